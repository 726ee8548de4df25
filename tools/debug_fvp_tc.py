"""Stage-by-stage check of the tcgen05 Fisher-vector chain (mlp_fvp_tc.cu) against numpy: run with
MRL_FVP_TC_DEBUG=1 on a GPU box.  Prints the relative error of every stage's unit values on the first
128-timestep tile and of every parameter block of the product."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MRL_FVP_TC_DEBUG", "1")
from modular_rl_b200 import _lib as L, device, synth  # noqa: E402
from oracle import policy_math as pm  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run(dims, head, N, seed=3):
    print("=== dims", dims, "head", head, "N", N)
    spec = pm.NetSpec(tuple(dims), pm.GAUSS if head == 0 else pm.CAT)
    wl = synth.Workload("dbg", tuple(dims), head, N, 200, seed)

    def fwd(th, ob):
        _, z = pm.forward(th, spec, ob)
        return z if head == 0 else pm.softmax(z)
    d = synth.policy_batch(wl, fwd)
    theta = synth.perturb(d["theta"], 0.05, 5)
    net = device.DeviceNet(dims, head)
    batch = device.DeviceBatch(dims[0], True)
    batch.set_obs(d["ob"]).set_paths(np.array([0, N], np.int64), np.array([1], np.uint8), 1000.0)
    batch.set_policy_inputs(head, dims[-1], d["act"], d["adv"], d["oldprob"])
    net.set_params(theta)
    v = np.random.default_rng(7).standard_normal(net.P).astype(np.float32)
    f = net.fvp(batch, v)
    of = pm.fisher_vector_product(theta, spec, d["ob"], v)
    print("fvp rel err", rel(f, of))
    for name, shape, a, b in pm.param_slices(spec):
        print("   %-7s %-12s rel %.3e   |ref| %.3e" % (name, shape, rel(f[a:b], of[a:b]), np.linalg.norm(of[a:b])))
    out = np.zeros((8, 128, 128), np.float32)
    rc = L.lib().mrl_debug_fvp_tc_read(net._h, out.ctypes.data_as(C.c_void_p))
    if rc:
        print("no debug dump:", L.lib().mrl_last_error())
        return
    # numpy stages on the first 128 timesteps
    th = theta.astype(np.float64)
    Ws, bs, logstd = pm.split_params(th, spec)
    Vs, vbs, vls = pm.split_params(v.astype(np.float64), spec)
    n = min(128, N)
    hs, z = pm.forward(th, spec, d["ob"][:n])
    nl = spec.n_layers
    want = []
    Rh = np.zeros_like(hs[0])
    Rz = None
    for l in range(nl):
        Rz = Rh @ Ws[l] + hs[l] @ Vs[l] + vbs[l]
        if l < nl - 1:
            Rh = (1 - hs[l + 1] ** 2) * Rz
            want.append(("Rh%d" % (l + 1), Rh))
    if head == 0:
        delta = Rz / np.exp(2 * logstd)
    else:
        p = pm.softmax(z)
        delta = p * Rz - p * (p * Rz).sum(1, keepdims=True)
    want.append(("delta%d" % nl, delta))
    for l in range(nl - 1, 0, -1):
        delta = (delta @ Ws[l].T) * (1 - hs[l] ** 2)
        want.append(("delta%d" % l, delta))
    for s, (name, w) in enumerate(want):
        got = out[s, :n, :w.shape[1]]
        print("   stage %d %-7s rel %.3e  |ref| %.3e  max|got| %.3e" % (s, name, rel(got, w), np.linalg.norm(w), np.abs(got).max()))


if __name__ == "__main__":
    run([11, 64, 64, 3], 0, 300)
    run([376, 100, 50, 25, 17], 0, 1000)
    run([128, 64, 64, 18], 1, 2000)
