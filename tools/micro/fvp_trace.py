"""Timeline of CTA 0 of fvp_tc_kernel (MRL_FVP_TC_TRACE=1): per role, clock64 stamps of the hand-overs."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["MRL_FVP_TC_TRACE"] = "1"
from modular_rl_b200 import _lib as L, synth
from modular_rl_b200.device import DeviceBatch, DeviceNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dims = (376, 100, 50, 25, 17)
rng = np.random.default_rng(0)
ob = synth.make_obs(N, dims[0], rng)
theta = synth.init_params(dims, synth.GAUSS, rng)
net = DeviceNet(dims, synth.GAUSS)
batch = DeviceBatch(dims[0], with_time_feature=False)
batch.set_obs(ob)
net.set_params(theta)
v = rng.standard_normal(net.P).astype(np.float32)
for _ in range(3):
    net.fvp(batch, v)
tr = np.zeros((8, 4096, 2), np.int64)
L.lib().mrl_debug_fvp_tc_trace(net._h, tr.ctypes.data_as(C.c_void_p))   # clears
net.fvp(batch, v)
L.check(L.lib().mrl_debug_fvp_tc_trace(net._h, tr.ctypes.data_as(C.c_void_p)))
names = ["R-mma", "D-mma", "K-mma", "R-epi", "D-epi", "conv", "-", "-"]
t0 = min(int(tr[r, 0, 0]) for r in range(6) if tr[r, 0, 0])
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "gpurun_out", "fvp_trace.npy"), tr)
for r in range(6):
    ev = tr[r]
    n = int((ev[:, 0] != 0).sum())
    print("== role", names[r], "events", n)
    # print the events between the start of the 3rd and 5th tile of this role (steady state)
    lines = []
    for i in range(n):
        t, code = int(ev[i, 0]) - t0, int(ev[i, 1])
        lines.append((t, code >> 24, (code >> 8) & 0xffff, code & 0xff))
    # tile boundaries: first event of each tile = (kind 1, stage/l = first, kg 0)
    first = lines[0][1:] if lines else None
    starts = [i for i, x in enumerate(lines) if x[1:] == first]
    print("  tile starts (cycles):", [lines[i][0] for i in starts[:8]])
    if len(starts) > 4:
        a, b = starts[2], starts[4]
        prev = lines[a][0]
        for t, k, s, g in lines[a:b]:
            print("   t=%8d (+%5d) ev%d stage/l=%d kg/c=%d" % (t, t - prev, k, s, g))
            prev = t
