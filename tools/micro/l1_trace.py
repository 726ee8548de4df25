#!/usr/bin/env python
"""Timeline of the layer-1 forward pipeline of CTA 0 (needs a library built with -DMRL_TRACE)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from modular_rl_b200 import _lib as L, synth
from modular_rl_b200.device import DeviceBatch, DeviceNet
N = 262144
dims = (376, 100, 50, 25, 17)
rng = np.random.default_rng(0)
net = DeviceNet(dims, synth.GAUSS)
batch = DeviceBatch(dims[0], with_time_feature=False)
batch.set_obs(synth.make_obs(N, dims[0], rng))
net.set_params(synth.init_params(dims, synth.GAUSS, rng))
for _ in range(3):
    net.forward(batch)
lib = C.CDLL(L.LIB_PATH)
out = np.zeros((8, 512), np.int64)
assert lib.mrl_debug_l1_trace(out.ctypes.data_as(C.c_void_p)) == 0
names = ["P:empty ok", "P:issued", "C:full ok", "C:conv done", "M:start", "M:full ok", "M:conv ok", "M:committed"]
t0 = out[1, 0]
print("use " + " ".join("%12s" % n for n in names))
for u in range(32, 64):
    print("%3d " % u + " ".join("%12d" % (out[e, u] - t0) for e in range(8)))
print("fence.proxy.async cycles (warp 6):", (out[5, 32:96] - out[4, 32:96]).mean(), " convert body:", (out[4, 32:96] - out[2, 32:96]).mean())
d = np.diff(out[7, 32:96])
print("mean cycles per stage (MMA commit to commit):", d.mean())
print("P issued -> C full ok (copy latency):", (out[2, 32:96] - out[1, 32:96]).mean())
print("C full ok -> conv done (convert):", (out[3, 32:96] - out[2, 32:96]).mean())
print("C conv done -> M conv ok:", (out[6, 32:96] - out[3, 32:96]).mean())
print("M conv ok -> committed (issue):", (out[7, 32:96] - out[6, 32:96]).mean())
print("M committed(u) -> P empty ok(u + nstages=5):", (out[0, 37:101] - out[7, 32:96]).mean())
print("P empty ok -> issued:", (out[1, 32:96] - out[0, 32:96]).mean())
