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
names = ["P:empty ok", "P:issued", "C:full ok", "C:conv done", "M:start", "-", "M:conv ok", "M:committed"]
t0 = out[1, 0]
print("use " + " ".join("%12s" % n for n in names if n != "-"))
for u in range(40, 64):
    print("%3d " % u + " ".join("%12d" % (out[e, u] - t0) for e in range(8) if e != 5))
R = slice(32, 96)
print("mean cycles per stage (MMA commit to commit):", np.diff(out[7, R]).mean())
print("P issued -> C full ok (copy latency):", (out[2, R] - out[1, R]).mean())
print("C full ok -> conv done (convert into tensor memory):", (out[3, R] - out[2, R]).mean())
print("C conv done -> M conv ok:", (out[6, R] - out[3, R]).mean())
print("M start -> conv ok (wait):", (out[6, R] - out[4, R]).mean())
print("M conv ok -> committed (issue of the stage's MMAs):", (out[7, R] - out[6, R]).mean())
print("M committed(u) -> P empty ok(u + 5 stages):", (out[0, 37:101] - out[7, R]).mean())
print("P empty ok -> issued:", (out[1, R] - out[0, R]).mean())
