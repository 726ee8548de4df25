"""ZFilter block scan on a Humanoid-sized block (for ncu): 262144 x 376 float64 observations."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from modular_rl_b200 import filters
N, d = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 376
X = np.random.default_rng(0).standard_normal((N, d))
for _ in range(3):
    t0 = time.perf_counter()
    y, st = filters.zfilter_scan(X, None, clip=5.0)
    print("zfilter_scan %d x %d: %.2f ms (host buffers, copies included)" % (N, d, 1e3 * (time.perf_counter() - t0)))
