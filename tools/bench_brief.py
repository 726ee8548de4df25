#!/usr/bin/env python
"""Print the per-kernel table of a bench.py JSON line read from stdin (experiment helper)."""
import json, sys
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print("ms_per_step %.2f  value %.3fM  e2e %s" % (d["ms_per_step"], d["value"] / 1e6,
      ("%.2f ms" % d["e2e"]["ms_per_step"]) if d.get("e2e") else None))
for k, v in d["kernels"].items():
    print("  %-18s %8.3f ms/step  avg %.4f ms  %6.2f TF" % (k, v["ms_per_step"], v["avg_ms"], v.get("algo_tflops", 0)))
