#!/usr/bin/env python
"""Count the tensor-core / TMA mnemonics per kernel in the built library (no GPU needed):

  python tools/sass_evidence.py > profiles/r02_sass_evidence.md
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "modular_rl_b200", "libmrl_b200.so")
WHAT = [("UTCHMMA", "tcgen05.mma"), ("LDTM", "tcgen05.ld"), ("STTM", "tcgen05.st"), ("UBLKCP", "cp.async.bulk"),
        ("UTCBAR", "tcgen05.commit"), ("USETMAXREG", "setmaxnreg"), ("SYNCS", "mbarrier ops"), ("HMMA", "legacy mma.sync"),
        ("NANOSLEEP", "nanosleep"), ("UCGABAR", "cluster barrier")]
KEEP = ("fvp_tc_kernel", "l1_forward_tc_kernel", "l1_grad_tc_kernel", "chain_fwd_kernel", "chain_bwd_kernel",
        "reduce_partials_kernel", "cg_step_cluster_kernel", "gae_kernel", "population_forward_kernel", "tc_peak")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name:
            funcs[name].append(line)
    print("# SASS evidence (`cuobjdump -sass modular_rl_b200/libmrl_b200.so`, made by `tools/sass_evidence.py`)\n")
    print("Mnemonics of `/opt/skills/guides/B200_PROFILING.md`: " + ", ".join("`%s` = %s" % w for w in WHAT) + ".\n")
    print("| kernel | " + " | ".join(w[0] for w in WHAT) + " |\n|---|" + "---:|" * len(WHAT))
    firsts = {}
    for fn in sorted(funcs):
        if not any(k in fn for k in KEEP):
            continue
        body = funcs[fn]
        counts = [sum(1 for ln in body if re.search(r"\b%s" % w[0], ln.split("/*")[1] if ln.count("/*") > 1 else ln)) for w in WHAT]
        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("| `%s` | " % short + " | ".join(str(c) for c in counts) + " |")
        mma = [ln.strip() for ln in body if "UTCHMMA" in ln][:3]
        if mma:
            firsts[short] = mma
    print()
    for k, v in firsts.items():
        print("First `tcgen05.mma` instructions of `%s`:\n```\n%s\n```\n" % (k, "\n".join(v)))


if __name__ == "__main__":
    sys.exit(main())
