#!/usr/bin/env python
"""Top source lines of an ncu report by warp-stall samples (needs -lineinfo and --import-source on).

  python tools/ncu_hot_lines.py gpurun_out/prof_X.ncu-rep [file-substring] [top-n]
"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ".cu"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        cur = {"file": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]) and r[0] != "":
        cur["rows"].append(r)   # source-line rows only (SASS rows have an empty line number)
STALLS = ["stall_long_sb", "stall_wait", "stall_math", "stall_barrier", "stall_short_sb", "stall_mio", "stall_lg",
          "stall_dispatch", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_branch_resolving"]
for s in sections:
    if want not in s["file"] or not s["rows"]:
        continue
    h = s["hdr"]
    ix = {n: h.index(n) for n in ["Line No", "Source", "# Samples", "Instructions Executed"] + STALLS if n in h}
    STALLS = [k for k in STALLS if k in ix]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in s["rows"])
    print(f"== {s['file']}: {tot} samples")
    for r in sorted(s["rows"], key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
        n = int(r[ix["# Samples"]] or 0)
        st = sorted(((int(r[ix[k]] or 0), k[6:]) for k in STALLS), reverse=True)[:3]
        print(f"{r[ix['Line No']]:>5} {100.0 * n / max(tot, 1):5.1f}% inst={r[ix['Instructions Executed']]:>10} "
              f"{' '.join(f'{k}:{v}' for v, k in st if v):40s} | {r[ix['Source']].strip()[:90]}")
