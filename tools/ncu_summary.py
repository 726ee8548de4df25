#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python tools/ncu_summary.py kernel  gpurun_out/prof_r01_X.ncu-rep profiles/r01_X.md
  python tools/ncu_summary.py traffic profiles/ncu_traffic.json name=rep [name=rep ...]   (DRAM bytes per launch)
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            rows.append((r["Kernel Name"].split("(")[0], v * scale))
    agg = OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): {len(rows)} launches, {tot:.2f} ms total (cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {ms / n:.4f} |\n")
    print("wrote", dst)


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            f.write(f"\n## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[h]) for h in hdr
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h]}
            tot = sum(st.values()) or 1.0
            f.write("\nstall reasons (pc sampling): " + ", ".join(
                f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    print("wrote", dst)


def _to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def traffic(dst, *pairs):
    """Rewrite the `kernels` table of profiles/ncu_traffic.json (bench.py's roofline.traffic) from --set full captures."""
    import json
    doc = json.load(open(dst))
    srcs = []
    for pair in pairs:
        name, rep = pair.split("=", 1)
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        doc["kernels"][name] = {"read": _to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
                                "write": _to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])}
        srcs.append(rep)
    doc["source"] = ", ".join(srcs)
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", dst, doc["kernels"])


if __name__ == "__main__":
    fn = {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]]
    fn(*sys.argv[2:])
