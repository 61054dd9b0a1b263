#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/r1_x.md
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    unit = None
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = unit or r[ui]
        name = r[ki].split("(")[0].replace("void ", "").replace("b200rl::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"launches: {sum(v[0] for v in agg.values())}, total kernel time {tot * scale:.1f} ms\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] * scale:.2f} | {100 * v[1] / tot:.1f}% | {1e3 * v[1] * scale / v[0]:.1f} |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    cols = [(k, H.index(k)) for k in ["Kernel Name"] + KEEP if k in H]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\nunits: " +
                ", ".join(f"{k}=[{units[i]}]" for k, i in cols[1:]) + "\n\n")
        f.write("| # | " + " | ".join(k for k, _ in cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n")
        for n, r in enumerate(rows[2:]):
            vals = [r[i] for _, i in cols]
            vals[0] = "`" + vals[0].split("(")[0].replace("void ", "").replace("b200rl::", "")[:70] + "`"
            f.write(f"| {n} | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
