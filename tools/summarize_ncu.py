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


if __name__ == "__main__" and sys.argv[1] in ("launches", "full"):
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])


# ------------------------------------------------------------------------------------------------ round 2 additions
def _source_kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    return kern


def roles(src, dst, samples=None):
    """Per captured launch: headline metrics (raw page) + where the issued warp instructions go (source page): static
    instructions bucketed by execution count separate the warp roles (epilogue / producers / MMA issuer) and expose
    barrier-polling loops; plus the stall-reason mix."""
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    keep = [k for k in KEEP + ["sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                               "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"] if k in H]
    kern = _source_kernels(src)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --import-source on: {src}\n\n"
                "Times under ncu are cold-cache and serialised; bench.py's CUDA-event numbers are the reported ones.\n")
        for n, r in enumerate(rows[2:]):
            name = r[H.index("Kernel Name")].split("(")[0].replace("void ", "").replace("b200rl::", "")
            f.write(f"\n## launch {n}: `{name}`\n\n| metric | value |\n|---|---|\n")
            for k in keep:
                f.write(f"| {k} | {r[H.index(k)]} |\n")
            norm = lambda z: z.replace("(int)", "").replace("(bool)", "").split("(")[0].replace("void ", "").replace("b200rl::", "")
            k = next((kk for kk in kern if norm(kk["name"]) == norm(name) and len(kk["rows"]) > 2), None)
            if k is not None:
                HH = k["rows"][0]
                ie, ism, isrc = HH.index("Instructions Executed"), HH.index("# Samples"), HH.index("Source")
                data = [(rr[isrc], int(rr[ie]), int(rr[ism] or 0)) for rr in k["rows"][1:] if len(rr) > ie and rr[ie].isdigit()]
                tot, tots = sum(d[1] for d in data) or 1, sum(d[2] for d in data) or 1
                b = collections.OrderedDict()
                for s_, e, sm in data:
                    if e:
                        x = b.setdefault(float(f"{e:.2g}"), [0, 0, 0, ""])
                        x[0] += 1; x[1] += e; x[2] += sm
                        if "TRYWAIT" in s_ or "BAR.SYNC" in s_:
                            x[3] = "barrier polling / sync"
                f.write(f"\nwarp instructions executed: {tot}; pc samples: {tots}\n\n"
                        "| executions per static instruction | static instrs | executed | share | pc samples | note |\n"
                        "|---:|---:|---:|---:|---:|---|\n")
                for key, (cnt, e, sm, note) in sorted(b.items(), key=lambda kv: -kv[1][1])[:8]:
                    note = note if cnt <= 6 else "warp-role body (epilogue / producer / MMA issuer)"
                    f.write(f"| ~{key:.0f} | {cnt} | {e} | {100 * e / tot:.1f}% | {100 * sm / tots:.1f}% | {note} |\n")
                st = [h for h in HH if h.startswith("stall_") and "Not Issued" not in h]
                acc = {h: 0 for h in st}
                for rr in k["rows"][1:]:
                    for h in st:
                        i = HH.index(h)
                        if len(rr) > i and rr[i].isdigit():
                            acc[h] += int(rr[i])
                ts = sum(acc.values()) or 1
                f.write("\nstall reasons (pc sampling): " + ", ".join(f"{h[6:]} {100 * v / ts:.0f}%" for h, v in
                                                                       sorted(acc.items(), key=lambda kv: -kv[1])[:7]) + "\n")


def traffic(src, dst, samples):
    """profiles/r2_traffic.json: DRAM bytes per sample of every captured kernel (dram__bytes_read + write of an
    `ncu --set full` capture whose launches processed `samples` samples each); bench.py scales it to its launch size."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    ir, iw, ik = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = json.load(open(dst)) if os.path.exists(dst) else {}
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("b200rl::", "")
        b = float(r[ir].replace(",", "")) * scale.get(U[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(U[iw], 1.0)
        out[name] = {"dram_bytes_per_sample": b / float(samples), "capture_samples": int(samples),
                     "src": f"{os.path.basename(src)} (ncu --set full --clock-control none)"}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] in ("roles", "traffic"):
    if sys.argv[1] == "roles":
        roles(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    sys.exit(0)
