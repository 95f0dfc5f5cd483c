"""Turns ncu outputs in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/profile_summary.py <tag> <launches.csv> [<full.ncu-rep>]"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
reps = sys.argv[3:]
rep = reps[0] if reps else None
out = open(os.path.join(ROOT, "profiles", "%s_summary.md" % tag), "w")
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = defaultdict(list)
for r in rows[1:]:
    if r[mi] == "gpu__time_duration.sum":
        agg[r[ki].split("(")[0][:48]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
out.write("# %s -- ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), python bench.py\n\n" % tag)
out.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
out.write("| kernel | launches | avg us | share |\n|---|---|---|---|\n")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.write("| %s | %d | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
if rep:
    rr = None
    for rp in reps:                     # several reports (e.g. a scoring and a training capture): rows concatenated
        raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        part = list(csv.reader(raw.splitlines()))
        if rr is None: rr = part; continue
        # align by column name and convert to the first report's units (ncu picks a unit per column per report)
        US = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
              "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
        idx = {n: i for i, n in enumerate(part[0])}
        for r in part[2:]:
            row = []
            for ci, n in enumerate(rr[0]):
                if n not in idx: row.append(""); continue
                v, u_src, u_dst = r[idx[n]], part[1][idx[n]], rr[1][ci]
                if u_src != u_dst and u_src in US and u_dst in US:
                    try: v = "%.6f" % (float(v.replace(",", "")) * US[u_src] / US[u_dst])
                    except ValueError: pass
                row.append(v)
            rr.append(row)
    h = rr[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum", "lts__cycles_elapsed.avg", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    out.write("\n## `ncu --set full` capture (one launch per kernel)\n\n| metric | unit | " + " | ".join(r[h.index("Kernel Name")].split("(")[0][:24] for r in rr[2:]) + " |\n")
    out.write("|---|---|" + "---|" * len(rr[2:]) + "\n")
    for i, name in enumerate(h):
        if name in want and name != "Kernel Name":
            out.write("| %s | %s | %s |\n" % (name, rr[1][i], " | ".join(r[i] for r in rr[2:])))
    # machine-readable DRAM traffic per launch for bench.py's roofline.traffic (it reads the newest profiles/*_traffic.json)
    import json
    ni, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kern = {}
    for r in rr[2:]:
        nm = r[ni].split("(")[0]
        # recur_kernel<HP, PAIR, TRAIN>: the scoring kernel is "recur", the training forward "recur_train"
        if "recur" in nm:
            key = "recur_train" if nm.replace(" ", "").endswith(",1>") else "recur"
        else:
            key = "proj" if "proj" in nm else nm.replace("void ", "").replace("kprn::", "")[:32]
        if key in kern:
            continue                    # first capture of a name wins (later launches of the same kernel are other steps)
        kern[key] = {"name": nm[:64], "dram_read_bytes": float(r[ri].replace(",", "")) * scale.get(rr[1][ri], 1.0),
                     "dram_write_bytes": float(r[wi].replace(",", "")) * scale.get(rr[1][wi], 1.0)}
    json.dump({"source": "profiles/%s_summary.md (ncu --set full, one launch per kernel)" % tag, "kernels": kern},
              open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
out.close()
print(open(os.path.join(ROOT, "profiles", "%s_summary.md" % tag)).read())
