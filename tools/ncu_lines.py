"""Joins the per-SASS-instruction stall samples of an ncu report (--page source --csv) with the source lines nvdisasm -g
prints for the same cubin, and aggregates samples per source line (the CLI source page has no CUDA-C view for this
report).  usage: ncu_lines.py <report.ncu-rep> <lib.so> <cubin name in lib> <kernel mangled substring> [top N]"""
import csv
import re
import subprocess
import sys
import collections

rep, lib, cubin, kern = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
subprocess.run(["cuobjdump", "-xelf", cubin, lib], cwd="/tmp", check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/" + cubin], capture_output=True, text=True).stdout.splitlines()
# instruction list of the kernel with the current (file, line)
ins = []
inside = False
cur = ("?", 0)
for ln in dis:
    if ln.startswith("\t.section") or ln.startswith(".section"):
        inside = (".text." in ln) and (kern in ln)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[6] if len(sys.argv) > 6 else "regex:recur"],
                        capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
h = rows[1]
body = rows[2:]
iN = h.index("# Samples"); iS = h.index("Source"); iE = h.index("Instructions Executed")
stall_cols = [(k, h.index(k)) for k in h if k.startswith("stall_") and "Not Issued" not in k]
print("sass in report %d, sass in disassembly %d" % (len(body), len(ins)))
n = min(len(body), len(ins))
mism = sum(1 for i in range(n) if body[i][iS].split()[0:1] != ins[i][1].split()[0:1] and not body[i][iS].startswith("@"))
print("opcode mismatches (unpredicated):", mism)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
total = 0
for i in range(n):
    s = int(body[i][iN] or 0)
    total += s
    a = agg[ins[i][2]]
    a[0] += s
    a[1] += int(body[i][iE] or 0)
    for k, c in stall_cols:
        v = int(body[i][c] or 0)
        if v:
            a[2][k[6:]] += v
print("total samples", total)
for key, (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-12s:%-5d %6d (%4.1f%%) inst %9d  %s" % (key[0], key[1], s, 100.0 * s / total, e, " ".join("%s=%d" % kv for kv in st.most_common(4))))
