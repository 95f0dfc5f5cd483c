"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples, with stall reasons."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
si = hdr.index("# Samples")
src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall") is False and "stall" in h.lower()]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot)
idx = {i: r for i, r in enumerate(body)}
order = sorted(idx, key=lambda i: -int(idx[i][si] or 0))[:top]
for i in sorted(order):
    r = idx[i]
    reasons = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols if r[c] not in ("", "0")), reverse=True)[:3]
    print("%5d  %5.1f%%  line %4d  %-70s %s" % (int(r[si]), 100.0 * int(r[si]) / max(tot, 1), i, r[src].strip()[:70],
                                              " ".join("%s=%d" % (n.replace("stall_", ""), v) for v, n in reasons)))
