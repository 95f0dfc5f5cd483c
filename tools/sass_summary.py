"""Counts the Blackwell-native SASS mnemonics per kernel of libkprn_b200.so (cuobjdump -sass) -> profiles/sass_summary.txt.
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "kprn_b200", "libkprn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG.2D.GATHER4", "UTMALDG", "UTCBAR", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "MUFU.RCP",
        "STG.E.ENL2.256", "RED", "ATOMG", "HMMA"]
cnt = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cnt.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op.startswith(p):
            cur[p] += 1
            break
with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
    f.write("SASS mnemonic counts per kernel, `cuobjdump -sass kprn_b200/libkprn_b200.so` (tools/sass_summary.py); sm_100a.\n"
            "UTCHMMA[.2CTA] = tcgen05.mma [cta_group::2]; LDTM/STTM = tcgen05.ld/st; UBLKCP = cp.async.bulk (1-D TMA); UTMALDG = cp.async.bulk.tensor\n"
            "(UTMALDG.2D.GATHER4 = tile::gather4); UTCBAR = tcgen05.commit; FFMA2/FADD2/FMUL2 = packed f32x2; HMMA (legacy mma.sync) must be 0.\n\n")
    f.write("%-46s %s\n" % ("kernel", " ".join("%s" % p for p in pats)))
    for name, c in cnt.items():
        if not any(c[p] for p in pats):
            continue
        f.write("%-46s %s\n" % (name[:46], " ".join("%*d" % (len(p), c[p]) for p in pats)))
print(open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read())
