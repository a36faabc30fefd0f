"""SASS opcode histogram per kernel of an object file (cuobjdump -sass): evidence for the tensor-core / TMA / cluster claims.

  python tools/sass_histogram.py gc-slam_b200/lib/gcs_bins_tc.o bin_scan_tc_kernel > profiles/r02_sass_bin_scan_tc.txt
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if pat in name else None
        if cur:
            hist[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            hist[cur][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "CREDUX", "REDUX", "MATCH", "UCGABAR", "MUFU", "FFMA2", "F2FP", "HADD2")
for name, h in hist.items():
    print(f"== {name[:150]}\n   {sum(h.values())} instructions")
    fam = collections.Counter()
    for op, c in h.items():
        for k in KEY:
            if op.startswith(k):
                fam[k] += c
    print("   families of interest:", ", ".join(f"{k} x{c}" for k, c in sorted(fam.items(), key=lambda x: -x[1])))
    for op, c in h.most_common(40):
        print(f"   {op:36s}{c:6d}")
