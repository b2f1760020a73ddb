#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libdpm_b200.so (cuobjdump -sass), restricted to the opcodes that tell which
hardware path a kernel uses: UTCHMMA (tcgen05.mma), UTCBAR / UTCCP, LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk),
UTMALDG / UTMASTG (tensor TMA), HMMA (mma.sync), REDUX, SYNCS (mbarrier), MUFU, F2FP / I2F conversions.
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "deeppointmap_b200", "libdpm_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "REDUX", "SYNCS", "MUFU",
       "LDGSTS", "ATOMS", "ATOMG", "RED", "BAR", "LDS", "STS", "LDG", "STG", "SHFL", "MATCH", "VOTE", "F2F", "FFMA", "DFMA")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, hist, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        hist[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEY:
            if op.startswith(k):
                hist[cur][k] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} -- instructions per kernel, opcodes of interest (prefix match)")
agg = collections.Counter()
for k, c in hist.items():
    agg.update(c)
print("# whole library: " + ", ".join(f"{k} {v}" for k, v in agg.items() if k in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "UTCBAR", "SYNCS")))
for k, c in hist.items():
    if total[k] == 0:
        continue
    print(f"{k}\n    {total[k]} instructions: " + ", ".join(f"{o} {n}" for o, n in c.most_common()))
