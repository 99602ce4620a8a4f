#!/usr/bin/env python
"""SASS opcode histogram of the kernels in libescort_b200.so (cuobjdump -sass), one line per kernel: the mnemonics that
tell which hardware path a kernel uses (UTCHMMA = tcgen05.mma kind::tf32, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, LDGSTS = cp.async,
SYNCS = mbarrier, FFMA2 = packed fp32 FMA, USETMAXREG = setmaxnreg).  python tools/sass_hist.py [regex] > profiles/..."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "caffe_escoin_b200", "libescort_b200.so")
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["FFMA2", "FFMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "LDGSTS", "SYNCS", "BAR", "USETMAXREG", "LDS", "STS", "STG", "ATOMG", "RED", "SHFL", "R2UR", "BRA"]
name, cnt, rows = None, Counter(), []


def flush():
    if name and (pat is None or pat.search(name)):
        rows.append((name, sum(cnt.values()), dict(cnt)))


for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, cnt = m.group(1), Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m:
        cnt[m.group(1).split(".")[0]] += 1
flush()
demangle = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines() if rows else []
print("%-72s %6s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
for (n, tot, c), d in zip(rows, demangle or [r[0] for r in rows]):
    d = (d[:d.rindex(">") + 1] if ">" in d else re.sub(r"\(.*", "", d)).replace("escort::", "").replace("void ", "").replace(" ", "")
    print("%-72s %6d " % (d[:72], tot) + " ".join("%7d" % c.get(k, 0) for k in KEYS))
