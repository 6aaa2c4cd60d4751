#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in libmvs_b200.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load, UBLKCP = bulk copy, UTCBAR = tcgen05.commit,
HMMA = warp-level mma.sync, SYNCS = mbarrier ops, plus register / stack use from the ELF symbol info.

  python tools/sass_summary.py [path/to/lib.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "self-supervised-mvs_b200", "libmvs_b200.so")
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "RED", "ATOMS", "ATOMG", "LDS", "LDG", "STG", "FFMA", "HFMA2", "FFMA2")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
names = iter(filt)
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(names)
        cur = re.sub(r"\(.*", "", cur).replace("(anonymous namespace)::", "")
        base = cur
        i = 2
        while cur in counts:
            cur = "%s #%d" % (base, i); i += 1
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
fn = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        fn = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = m.groups()
print("# %s" % os.path.relpath(lib, ROOT))
print("# instruction counts per kernel (static SASS, sm_100a); '_all' = every instruction")
print("%-78s %7s %s" % ("kernel", "_all", " ".join("%7s" % k for k in KEYS)))
for k, c in counts.items():
    if c["_all"] == 0:
        continue
    print("%-78s %7d %s" % (k[:78], c["_all"], " ".join("%7d" % c[x] for x in KEYS)))
    total.update(c)
print("%-78s %7d %s" % ("TOTAL", total["_all"], " ".join("%7d" % total[x] for x in KEYS)))
