"""SASS evidence that the contraction kernels are Blackwell-native: per kernel of libtortto_b200.so (sm_100a cubin), the
count of tcgen05 / TMA / TMEM instructions.  UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load; .IM2COL = the
im2col-mode tensor maps), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS/UTCALLOC-class = TMEM allocation,
SYNCS = mbarrier ops.  Usage: python scripts/sass_summary.py [object-or-so ...] > profiles/r2_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
files = sys.argv[1:] or sorted(glob.glob(os.path.join(HERE, "..", "pytortto_b200", "_build", "*.o")))
PATTERNS = ["UTCHMMA", "UTMALDG", "IM2COL", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "ELECT", " HMMA", "FFMA"]  # (" HMMA" = mma.sync)
print("kernel".ljust(78), " ".join(p.rjust(8) for p in PATTERNS))
tot = collections.Counter()
for f in files:
    out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
    name, counts = None, None
    rows = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), collections.Counter()
            continue
        if name and "/*" in line:
            for p in PATTERNS:
                if p in line:
                    counts[p] += 1
    if name:
        rows.append((name, counts))
    for name, counts in rows:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0].replace("void ", "")
        if not any(counts[p] for p in PATTERNS[:7]):
            continue  # only the kernels that use tensor cores / TMA / mbarriers
        print(d[:78].ljust(78), " ".join(str(counts[p]).rjust(8) for p in PATTERNS))
        tot.update(counts)
print("TOTAL".ljust(78), " ".join(str(tot[p]).rjust(8) for p in PATTERNS))
