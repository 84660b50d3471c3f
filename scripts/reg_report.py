"""Registers / stack / spills per kernel from the -Xptxas -v build logs (pytortto_b200/_build/*.log)."""
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(HERE, "..", "pytortto_b200", "_build", "*.cu.log"))):
    t = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores"
                         r".*?Used (\d+) registers", t, re.S):
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        d = re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0]
        if pat in d:
            print(f"{m.group(4):>4} regs  stack {m.group(2):>4}  spill {m.group(3):>4}  {d}")
