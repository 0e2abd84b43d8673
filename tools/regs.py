#!/usr/bin/env python
"""Registers / spills per kernel from the ptxas logs of the last build (fots/pytorch_b200/csrc/build/*.ptxas.log).

    python tools/regs.py [translation unit] [kernel-name filter]
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
unit = sys.argv[1] if len(sys.argv) > 1 else "*"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for path in sorted(glob.glob(os.path.join(ROOT, "fots/pytorch_b200/csrc/build", unit + ".ptxas.log"))):
    cur, sp = None, ""
    for line in open(path):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            sp = "spill %s/%s" % (m.group(1), m.group(2)) if m.group(1) != "0" or m.group(2) != "0" else ""
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            if flt in cur:
                print("%4s regs %-16s %s" % (m.group(1), sp, cur.split("(")[0][:110]))
            cur = None
