"""Summarise registers / spills per kernel from the ptxas logs written by ramannoodle_b200/_build.py."""
import glob
import os
import re
import subprocess
import sys

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ramannoodle_b200", "csrc", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(log).read()
    items = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt)
    for name, stack, ss, sl, regs in items:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void rn::", "")
        if pat in dem:
            print(f"{dem[:60]:60s} regs={regs:>3s} stack={stack:>3s} spill={ss}/{sl}")
