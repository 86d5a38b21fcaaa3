#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output (registers / spills per kernel).  Usage: ptxas_summary.py LOG [filter]"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", txt)
for name, stack, spill, regs in ents:
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("pfx::", "").replace("void ", "")
    if flt in dem:
        print(f"{regs:>4} regs stack {stack:>4} spill {spill:>4}  {dem}")
