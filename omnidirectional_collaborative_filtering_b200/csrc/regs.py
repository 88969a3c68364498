#!/usr/bin/env python
"""Summarise `ptxas -v` output in build.log: registers / spills / smem per kernel."""
import re, subprocess, sys
txt = open(sys.argv[1] if len(sys.argv) > 1 else "build.log").read()
pat = re.compile(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")
for name, stack, ss, sl, regs, smem in pat.findall(txt):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    print("%4s regs %6s smem  stack %-4s spill %s/%s  %s" % (regs, smem or 0, stack, ss, sl, dem.split("(")[0][:90]))
