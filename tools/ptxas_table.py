#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output (stdin): kernel, registers, stack bytes."""
import re, subprocess, sys
name = None
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        stack = 0
        continue
    m = re.search(r"(\d+) bytes stack frame", line)
    if m:
        stack = int(m.group(1))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        print(f"{int(m.group(1)):4d} regs {stack:6d} B stack  {name}")
        name = None
