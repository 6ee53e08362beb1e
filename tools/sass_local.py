#!/usr/bin/env python3
"""Development: which source lines of a kernel carry its thread-local loads / stores (LDL / STL)?
usage: sass_local.py <lib.so> <kernel-name-substring> [top]   (the library must be built with -lineinfo)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
cur_fn, cur_line, n_ins = None, None, collections.Counter()
loc = collections.defaultdict(collections.Counter)
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur_fn and pat in cur_fn and re.search(r"/\*[0-9a-f]{4,6}\*/", l):
        n_ins[cur_fn] += 1
        mm = re.search(r"\b(LDL|STL)\b", l)
        if mm:
            loc[cur_fn][(cur_line, mm.group(1))] += 1
for fn, c in loc.items():
    print(f"== {fn}: {n_ins[fn]} instructions, {sum(v for (k, op), v in c.items() if op == 'LDL')} LDL, {sum(v for (k, op), v in c.items() if op == 'STL')} STL")
    for (line, op), v in c.most_common(top):
        print(f"   {op} x{v:3d}  {line}")
