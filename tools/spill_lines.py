#!/usr/bin/env python
"""Attribute local-memory (spill) instructions of one kernel to source lines.
   usage: spill_lines.py <object.o> <mangled kernel substring>"""
import subprocess, sys, os, tempfile, collections, re
obj, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
inside, line, cnt, inl = False, None, collections.Counter(), None
for t in txt:
    if t.startswith("//---") and ".text." in t:
        inside = kern in t
    if not inside:
        continue
    m = re.search(r'//## File ".*?", line (\d+)(.*)', t)
    if m:
        line = int(m.group(1)); inl = m.group(2)
        continue
    if re.search(r"\b(STL|LDL)\b", t):
        cnt[(line, "STL" if "STL" in t else "LDL")] += 1
for (ln, k), c in sorted(cnt.items()):
    print(ln, k, c)
