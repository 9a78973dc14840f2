#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels (cuobjdump on the built objects; no GPU needed):

    python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt

Shows that the contraction is Blackwell-native (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM / STTM = tcgen05.ld /
st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch) and what the epilogues are made of.
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "MUFU", "SYNCS", "RED", "ATOM", "ATOMG", "REDG", "UTCATOMSWS", "CCTL")
INS = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)")


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


for obj in ("build/chain.o", "build/dw.o", "build/aux.o"):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, obj)], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            hist[fn] = collections.Counter()
            continue
        m = INS.match(line)
        if m and fn:
            op = m.group(1).rstrip(";")
            base = op.split(".")[0]
            hist[fn][op if base in FULL else base] += 1
    for fn, c in hist.items():
        tot = sum(c.values())
        if tot < 400:
            continue
        print(f"## {obj} :: {demangle(fn)}  ({tot} SASS instructions)")
        special = [(k, v) for k, v in sorted(c.items()) if k.split(".")[0] in FULL]
        print("   tensor / TMEM / bulk-copy / MUFU / reduction opcodes: " + ", ".join(f"{k} x{v}" for k, v in special))
        print("   most frequent: " + ", ".join(f"{k} x{v}" for k, v in c.most_common(16)))
        print()
