#!/usr/bin/env python
"""SASS opcode histogram per kernel of liboatgpu.so (cuobjdump -sass; CPU side, no GPU needed) -> text.
What to look for (B200_PROFILING.md): UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier, MEMBAR.ALL.GPU = the
release fence of the tile hand-off, LDG.E.STRONG.GPU + CCTL.IVALL = ld.acquire.gpu, FENCE.VIEW.ASYNC = proxy fences.
usage: tools/sass_histogram.py [lib] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oat_b200", "liboatgpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, order = None, {}, []
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        order.append(kern)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# SASS opcode histogram of {os.path.basename(lib)} (sm_100a), cuobjdump -sass; full mnemonics with modifiers")
KEY = ("UBLKCP", "UTMA", "SYNCS", "MEMBAR", "FENCE", "CCTL", "LDG.E.STRONG", "ATOMG", "REDG", "LDGSTS", "NANOSLEEP", "BPT", "ERRBAR", "ACQBULK", "UCGABAR")
for k in order:
    c = hist[k]
    total = sum(c.values())
    print(f"\n## {k}: {total} instructions")
    keys = {op: n for op, n in c.items() if any(op.startswith(p) for p in KEY)}
    if keys:
        print("   async copy / barrier / ordering: " + ", ".join(f"{op} x{n}" for op, n in sorted(keys.items())))
    fam = collections.Counter()
    for op, n in c.items():
        fam[op.split(".")[0]] += n
    print("   by opcode: " + ", ".join(f"{op} {n}" for op, n in fam.most_common(24)))
