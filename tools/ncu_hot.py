#!/usr/bin/env python
"""Top stall sites from an ncu report's source page (SASS view). usage: ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# may contain several kernels: split at "Kernel Name" rows
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; H = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        si = H.index("# Samples"); src = H.index("Source"); ie = H.index("Instructions Executed")
        stall_cols = [k for k, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si] or 0) for r in body)
        print(f"== {name}: {len(body)} SASS instrs, {tot} samples, {sum(int(r[ie] or 0) for r in body)} warp-instr")
        agg = {}
        for r in body:
            for k in stall_cols:
                agg[H[k]] = agg.get(H[k], 0) + int(r[k] or 0)
        print("   stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
        idx = sorted(range(len(body)), key=lambda k: -int(body[k][si] or 0))[:N]
        for k in sorted(idx):
            r = body[k]
            st = {H[c][6:]: int(r[c]) for c in stall_cols if int(r[c] or 0)}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"   #{k:5d} {int(r[si]):6d} ({100*int(r[si])/max(tot,1):4.1f}%) exec={r[ie]:>7s}  {r[src].strip()[:70]:70s} {top}")
        i = j
    else:
        i += 1
