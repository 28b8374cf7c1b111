#!/usr/bin/env python
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki = H.index('Kernel Name'); vi = H.index('Metric Value'); gi = H.index('Grid Size'); bi = H.index('Block Size')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    k = (r[ki][:60], r[gi], r[bi])
    agg.setdefault(k, []).append(float(r[vi].replace(',', '')))
for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:62s} grid={g:>12s} block={b:>12s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us min={min(v)/1e3:8.2f} max={max(v)/1e3:8.2f}")
