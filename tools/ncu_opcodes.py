#!/usr/bin/env python
"""Executed-instruction histogram by opcode + top stall sites from an ncu source-page CSV.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; ncu_opcodes.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iW = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
tot = 0; cat = collections.Counter(); stall = collections.Counter(); data = []
for r in rows[2:]:
    if len(r) <= iE: continue
    try: n = int(r[iE]); w = int(r[iW])
    except ValueError: continue
    toks = r[iS].split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    cat[op] += n; stall[op] += w; tot += n; data.append((n, w, r[iS]))
print('total executed warp-instructions', tot)
for k, v in cat.most_common(24): print(f'{k:10s} {v:11d} {100*v/tot:5.1f}%  stall-samples {stall[k]}')
print('--- top stall sites')
for n, w, s in sorted(data, key=lambda x: -x[1])[:12]: print(w, n, s[:90])
