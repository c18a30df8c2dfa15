#!/usr/bin/env python
"""Summarise an ncu source-page CSV (ncu -i rep --page source --csv) by SASS regions: share of warp-stall samples, dominant
opcodes and stall reasons.  Usage: ncu_src_regions.py file.csv [window=80] [min_share=0.01]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 80
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = [r]; blocks.append(cur)
    elif cur is not None: cur.append(r)
seen = set()
for b in blocks:
    name = b[0][1]
    if name in seen: continue
    seen.add(name)
    h = b[1]; R = [r for r in b[2:] if len(r) == len(h)]
    iS = h.index('# Samples'); iE = h.index('Instructions Executed'); isrc = h.index('Source')
    tot = sum(int(r[iS]) for r in R)
    print('==', name[:70], 'samples', tot)
    stall = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    allst = {c: sum(int(r[h.index(c)]) for r in R) for c in stall}
    print('   overall:', [(k[6:], round(100 * v / tot, 1)) for k, v in sorted(allst.items(), key=lambda x: -x[1])[:8]])
    for i in range(0, len(R), win):
        seg = R[i:i + win]; s = sum(int(r[iS]) for r in seg); e = sum(int(r[iE]) for r in seg)
        if s / tot < minshare: continue
        def op(r):
            t = r[isrc].split()
            return t[1] if t[0].startswith('@') else t[0]
        ops = collections.Counter(op(r) for r in seg)
        st = {c: sum(int(r[h.index(c)]) for r in seg) for c in stall}
        top = sorted(st.items(), key=lambda x: -x[1])[:4]
        print(f'   {i:5d} {100*s/tot:5.1f}% exec={e:>11d}', dict(ops.most_common(4)), [(k[6:], round(100 * v / tot, 1)) for k, v in top])
