#!/usr/bin/env python
"""Operand-bandwidth model of an FP64 instruction stream (profiling aid).  Reads the compact SASS listing of tools/sass.sh on stdin
and, for every maximal run of >= MINRUN instructions without a branch, reports: FP64 instructions, the cycles the FP64 pipe needs
(2 per warp instruction) and the cycles the register file needs when a DFMA/DMUL/DADD whose three / two sources are all fresh
vector-register reads costs one cycle per 64-bit source (tools/rf_probe.cu: 3 fresh sources -> 2/3 of the peak rate); a source is
not fresh when the previous FP64 instruction held the same register in the same slot with .reuse, or when it is RZ / a constant /
a uniform register.  usage: tools/sass.sh <kernel> | python tools/sass_operand_model.py [MINRUN]"""
import re, sys
minrun = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rows = []
for line in sys.stdin:
    parts = line.strip().split(None, 1)
    if len(parts) < 2:
        continue
    rows.append((parts[0], parts[1].rstrip(' ;')))
runs, cur = [], []
for a, ins in rows:
    op = ins.split()[0] if not ins.startswith('@') else ins.split()[1]
    if op.startswith(('BRA', 'EXIT', 'RET', 'BSYNC', 'BSSY', 'WARPSYNC', 'CALL')):
        if len(cur) >= minrun:
            runs.append(cur)
        cur = []
    else:
        cur.append((a, ins))
if len(cur) >= minrun:
    runs.append(cur)
for run in runs:
    prev = {}
    n64 = pipe = rf = 0
    other = {}
    for a, ins in run:
        toks = ins.split(None, 1)
        if toks[0].startswith('@'):
            toks = toks[1].split(None, 1)
        op = toks[0]
        if op.split('.')[0] in ('DFMA', 'DMUL', 'DADD'):
            srcs = [s.strip() for s in toks[1].split(',')][1:]
            fresh = 0
            cur_reuse = {}
            for slot, s in enumerate(srcs):
                reg = s.lstrip('-|').rstrip('|')
                reuse = reg.endswith('.reuse')
                reg = reg.replace('.reuse', '')
                if re.fullmatch(r'R\d+', reg):
                    if prev.get(slot) != reg:
                        fresh += 1
                    if reuse:
                        cur_reuse[slot] = reg
            prev = cur_reuse
            n64 += 1; pipe += 2; rf += max(2, fresh)
        else:
            k = op.split('.')[0]
            other[k] = other.get(k, 0) + 1
    if n64 >= minrun // 2:
        print("%s..%s: %4d instr, %4d FP64 -> pipe %4d cyc, register file %4d cyc, bound %.1f%% of FP64 peak; other: %s" %
              (run[0][0], run[-1][0], len(run), n64, pipe, rf, 100.0 * pipe / max(rf, 1), dict(sorted(other.items(), key=lambda kv: -kv[1])[:6])))
