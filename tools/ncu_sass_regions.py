#!/usr/bin/env python3
"""Splits the SASS of the first kernel in `ncu --page source --csv --print-source sass` output into windows of W
instructions and prints, per window, the share of stall samples / executed instructions, the instruction mix and the
top stall reasons.  usage: ncu -i rep --page source --csv --print-source sass > x.csv; ncu_sass_regions.py x.csv [W]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 200
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
iS, iI, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
totI = sum(int(r[iI]) for r in data)
print(rows[0][1][:100])
print('samples', tot, 'warp instructions', totI, 'SASS instructions', len(data))
allst = {hdr[i]: sum(int(r[i]) for r in data) for i in stall_cols}
print('stall mix:', ' '.join(f"{k[6:]}:{v / tot:.3f}" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
for s in range(0, len(data), W):
    blk = data[s:s + W]
    sm = sum(int(r[iS]) for r in blk)
    ins = sum(int(r[iI]) for r in blk)
    ops = {}
    for r in blk:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    st = {hdr[i]: sum(int(r[i]) for r in blk) for i in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    key = [(k, ops.get(k, 0)) for k in ('SHFL', 'LDG', 'LDS', 'STS', 'STG', 'DMUL', 'DADD', 'DFMA', 'MUFU', 'CALL', 'BRA')]
    print(f"{s:5d} samp {sm / tot:6.3f} instr {ins / totI:6.3f}", ' '.join(f"{k}{v}" for k, v in key if v), '|', ' '.join(f"{k[6:]}:{v}" for k, v in top))
