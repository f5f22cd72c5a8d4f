"""Condense `ncu --page source --csv` of one kernel into blocks of SASS instructions with executed-instruction and stall-sample
totals: python tools/ncu_hot_blocks.py src.csv [block]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
ia, isrc, ismp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
body = rows[2:]
tot_ex = sum(int(r[iex]) for r in body)
tot_s = sum(int(r[ismp]) for r in body)
print('instructions executed (warp level) {:,}   samples {:,}'.format(tot_ex, tot_s))
for b in range(0, len(body), blk):
    seg = body[b:b + blk]
    ex = sum(int(r[iex]) for r in seg)
    sm = sum(int(r[ismp]) for r in seg)
    ops = {}
    for r in seg:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith('@') else r[isrc].split()[1]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
    top = ' '.join('{}x{}'.format(k, v) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print('{:5d} ex {:5.1f}%  samples {:5.1f}%   {}'.format(b, 100.0 * ex / tot_ex, 100.0 * sm / max(tot_s, 1), top))
