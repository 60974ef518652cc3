#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts (ncu --page source --csv) to innermost source lines
using nvdisasm -gi line info.  usage: ncu_lines.py <source.csv> <nvdisasm -gi dump> <kernel substr> [ncells]"""
import csv, re, sys
from collections import defaultdict
src_csv, sass, kern = sys.argv[1:4]
cells = float(sys.argv[4]) if len(sys.argv) > 4 else 1024.0 ** 3
txt = open(sass).read().split('\n')
start = next(i for i, l in enumerate(txt) if l.startswith('.text.') and kern in l)
cur, ins, prev_annot = None, [], False
pat = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
for l in txt[start + 1:]:
    if (l.startswith('.text.') or l.startswith('.section')) and ins:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not prev_annot:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
        prev_annot = True
        continue
    prev_annot = False
    m = pat.match(l)
    if m:
        ins.append((m.group(2), cur))
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
n = min(len(ins), len(data))
agg = defaultdict(lambda: [0, 0])
for k in range(n):
    key = ins[k][1] or ('?', 0)
    agg[key][0] += int(data[k][iex] or 0)
    agg[key][1] += int(data[k][ismp] or 0)
tots = sum(v[1] for v in agg.values()) or 1
byfile = defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
print(f"{len(ins)} sass instr, {len(data)} csv rows")
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{f:24s} thread-instr/cell {v[0]*32/cells:8.1f}  stall samples {v[1]/tots*100:5.1f}%")
print('--- top lines by instructions')
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{f}:{l:<5d} instr/cell {v[0]*32/cells:7.1f}  samples {v[1]/tots*100:5.1f}%")
