#!/usr/bin/env python
"""Per-opcode stall-sample summary of the hot loop of one kernel in an .ncu-rep (read on the CPU box).
usage: python tools/ncu_source_hot.py prof.ncu-rep [window_start window_len]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
ex = [int(r[ix['Instructions Executed']]) for r in data]
big = collections.Counter(e for e in ex if e > 10000)
top = big.most_common(1)[0][0]
loop = [r for r in data if top * 0.9 <= int(r[ix['Instructions Executed']]) <= top * 1.1]
tot = sum(int(r[ix['# Samples']]) for r in data)
print("kernel:", rows[0][1][:100])
print("total samples", tot, "| loop instrs", len(loop), "| samples in loop", sum(int(r[ix['# Samples']]) for r in loop))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
byop, cnt, st = collections.Counter(), collections.Counter(), collections.Counter()
for r in loop:
    toks = r[ix['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    key = op.split('.')[0] + ('.WIDE' if 'WIDE' in op else '') + ('.X' if '.X' in op else '')
    byop[key] += int(r[ix['# Samples']]); cnt[key] += 1
    for c in stall_cols:
        st[c] += int(r[ix[c]])
print("opcode: count, samples, samples/instr")
for op, v in byop.most_common(14):
    print("  %-14s %4d %6d %6.1f" % (op, cnt[op], v, v / cnt[op]))
print("stalls in loop:", [(k[6:], v) for k, v in st.most_common(8)])
if len(sys.argv) > 3:
    a, n = int(sys.argv[2]), int(sys.argv[3])
    for r in loop[a:a + n]:
        print(r[ix['# Samples']].rjust(5), {c[6:]: r[ix[c]] for c in stall_cols if int(r[ix[c]]) > 3}, r[ix['Source']][:100])
