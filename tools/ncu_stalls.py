#!/usr/bin/env python
"""Stall summary of one kernel of an ncu report: overall stall-reason shares and the hottest instructions.
usage: tools/ncu_stalls.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
iS, iI, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iSm]) for r in data)
agg = {}
for r in data:
    for i in cols: agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print('total warp-inst %.1fM' % (sum(int(r[iI]) for r in data) / 1e6))
print('  '.join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
top = sorted(enumerate(data), key=lambda t: -int(t[1][iSm]))[:topn]
for n, r in sorted(top):
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in cols), reverse=True)[:2]
    print(f"{n:4d} exec={int(r[iI])/1e6:6.2f}M smp={100*int(r[iSm])/tot:4.1f}% {r[iS].strip()[:58]:58s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
