#!/usr/bin/env python
"""Summarise the source page of an ncu report: per-instruction execution counts bucketed, plus hot SASS lines.
usage: tools/ncu_sass.py report.ncu-rep kernel_regex [min_Mexec]"""
import csv, subprocess, sys, collections, io
rep, rx = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iI, iT, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iI]) for r in rows[2:]); smp = sum(int(r[iSm]) for r in rows[2:])
print(rows[0][1][:100]); print('total warp-inst %.1fM, samples %d' % (tot / 1e6, smp))
b = collections.OrderedDict()
for r in rows[2:]:
    k = round(int(r[iI]) / 1e6, 1)
    e = b.setdefault(k, [0, 0, 0, 0]); e[0] += 1; e[1] += int(r[iI]); e[2] += int(r[iSm]); e[3] += int(r[iT])
for k in sorted(b, reverse=True)[:14]:
    e = b[k]
    print(f"exec={k:7.1f}M n_instr={e[0]:4d} warp-inst={e[1]/1e6:8.1f}M ({100*e[1]/tot:4.1f}%) samples={100*e[2]/smp:4.1f}% lanes={e[3]/max(e[1],1):4.1f}")
if thr is not None:
    for n, r in enumerate(rows[2:]):
        I = int(r[iI])
        if I / 1e6 >= thr:
            print(f"{n:4d} {I/1e6:8.2f}M l={int(r[iT])/max(I,1):4.1f} s={r[iSm]:>6s} {r[iS].strip()[:80]}")
