#!/usr/bin/env python
"""Group the SASS of one kernel of an ncu report into runs of equal execution count and print, per run, the
instruction total, stall samples and shared-memory wavefronts: a cheap 'which loop costs what' view.
usage: tools/ncu_regions.py report.ncu-rep kernel_regex [min_Minst]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; data = rows[2:]
iS, iI, iSm, iT = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
iW = hdr.index('L1 Wavefronts Shared')
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iSm]) for r in data)
runs = []
for n, r in enumerate(data):
    e = int(r[iI])
    if runs and abs(runs[-1]['e'] - e) <= 0.02 * max(e, 1):
        k = runs[-1]
    else:
        k = dict(e=e, n0=n, cnt=0, inst=0, smp=0, thr=0, wf=0, first=r[iS].strip()[:40]); runs.append(k)
    k['cnt'] += 1; k['inst'] += e; k['smp'] += int(r[iSm]); k['thr'] += int(r[iT]); k['wf'] += int(r[iW] or 0); k['n1'] = n
print(f"total {tot/1e6:.1f}M warp-inst, {tots} samples")
for k in runs:
    if k['inst'] / 1e6 >= thr:
        print(f"sass {k['n0']:5d}-{k['n1']:5d} exec={k['e']/1e6:7.2f}M x{k['cnt']:4d} = {k['inst']/1e6:7.1f}M ({100*k['inst']/tot:4.1f}%) smp={100*k['smp']/tots:4.1f}% lanes={k['thr']/max(k['inst'],1):4.1f} smem_wf={k['wf']/1e6:6.1f}M  {k['first']}")
