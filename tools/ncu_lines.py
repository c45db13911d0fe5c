"""Summarise an ncu report per CUDA source line: instructions executed + stall samples (needs -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, agg, done = None, None, {}, set()
inst_first = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        key = (fname, r[1])
        if inst_first is None:
            inst_first = r[1]
        skip = (fname, r[1]) in done
        done.add((fname, r[1]))
        continue
    if r[0] == 'Line No':
        hdr = r
        ie, sm = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or skip or len(r) <= ie or r[2] != '-':
        continue       # only the per-line summary rows (Address == '-')
    try:
        n, s = int(r[ie]), int(r[sm])
    except ValueError:
        continue
    k = (fname, int(r[0]), r[1].strip()[:100])
    a = agg.setdefault(k, [0, 0])
    a[0] += n; a[1] += s
T = sum(a[0] for a in agg.values()); S = sum(a[1] for a in agg.values())
print('total instructions %d, samples %d' % (T, S))
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * n / T, 100.0 * s / max(S, 1), f, ln, src))
