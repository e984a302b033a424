#!/usr/bin/env python
"""Top CUDA source lines of an ncu report by executed warp instructions and stall samples.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv ; python profiles/ncu_lines.py X.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, fname, out = None, '', []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
        d = dict(zip(hdr[4:], r[4:]))
        out.append((fname, int(r[0]), r[1], d))
tot_i = sum(float(o[3]['Instructions Executed'] or 0) for o in out)
tot_s = sum(float(o[3]['# Samples'] or 0) for o in out)
print('total warp instructions %.0f, samples %.0f' % (tot_i, tot_s))
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
print('--- by instructions')
for f, ln, src, d in sorted(out, key=lambda o: -float(o[3]['Instructions Executed'] or 0))[:top_n]:
    print('%6.2f%% %5s smp  %s:%d  %s' % (100 * float(d['Instructions Executed'] or 0) / max(tot_i, 1), d['# Samples'], f, ln, src.strip()[:110]))
print('--- by samples')
for f, ln, src, d in sorted(out, key=lambda o: -float(o[3]['# Samples'] or 0))[:top_n]:
    st = sorted(((float(d[k] or 0), k[6:]) for k in stalls), reverse=True)[:3]
    print('%6.2f%% %s:%d  %s   [%s]' % (100 * float(d['# Samples'] or 0) / max(tot_s, 1), f, ln, src.strip()[:90],
                                       ', '.join('%s %.0f' % (k, v) for v, k in st if v)))
