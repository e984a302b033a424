"""Summarise an ncu report: scratch/ncu_top.py <src csv (--page source --print-source sass,cuda)> <raw csv> [n] [kernel substring]"""
import csv, sys
src, raw = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
want = sys.argv[4] if len(sys.argv) > 4 else ''
rows = list(csv.reader(open(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if want not in r[idx['Kernel Name']]: continue
    print('==', r[idx['Kernel Name']][:60])
    for w in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
              'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
              'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct']:
        if w in idx: print('  ', w, r[idx[w]])
rows = list(csv.reader(open(src)))
cur = None; fn = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) == 2: continue
    if r[0] == 'Line No': hdr = r; ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples'); continue
    if r[2] == '-' and want in (fn or ''):
        try:
            k = (cur, int(r[0]), r[1].strip()[:100])
            a = agg.setdefault(k, [0, 0]); a[0] += int(r[ie] or 0); a[1] += int(r[isamp] or 0)
        except ValueError: pass
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print('total inst', tot, 'samples', ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print('%5.1f%% inst %5.1f%% smp  %s:%s  %s' % (100 * a[0] / max(tot, 1), 100 * a[1] / max(ts, 1), k[0], k[1], k[2]))
