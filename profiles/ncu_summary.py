#!/usr/bin/env python
"""Summarises an ncu report (one row per captured kernel launch): duration, DRAM bytes, issue utilisation, occupancy and
the warp-stall breakdown (pc sampling), as CSV on stdout.

    python profiles/ncu_summary.py gpurun_out/X.ncu-rep > profiles/X_summary.csv
"""
import csv
import subprocess
import sys

KEEP = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_peak'),
    ('lts__t_sectors_op_atom.sum', 'l2_atom_sectors'),
    ('lts__t_sectors_op_red.sum', 'l2_red_sectors'),
    ('smsp__inst_executed.sum', 'warp_instructions'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('launch__registers_per_thread', 'registers'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__occupancy_limit_shared_mem', 'occ_limit_smem_blocks'),
    ('launch__occupancy_limit_registers', 'occ_limit_reg_blocks'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
]
STALLS = ['barrier', 'branch_resolving', 'dispatch_stall', 'drain', 'lg_throttle', 'long_scoreboard', 'math_pipe_throttle', 'membar',
          'mio_throttle', 'misc', 'no_instructions', 'not_selected', 'selected', 'short_scoreboard', 'sleeping', 'tex_throttle', 'wait']

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
w = csv.writer(sys.stdout)
w.writerow(['kernel'] + [n for _, n in KEEP] + ['stall_samples_total'] + ['stall_' + s + '_pct' for s in STALLS])
for r in data:
    name = r[col['Kernel Name']].split('(')[0]
    out = [name]
    for key, _ in KEEP:
        i = col.get(key)
        out.append('%s %s' % (r[i], units[i]) if i is not None and units[i] else (r[i] if i is not None else ''))
    samples = []
    for s in STALLS:
        i = col.get('smsp__pcsamp_warps_issue_stalled_' + s)
        samples.append(float(r[i]) if i is not None and r[i] else 0.0)
    tot = sum(samples) or 1.0
    out.append('%.0f' % sum(samples))
    out += ['%.1f' % (100 * x / tot) for x in samples]
    w.writerow(out)
