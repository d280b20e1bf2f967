#!/usr/bin/env python
"""Short CSV summary of `ncu --set full` reports:  python tools/ncu_summary.py out.csv a.ncu-rep [b.ncu-rep ...]
One row per captured launch with the metrics the roofline discussion uses (duration, DRAM bytes, pipe / issue
utilisation, occupancy, top warp-stall reasons)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'time_us'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
    ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('smsp__issue_active.avg.pct', 'issue_active_pct'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu_pct'),
    ('sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'uniform_pct'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
    ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
]
STALL = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'mio_throttle', 'lg_throttle', 'barrier',
          'not_selected', 'no_instruction', 'membar', 'sleeping']

out = csv.writer(open(sys.argv[1], 'w', newline=''))
out.writerow(['report', 'kernel'] + [k[1] for k in KEYS] + ['units(time,read,write)'] + ['stall_' + s for s in STALLS])
for rep in sys.argv[2:]:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        def g(k):
            return r[h.index(k)] if k in h else ''
        u = '/'.join(units[h.index(k)] if k in h else '' for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum'))
        out.writerow([rep.split('/')[-1], g('Kernel Name')[:80]] + [g(k[0]) for k in KEYS] + [u] + [g(STALL % s) for s in STALLS])
