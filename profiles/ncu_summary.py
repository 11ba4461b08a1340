"""Summarise an ncu --page raw --csv export: python profiles/ncu_summary.py raw.csv [row indices]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
sel = [int(x) for x in sys.argv[2:]] or list(range(len(rows) - 2))
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum']
want += [n for n in h if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio')]
for k in sel:
    r = rows[k + 2]
    print('--- row', k, r[h.index('Kernel Name')][:50])
    for n in want:
        if n in h:
            v = r[h.index(n)]
            if n.startswith('smsp__average_warps_issue_stalled'):
                try:
                    if float(v.replace(',', '')) < 0.05:
                        continue
                except ValueError:
                    pass
            print('  %-95s %s %s' % (n, v, rows[1][h.index(n)]))
