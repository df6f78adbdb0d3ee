"""Condenses an ncu report (ncu -i X.ncu-rep --page raw --csv on stdin) to the metrics quoted in DESIGN.md / bench.py."""
import csv
import sys

KEEP = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
r = list(csv.reader(sys.stdin))
idx = [i for i, h in enumerate(r[0]) if h in KEEP]
w = csv.writer(sys.stdout)
for row in r:
    w.writerow([row[i] for i in idx])
