"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns the rooflines use (one row per captured launch)."""
import csv, sys
KEEP = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active"]
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
head = rows[0]
idx = [head.index(k) for k in KEEP if k in head]
w = csv.writer(sys.stdout)
for r in rows:
    if len(r) >= len(head):
        w.writerow([r[i] for i in idx])
