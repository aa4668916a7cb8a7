"""Print selected metrics from an `ncu --page raw --csv` export (one column per captured launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct",
    "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct", "issue_stalled", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "dram__throughput", "achieved_occupancy", "sm__inst_executed_pipe"]
for i, h in enumerate(hdr):
    if any(p in h for p in pats) or h == "Kernel Name":
        print(f"{h:88s} {units[i]:12s}", [r[i] for r in data])
