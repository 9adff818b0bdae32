"""Print the interesting metrics of an `ncu --page raw --csv` export: python tools/ncu_show.py file.csv [more.csv]"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_xu.sum",
        "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled"]
for path in sys.argv[1:]:
    raw = list(csv.reader(open(path)))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    for r in rows:
        print("==", path, r[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS) or "stall" in h and "pct" in h or "warp_issue_stalled" in h and h.endswith("ratio"):
                print(f"  {h:90s} {r[i]:>16s} {units[i]}")
