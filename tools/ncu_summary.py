"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [pattern ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "registers_per_thread",
    "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_fma", "smsp__inst_executed.avg.per_cycle_active",
    "bank_conflicts", "wavefronts_mem_shared", "issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct",
    "sm__pipe_tensor", "dram__throughput.avg.pct", "occupancy", "local", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__inst_executed.sum", "smsp__inst_executed_op_shared", "l1tex__data_pipe_lsu_wavefronts.sum"]
for r in rows[2:]:
    print("==", r[4][:80], "grid", r[8], "block", r[7])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats):
            print(f"  {h} [{u}] = {v}")
