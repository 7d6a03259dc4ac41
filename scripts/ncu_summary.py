"""condense an `ncu --page raw --csv` export into one row per launch with the metrics the design notes quote:
usage: python scripts/ncu_summary.py <raw.csv> <out.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
I = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed_op_local_ld.sum", "local_ld"), ("smsp__inst_executed_op_local_st.sum", "local_st")]
cols = [(src, dst) for src, dst in want if src in I]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([dst + (" [%s]" % units[I[src]] if units[I[src]] else "") for src, dst in cols])
    for r in data:
        w.writerow([r[I[src]][:90] for src, dst in cols])
print(len(data), "launches ->", sys.argv[2])
