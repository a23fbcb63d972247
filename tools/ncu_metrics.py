"""key metrics of the first kernel in an ncu report: python tools/ncu_metrics.py REP"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:90])
    for k in hdr:
        if any(t in k for t in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "warps_active.avg.pct",
                                "registers_per_thread", "smsp__inst_executed.sum", "issue_active.avg.pct", "dram_throughput.avg.pct",
                                "bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "lts__t_sector_hit_rate.pct",
                                "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_xu", "inst_executed_pipe_alu.sum", "inst_executed_pipe_fma.sum", "inst_executed_pipe_lsu.sum")) or \
           ("issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if "stalled" in k and v < 0.05:
                continue
            print(f"  {k:95s} {d[k]}")
