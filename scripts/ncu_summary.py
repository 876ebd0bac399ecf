"""ncu report -> small CSV (one row per captured launch) with the metrics the roofline / DESIGN cite.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv [max_rows]"""
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_global_red.sum",
]
rep, out = sys.argv[1], sys.argv[2]
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
idx = [h.index(k) for k in KEEP if k in h]
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([h[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:2 + limit]:
        w.writerow([r[i] for i in idx])
print(out, len(rows) - 2, "launches in report")
