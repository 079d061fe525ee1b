"""Summarise an .ncu-rep (raw page) into the few metrics we track.  Usage: ncu_summary.py <rep> [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:10]  # (the first eight captured launches)
idx = {h: i for i, h in enumerate(hdr)}
want = [
 ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
 ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
 ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
 ("lts__t_sectors_op_read.sum", "l2_rd_sectors"), ("lts__t_sectors_op_write.sum", "l2_wr_sectors"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
 ("smsp__inst_executed.sum", "warp_inst"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
 ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
 ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
 ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
 ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
 ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
 ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
]
lines = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
for key, name in want:
    if key not in idx:
        continue
    vals = []
    for r in data:
        v = r[idx[key]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.4g}"
        except ValueError:
            v = v[:40]
        vals.append(v)
    lines.append(f"| {name} | {units[idx[key]]} | " + " | ".join(vals) + " |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
