"""Extract a few metrics of every captured launch from an .ncu-rep into the small CSV kept under profiles/
(metric,unit,launch0,launch1,...).  Usage: python tools/ncu_extract.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
with open(sys.argv[2], "w") as f:
    f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))) + "\n")
    f.write("kernel,," + ",".join(r[kn].split("(")[0].replace(",", ";") for r in data) + "\n")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            f.write(",".join([w, units[i]] + [r[i].replace(",", "") for r in data]) + "\n")
print(open(sys.argv[2]).read())
