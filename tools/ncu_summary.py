"""Developer script: one block of key metrics per profiled launch of an .ncu-rep (ncu --set full), for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<what was profiled>" > profiles/rNN_<name>_full_summary.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
print("ncu --set full --clock-control none --import-source on ; %s (1 GPU, B200)" % (sys.argv[2] if len(sys.argv) > 2 else ""))
print("cold-cache (ncu flushes caches between replays), serialised launches; dram bytes are per launch\n")
ik = hdr.index("Kernel Name")
for n, r in enumerate(rows[2:]):
    print("[%d] %s" % (n, r[ik][:150]))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("    %-90s %14s %s" % (w, r[i], units[i]))
    print()
