#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the per-kernel table kept under profiles/.
usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep > profiles/X_summary.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per instruction"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu % (int / compare / select / min-max)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma % (fp32 add / fma, imad)"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe fmaheavy %"),
    ("sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "pipe fmalite %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "pipe uniform %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu data pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle / issue"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__occupancy_limit_registers", "occupancy limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit smem (blocks)"),
    ("sm__cycles_elapsed.max", "cycles elapsed"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# ncu --set full summary of `%s`\n" % rep.split("/")[-1])
    print("Per-launch values (ncu replays: cold cache, serialised; compare shares, not absolutes).\n")
    for r in rows[2:]:
        print("## %s  (launch id %s)\n" % (r[ki].split("(")[0], r[0]))
        print("| metric | value | unit |\n|---|---|---|")
        for name, label in WANT:
            if name in hdr:
                i = hdr.index(name)
                print("| %s (`%s`) | %s | %s |" % (label, name, r[i], units[i]))
        print()


if __name__ == "__main__":
    main()
