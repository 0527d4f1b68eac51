#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the per-kernel metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_<tag>.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "uniform branch targets"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem), blocks"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sass__inst_executed_local_loads", "local (spill) loads"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per instruction issued (cycles)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (i-cache)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L1/L2/DRAM)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem/XU)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of `{rep.split('/')[-1]}`\n\n")
        f.write("Per launch (one captured launch per kernel; cold caches, serialised -- compare shares, not absolutes).\n"
                "Raw report is scratch (gpurun_out/); this file is the committed evidence.\n\n")
        for r in data:
            f.write(f"## {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n\n| metric | value |\n|---|---|\n")
            for k, label in KEYS:
                if k in col and r[col[k]] not in ("", "nan", "-nan"):
                    f.write(f"| {label} (`{k}`) | {r[col[k]]} {units[col[k]]} |\n")
            f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
