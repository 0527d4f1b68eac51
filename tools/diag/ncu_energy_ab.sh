# Issue / pipe / stall metrics of energy_kernel for one or two builds of the library (A/B of a tuning build made with
# `python -m graspit_b200.build -D... --out=...`).   usage (GPU box): tools/diag/ncu_energy_ab.sh [other_lib.so]
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct
i=0
for lib in "" "$@"; do
  [ -n "$lib" ] && export B2C_LIB=$lib
  ncu --metrics $M --clock-control none -k regex:energy_kernel -s 6 -c 2 --csv --log-file gpurun_out/ncu_energy_v$i.csv python bench.py --steps 4 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/ncu_energy_v$i.log 2>&1
  i=$((i+1))
done
