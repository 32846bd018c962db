#!/bin/bash
# Per-launch instruction / byte counts of the hot kernels on their bench workloads (metrics pass, a few replays each).
# Output: gpurun_out/counts_<kernel>.csv ; profiles/tools/ncu_counts.py turns them into profiles/kernel_counts.json
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__thread_inst_executed_pipe_fp64_pred_on.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
mkdir -p gpurun_out
run() { # name, kernel regex, skip, workload args...
  name=$1; rx=$2; skip=$3; shift 3
  python profiles/prof_workload.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:$rx -s $skip -c 1 --csv --log-file gpurun_out/counts_$name.csv python profiles/prof_workload.py "$@" > gpurun_out/ncu_counts_$name.log 2>&1
  echo "$name rc=$?"
}
run rrt_kernel rrt_kernel_spec 1 rrt 4096 5001 32
run los_tiled_kernel los_tiled 1 los
run nearest_tile_kernel nearest_tile 1 nearest
run theta_kernel theta_kernel 1 theta 8192
