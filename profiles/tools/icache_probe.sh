# instruction-cache counters of the fused RRT kernel for several CTA shapes (experiment)
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for cfg in "512 1" "128 4"; do set -- $cfg
  python theta_rrt_b200/build.py -DTRRT_SPEC_THREADS=$1 -DTRRT_SPEC_BLOCKS_PER_SM=$2 -DTRRT_GRID_MIN_NODES=100000000 > /dev/null 2>&1
  python profiles/prof_workload.py rrt 4096 5001 32 > gpurun_out/prof_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:rrt_kernel -s 1 -c 1 --csv --log-file gpurun_out/icache_t$1.csv python profiles/prof_workload.py rrt 4096 5001 32 > /dev/null 2>&1
  echo "== threads $1 x $2"; grep -E '^"' gpurun_out/icache_t$1.csv | cut -d, -f13-15
done
