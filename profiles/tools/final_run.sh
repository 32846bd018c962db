#!/bin/bash
# Round-end measurement pass on the GPU box: tests, smoke, both bench arms, ncu launch list, ncu captures of the
# fused RRT kernel and of the strip LOS kernel (skipped with a second argument "nocapture").  Everything lands in
# gpurun_out/ (tag = $1).
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_$tag.log 2>&1; echo "launch list rc=$?"
[ "$2" = "nocapture" ] && exit 0
ncu --set full --import-source on --clock-control none -k regex:rrt_kernel_spec -c 1 -s 1 -o gpurun_out/rrt_r1_$tag -f \
    python profiles/prof_workload.py rrt 4096 5001 32 > gpurun_out/ncu_rrt_$tag.log 2>&1; echo "ncu rrt rc=$?"
ncu --set full --import-source on --clock-control none -k regex:los_tiled -c 1 -s 1 -o gpurun_out/los_r1_$tag -f \
    python profiles/prof_workload.py los > gpurun_out/ncu_los_$tag.log 2>&1; echo "ncu los rc=$?"
