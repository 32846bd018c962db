#!/bin/bash
# Round-end measurement pass on the GPU box (1 GPU): tests, smoke, both bench arms, ncu launch list, per-kernel ncu counts.
# Everything lands in gpurun_out/ (tag = $1).
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_bench_$tag.log 2>&1; echo "launch list rc=$?"
bash profiles/tools/ncu_counts.sh
