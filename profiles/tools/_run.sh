mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final2.log 2>&1; tail -2 gpurun_out/pytest_gpu_final2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final2.log 2>&1; tail -1 gpurun_out/smoke_final2.log
timeout 900 python -m pytest tests -q -m gpu --trrt-so profiles/tools/_variants/checked.so > gpurun_out/checked_build_gpu_tests2.log 2>&1; tail -2 gpurun_out/checked_build_gpu_tests2.log
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
