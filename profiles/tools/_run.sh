timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q --trrt-so profiles/tools/_variants/checked.so > gpurun_out/checked_build_gpu_tests.log 2>&1; tail -1 gpurun_out/checked_build_gpu_tests.log
python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json;d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]);print('ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), d['e2e']['ms_per_step'])"
