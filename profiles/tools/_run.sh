timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rrt or cfg5 or cfg3" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json;d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]);print('ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), d['e2e']['ms_per_step'])"
python profiles/tools/phase_prof.py > gpurun_out/phase1.txt 2>&1; cat gpurun_out/phase1.txt
