profiles/tools/mb_peaks > gpurun_out/peaks.json; cat gpurun_out/peaks.json
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rrt or cfg5 or cfg3" > gpurun_out/t_rrt.log 2>&1; tail -3 gpurun_out/t_rrt.log
python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err; python -c "
import json;d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1]);print('ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), d['e2e']['ms_per_step'])"
python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu --valid-rows-d2h > gpurun_out/b2.json 2> gpurun_out/b2.err; python -c "
import json;d=json.loads(open('gpurun_out/b2.json').read().strip().splitlines()[-1]);print('vro: ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), d['e2e']['ms_per_step'], d['e2e']['serial_ms_per_step'])"
