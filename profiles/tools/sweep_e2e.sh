for ch in 1 2 4 8 16; do
  python bench.py --steps 3 --warmup 3 --skip-secondary --skip-cpu --chunks $ch > gpurun_out/sw_e.json 2>gpurun_out/sw.err || tail -5 gpurun_out/sw.err
  python -c "
import json;d=json.loads(open('gpurun_out/sw_e.json').read().strip().splitlines()[-1]);print('chunks',$ch,'kernel ms',round(d['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'e2e Mexp/s',round(d['e2e']['value']/1e6,1))"
done
