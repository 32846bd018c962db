set -e
for per in 2 4 1; do
  python theta_rrt_b200/build.py -DTRRT_SPEC_BARRIER_PERIOD=$per > /dev/null 2>&1
  python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/sw_x.json 2>gpurun_out/sw.err
  python -c "
import json;d=json.loads(open('gpurun_out/sw_x.json').read().strip().splitlines()[-1]);print('barrier period',$per,'ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1))"
done
