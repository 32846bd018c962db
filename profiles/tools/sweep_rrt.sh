set -e
for cfg in "128 6 0" "256 3 0" "384 2 0" "128 4 0"; do set -- $cfg
  python theta_rrt_b200/build.py -DTRRT_SPEC_THREADS=$1 -DTRRT_SPEC_BLOCKS_PER_SM=$2 -DTRRT_SPEC_LOCKSTEP=$3 > /dev/null 2>&1
  python bench.py --steps 3 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/sw_x.json 2>gpurun_out/sw.err
  python -c "
import json;d=json.loads(open('gpurun_out/sw_x.json').read().strip().splitlines()[-1]);print('threads',$1,'blocks/SM',$2,'lockstep',$3,'ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1))"
done
