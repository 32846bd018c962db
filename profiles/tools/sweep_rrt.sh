for cfg in "384 2 16" "768 1 16" "384 2 32" "256 3 16"; do set -- $cfg
  python theta_rrt_b200/build.py -DTRRT_SPEC_THREADS=$1 -DTRRT_SPEC_BLOCKS_PER_SM=$2 -DTRRT_TILE_PAIRS=$3 > /dev/null 2>&1 || { echo "build failed $1 $2"; continue; }
  python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/sw_x.json 2>gpurun_out/sw.err
  python -c "
import json;d=json.loads(open('gpurun_out/sw_x.json').read().strip().splitlines()[-1]);print('threads',$1,'blocks/SM',$2,'tile pairs',$3,'ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1))"
done
