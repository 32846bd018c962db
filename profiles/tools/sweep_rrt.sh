for cfg in "320 2" "352 2" "288 2" "384 2"; do set -- $cfg
  python theta_rrt_b200/build.py -DTRRT_SPEC_THREADS=$1 -DTRRT_SPEC_BLOCKS_PER_SM=$2 -Xptxas -v 2>&1 | grep -A2 "rrt_kernel_specILi32" | grep -E "spill|registers" | tr '\n' ' '; echo
  python bench.py --steps 5 --warmup 3 --skip-secondary --skip-cpu > gpurun_out/sw_x.json 2>gpurun_out/sw.err
  python -c "
import json;d=json.loads(open('gpurun_out/sw_x.json').read().strip().splitlines()[-1]);print('threads',$1,'blocks/SM',$2,'ms',round(d['ms_per_step'],2),'Mexp/s',round(d['value']/1e6,1))"
done
