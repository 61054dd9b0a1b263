#!/bin/bash
for v in 0 1; do
  B200RL_XFOLD_FWD=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-targets --no-others > gpurun_out/bp_$v.json 2> gpurun_out/bp_$v.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bp_$v.json') if l.startswith('{"metric"')][0])
print('XFOLD_FWD=$v ms_per_step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:12]:
    print('  %-32s %7.2f ms %5.0f'%(k, v['ms_per_step'], v['launches_per_step']))
PY
done
