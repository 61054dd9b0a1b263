#!/bin/bash
# quick iteration pass: conv kernel tests, then the whole GPU suite, then a short bench (no e2e / cpu legs)
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "uint8 or shift or xfold" > gpurun_out/it_ktests.log 2>&1; rc=$?; echo "ktests rc=$rc t=$((SECONDS-S))"; tail -3 gpurun_out/it_ktests.log
if [ $rc -ne 0 ]; then grep -n "^E " gpurun_out/it_ktests.log | head -20; exit 1; fi
S=$SECONDS
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/it_tests.log 2>&1; echo "tests rc=$? t=$((SECONDS-S))"; tail -3 gpurun_out/it_tests.log
S=$SECONDS
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-targets --no-others > gpurun_out/it_bench.json 2> gpurun_out/it_bench.err; echo "bench rc=$? t=$((SECONDS-S))"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/it_bench.json') if l.startswith('{"metric"')][0])
print('ms_per_step', d['ms_per_step'], d['clocks'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:14]:
    print('  %-32s %7.2f ms %5.0f'%(k, v['ms_per_step'], v['launches_per_step']))
PY
