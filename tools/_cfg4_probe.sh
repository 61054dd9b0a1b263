#!/bin/bash
B200RL_NO_GRAPHS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 400 --csv --log-file gpurun_out/cfg4_launches.csv \
  python bench.py --config cfg4 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-targets --no-others --no-profile > gpurun_out/cfg4_ncu.log 2>&1; echo "launches rc=$?"
tail -3 gpurun_out/cfg4_ncu.log | cut -c1-300
