#!/bin/bash
B200RL_NO_GRAPHS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/cfg3_launches.csv \
  python bench.py --config cfg3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-targets --no-others --no-profile > gpurun_out/cfg3_ncu.log 2>&1; echo "launches rc=$?"
B200RL_NO_GRAPHS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gauss_loss -c 1 -o gpurun_out/cfg3_gauss \
  python bench.py --config cfg3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-targets --no-others --no-profile > gpurun_out/cfg3_ncu2.log 2>&1; echo "gauss rc=$?"
