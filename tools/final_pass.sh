#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/fp_tests.log 2>&1; echo "tests rc=$? t=$((SECONDS-S))"; tail -2 gpurun_out/fp_tests.log
S=$SECONDS
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fp_smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-S))"; tail -1 gpurun_out/fp_smoke.log
S=$SECONDS
timeout 600 python bench.py > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err; echo "bench rc=$? t=$((SECONDS-S))"
S=$SECONDS
timeout 400 python bench.py --impl reference > gpurun_out/fp_bench_reference.json 2> gpurun_out/fp_bench_reference.err; echo "ref rc=$? t=$((SECONDS-S))"
S=$SECONDS
B200RL_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/fp_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-targets --no-others --no-profile > gpurun_out/fp_ncu_bench.log 2>&1; echo "launches rc=$? t=$((SECONDS-S))"
S=$SECONDS
PROF_B=16384 PROF_ITERS=1 B200RL_NO_GRAPHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_shift -o gpurun_out/fp_convs \
  python tools/profile_convs.py > gpurun_out/fp_ncu_convs.log 2>&1; echo "convs rc=$? t=$((SECONDS-S))"
ls -la gpurun_out/ | tail -12
