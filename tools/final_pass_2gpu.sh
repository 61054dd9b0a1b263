#!/bin/bash
# N=1 and N=2 on the same 2-GPU box (run under `gpurun --gpus 2` from the repo root)
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-targets --no-others --no-e2e > gpurun_out/fp2_n1.json 2> gpurun_out/fp2_n1.err; echo "n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/fp2_n2.json 2> gpurun_out/fp2_n2.err; echo "n2 rc=$?"
timeout 300 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/fp2_dist_tests.log 2>&1; echo "dist tests rc=$?"; tail -2 gpurun_out/fp2_dist_tests.log
python - <<'PY'
import json
for f in ("fp2_n1", "fp2_n2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith('{"metric"')][0])
        print(f, d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"]), d.get("per_rank_kernel_ms"), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
