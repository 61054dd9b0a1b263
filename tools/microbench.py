#!/usr/bin/env python
"""Stand-alone timings of the two north-star kernels (GAE scan, NatureCNN fc1 GEMM) with CUDA events on the
launching stream, >=3 warm-ups, and an L2 flush (write of a 512 MB buffer) between timed iterations.

    python tools/microbench.py            # prints one JSON object
Used by bench.py (extra `targets` key) and under ncu for profiles/ (tools/microbench.py --only gae|fc1)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baselines_b200 import ops  # noqa: E402


def _time(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def gae_case(T, N, variant, flush):
    g = torch.Generator(device="cuda").manual_seed(0)
    rew = torch.randn(T, N, device="cuda", generator=g)
    val = torch.randn(T, N, device="cuda", generator=g)
    done = (torch.rand(T, N, device="cuda", generator=g) < 0.01).to(torch.uint8)
    lv = torch.randn(N, device="cuda", generator=g)
    ld = torch.zeros(N, dtype=torch.uint8, device="cuda")
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    med, best = _time(lambda: ops.gae_scan(rew, val, done, lv, ld, adv, ret, 0.99, 0.95, variant), flush=flush)
    nbytes = 17.0 * T * N + 5.0 * N
    return {"T": T, "N": N, "variant": variant, "ms": med, "ms_best": best, "bytes": nbytes,
            "gbs": nbytes / (med * 1e-3) / 1e9}


def fc1_case(M, kind, flush, K=3136, N=512):
    A = (torch.randn(M, K, device="cuda") * 0.1).half()
    W = (torch.randn(N, K, device="cuda") * 0.1).half()
    if kind == "fwd":
        C = torch.empty(M, N, dtype=torch.float16, device="cuda")
        bias = torch.zeros(N, device="cuda")
        fn = lambda: ops.gemm(A, W, C, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, mode=ops.MODE_F16_ACT,
                              act=ops.ACT_RELU)
    elif kind == "dgrad":
        dz = (torch.randn(M, N, device="cuda") * 0.1).half()
        Wb = (torch.randn(K, N, device="cuda") * 0.1).half()
        out = torch.empty(M, K, dtype=torch.float16, device="cuda")
        fn = lambda: ops.gemm(dz, Wb, out, M=M, N=K, K=N, lda=N, ldb=N, ldc=K, saved=A, ld_saved=K,
                              mode=ops.MODE_F16_DACT, act=ops.ACT_RELU)
    else:
        dz = (torch.randn(M, N, device="cuda") * 0.1).half()
        gw = torch.zeros(K, N, device="cuda")
        tiles = -(-K // 128) * -(-N // 128)
        split = max(1, min((M // 64) // 2, -(-296 // tiles)))
        fn = lambda: ops.gemm(A, dz, gw, M=K, N=N, K=M, lda=K, ldb=N, ldc=N, mn_major=True,
                              mode=ops.MODE_F32_ATOMIC, alpha=1.0, split_k=split)
    med, best = _time(fn, flush=flush)
    flops = 2.0 * M * N * K
    return {"M": M, "K": K, "N": N, "kind": kind, "ms": med, "ms_best": best, "flops": flops,
            "tflops": flops / (med * 1e-3) / 1e12}


def run(only=None, quick=False):
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")      # 512 MB > 126 MB L2
    out = {"l2_flush": "512 MB write between iterations"}
    if only in (None, "gae"):
        out["gae"] = [gae_case(128, 4096, -1, flush), gae_case(512, 16384, 1, flush), gae_case(512, 16384, 0, flush)]
    if only in (None, "fc1"):
        Ms = [8192, 131072] if not quick else [131072]
        out["fc1"] = [fc1_case(M, k, flush) for M in Ms for k in ("fwd", "dgrad", "wgrad")]
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--quick", action="store_true", help="fc1 at M=131072 only")
    a = ap.parse_args()
    print(json.dumps(run(a.only, a.quick)))
