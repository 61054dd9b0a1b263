"""Experiment: may a K-major SWIZZLE_128B UMMA descriptor start `shift` rows (128 B each) into a tile, and does it
need the descriptor's base_offset field?  C[i] should equal A[i+shift] @ B^T for rows with (i % 128) < 128-shift."""
import os, subprocess, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(shift, bo):
    os.environ["B200RL_EXP_SHIFT"], os.environ["B200RL_EXP_BO"] = str(shift), str(bo)
    from baselines_b200 import ops
    torch.manual_seed(0)
    M, N, K = 512, 64, 192
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.5).half()
    C = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A, B, C, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, mode=ops.MODE_F32_STORE)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    rows = [i for i in range(M) if (i % 128) < 128 - shift]
    want = ref[[i + shift for i in rows]]
    got = C[rows]
    return float((got - want).abs().max()), float((got - ref[rows]).abs().max())

if __name__ == "__main__":
    out = {}
    for shift in (0, 1, 2, 3, 5, 8, 9, 11):
        for bo in sorted({0, shift % 8}):
            out[f"shift={shift},bo={bo}"] = run(shift, bo)
    print(json.dumps(out, indent=0))
