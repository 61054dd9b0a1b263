"""Role isolation of the shift-GEMM conv kernels (B200RL_CONV_DEBUG masks: 1 producers move no data, 2 no MMAs,
4 no epilogue / bias-sum work): which warp role bounds a tile?  Results of masked runs are garbage; only times matter.

    python tools/conv_roles.py [B]        -> one JSON line per (kernel, mask)
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baselines_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = "cuda"
torch.manual_seed(0)
f16 = dict(dtype=torch.float16, device=dev)


def timed(fn, iters=5):
    fn(); fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))


def run(name, fn, masks):
    base = None
    for m in masks:
        os.environ["B200RL_CONV_DEBUG"] = str(m)
        ms = timed(fn)
        base = base or ms
        print(json.dumps({"kernel": name, "mask": m, "ms": round(ms, 4), "vs_full": round(ms / base, 3), "B": B}), flush=True)
    os.environ["B200RL_CONV_DEBUG"] = "0"


# conv1 (uint8-fed): 84x84x4 frames, s2d 4 -> 21x21x64 grid, 2x2 taps, 32 channels, output s2d 2 for conv2
H = W = 84
N1 = 32
frames = torch.randint(0, 256, (B, H, W, 4), dtype=torch.uint8, device=dev)
idx = torch.randperm(B, device=dev)
u8 = (frames, idx, H, W, 4, 4)
w1 = (torch.randn(N1, 256, device=dev) * 0.01).half()
b1 = torch.randn(N1, device=dev)
h1 = torch.zeros(B, 10, 10, 4 * N1, **f16)
bits1 = torch.zeros(B * 100 * 4 * N1 // 16, dtype=torch.int16, device=dev)
omap1 = (2, 100 * 4 * N1, 10 * 4 * N1, 4 * N1, N1, 2)
run("c1.fwd(u8)", lambda: ops.conv_shift_fwd(None, B, 21, 21, 64, w1, 256, N1, [0, 1, 21, 22], 20, 20, h1, omap1, bias=b1,
                                             act=ops.ACT_RELU, u8=u8, bits_out=bits1), [0, 1, 2, 4, 3, 5, 6, 7])
dz1 = torch.zeros(B, 21, 21, N1, **f16)
dz1[:, :20, :20] = (torch.randn(B, 20, 20, N1, device=dev) * 0.5).half()
G1 = torch.zeros(256, N1, dtype=torch.float32, device=dev)
gb1 = torch.zeros(N1, dtype=torch.float32, device=dev)
run("c1.wgrad(u8,kx2)", lambda: ops.conv_shift_wgrad(None, B * 441, 64, dz1, N1, [0, 21], G1, N1, alpha=1.0 / 255, gbias=gb1,
                                                     alpha_b=1.0, u8=u8, kx=2), [0, 1, 2, 4, 3, 5, 6, 7])
del frames, dz1
# conv2 (TMA-fed): 10x10x128 grid, 2x2 taps, 64 channels
x2 = (torch.randn(B, 100 * 128, device=dev) * 0.5).half()
w2 = (torch.randn(64, 4 * 128, device=dev) * 0.01).half()
b2 = torch.randn(64, device=dev)
h2 = torch.zeros(B, 81 * 64, **f16)
bits2 = torch.zeros(B * 81 * 64 // 16, dtype=torch.int16, device=dev)
run("c2.fwd", lambda: ops.conv_shift_fwd(x2, B, 10, 10, 128, w2, 512, 64, [0, 1, 10, 11], 9, 9, h2, (0, 81 * 64, 9 * 64, 64, 0, 0),
                                         bias=b2, act=ops.ACT_RELU, bits_out=bits2), [0, 2, 4, 6])
# conv3: 9x9x64 grid, 3x3 taps, 64 channels
w3 = (torch.randn(64, 9 * 64, device=dev) * 0.01).half()
h3 = torch.zeros(B, 49 * 64, **f16)
bits3 = torch.zeros(B * 49 * 64 // 16, dtype=torch.int16, device=dev)
run("c3.fwd", lambda: ops.conv_shift_fwd(h2, B, 9, 9, 64, w3, 576, 64, [a * 9 + b for a in range(3) for b in range(3)], 7, 7, h3,
                                         (0, 49 * 64, 7 * 64, 64, 0, 0), bias=b2, act=ops.ACT_RELU, bits_out=bits3), [0, 2, 4, 6])
