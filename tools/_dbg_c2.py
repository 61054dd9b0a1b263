import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baselines_b200 import ops
B = int(sys.argv[1]); masks = [int(m) for m in sys.argv[2].split(",")]; reps = int(sys.argv[3])
dev = "cuda"; f16 = dict(dtype=torch.float16, device=dev)
x2 = (torch.randn(B, 100 * 128, device=dev) * 0.5).half()
w2 = (torch.randn(64, 4 * 128, device=dev) * 0.01).half()
b2 = torch.randn(64, device=dev)
h2 = torch.zeros(B, 81 * 64, **f16)
bits2 = torch.zeros(B * 81 * 64 // 16, dtype=torch.int16, device=dev)
for m in masks:
    os.environ["B200RL_CONV_DEBUG"] = str(m)
    for r in range(reps):
        ops.conv_shift_fwd(x2, B, 10, 10, 128, w2, 512, 64, [0, 1, 10, 11], 9, 9, h2, (0, 81 * 64, 9 * 64, 64, 0, 0),
                           bias=b2, act=ops.ACT_RELU, bits_out=bits2)
        torch.cuda.synchronize()
    print("mask", m, "ok", flush=True)
