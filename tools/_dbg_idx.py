import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baselines_b200 import ops
B = int(sys.argv[1]); POOL = int(sys.argv[2])
dev = "cuda"; f16 = dict(dtype=torch.float16, device=dev)
frames = torch.randint(0, 256, (POOL, 84, 84, 4), dtype=torch.uint8, device=dev)
N1 = 32
w1 = (torch.randn(N1, 256, device=dev) * 0.01).half(); b1 = torch.randn(N1, device=dev)
h1 = torch.zeros(B, 10, 10, 4 * N1, **f16)
omap1 = (2, 100 * 4 * N1, 10 * 4 * N1, 4 * N1, N1, 2)
def timed(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
perm = torch.randperm(POOL, device=dev)[:B].contiguous()
for name, idx in (("random", perm), ("sorted", torch.sort(perm).values.contiguous()), ("arange", torch.arange(B, device=dev))):
    u8 = (frames, idx, 84, 84, 4, 4)
    for m in (0, 6):
        os.environ["B200RL_CONV_DEBUG"] = str(m)
        ms = timed(lambda: ops.conv_shift_fwd(None, B, 21, 21, 64, w1, 256, N1, [0, 1, 21, 22], 20, 20, h1, omap1, bias=b1, act=ops.ACT_RELU, u8=u8))
        print(json.dumps({"idx": name, "mask": m, "ms": round(ms, 4), "B": B, "pool": POOL}), flush=True)
