"""Times the per-minibatch gradient all-reduce (6.75 MB fp32, mpi_adam_optimizer.py:39) by itself and prints which NCCL
transport the box gives us.    torchrun --nproc-per-node N tools/nccl_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
g = torch.randn(1687720, device="cuda")
for _ in range(20):
    dist.all_reduce(g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_reduce(g)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    peer = torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None
    print({"allreduce_6.75MB_us": 1e3 * e0.elapsed_time(e1) / 200, "world": dist.get_world_size(), "p2p_0_1": peer})
dist.destroy_process_group()
