"""One forward + backward of the NatureCNN tower at training-chunk size (for ncu captures of the conv kernels)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baselines_b200.common import spaces
from baselines_b200.common.policies import build_policy
from baselines_b200.ppo2.model import Model

B = int(os.environ.get("PROF_B", 16384))
class E:
    observation_space = spaces.Box(0, 255, (84, 84, 4), np.uint8)
    action_space = spaces.Discrete(6)
    num_envs = 64
np.random.seed(0)
m = Model(policy=build_policy(E, "cnn"), ob_space=E.observation_space, ac_space=E.action_space, nbatch_act=64,
          nbatch_train=B, nsteps=4, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5, comm=False, train_chunk=B)
dev = m.device
obs = torch.randint(0, 256, (B, 84, 84, 4), dtype=torch.uint8, device=dev)
a = torch.randint(0, 6, (B,), device=dev)
v = torch.randn(B, device=dev)
r = v + torch.randn(B, device=dev)
n = torch.full((B,), 1.79, device=dev)
for _ in range(int(os.environ.get("PROF_ITERS", 2))):
    m.train_rollout(2.5e-4, 0.1, obs, a, r, v, n, torch.randperm(B, device=dev) if os.environ.get("PROF_GATHER", "1") == "1" else None)
torch.cuda.synchronize()
print("done")
