"""2-GPU NCCL data parallel (needs 2 devices; skipped otherwise): semantics of the reference's MPI mode
(mpi_adam_optimizer.py:39-40 mean-allreduce BEFORE the clip, model.py:107; local per-rank advantage
normalisation, model.py:139; root broadcast model.py:131)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from baselines_b200.common import spaces
    from baselines_b200.common.policies import build_policy
    from baselines_b200.ppo2.model import Model

    class E:
        observation_space = spaces.Box(0, 255, (84, 84, 4), np.uint8)
        action_space = spaces.Discrete(6)
        num_envs = 16
    np.random.seed(100 + rank)                      # DIFFERENT init per rank: sync_from_root must fix it
    model = Model(policy=build_policy(E, "cnn"), ob_space=E.observation_space, ac_space=E.action_space, nbatch_act=16,
                  nbatch_train=data["M"], nsteps=4, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5)
    p0 = model.get_params()
    d = data["shards"][rank]
    st = model.train(2.5e-4, 0.1, d["obs"], d["returns"], None, d["actions"], d["values"], d["nlp"])
    q.put((rank, p0, model.get_params(), st, model.dist.check_synced(model.net.store)))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradient_mean_before_clip():
    import torch.multiprocessing as mp
    from oracle import nets
    M = 128
    rng = np.random.RandomState(0)
    shards = []
    for r in range(2):
        values = rng.randn(M).astype(np.float32)
        shards.append(dict(obs=rng.randint(0, 256, (M, 84, 84, 4)).astype(np.uint8),
                           actions=rng.randint(0, 6, M), values=values,
                           returns=(values + rng.randn(M) * (1 + r)).astype(np.float32),
                           nlp=np.full(M, np.log(6), np.float32)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dict(M=M, shards=shards), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, p0a, p1a, sta, syn_a), (_, p0b, p1b, stb, syn_b) = res
    for k in p0a:                                          # rank 1 adopted rank 0's initial parameters
        assert np.array_equal(p0a[k], p0b[k]), k
        assert np.array_equal(p1a[k], p1b[k]), k           # identical after the step on every rank
    assert syn_a and syn_b
    # oracle: per-rank grads (each with ITS OWN advantage normalisation), mean, then clip + Adam
    o = [nets.PPO2Oracle(p0a, "cnn", 0.01, 0.5, 0.5) for _ in range(3)]
    gs = []
    for r in range(2):
        d = shards[r]
        _, g = o[r].grads(0.1, d["obs"], d["returns"], d["actions"], d["values"], d["nlp"])
        gs.append(g)
    d = shards[0]
    o[2].train(2.5e-4, 0.1, d["obs"], d["returns"], None, d["actions"], d["values"], d["nlp"],
               grad_transform=lambda g: [(a + b) / 2 for a, b in zip(gs[0], gs[1])])
    po = o[2].params_np()
    err = max(float(np.abs(p1a[k] - po[k]).max()) for k in po)
    assert err < 3e-3, err
    assert not np.allclose(sta, stb)                       # loss statistics stay local (different shards)
