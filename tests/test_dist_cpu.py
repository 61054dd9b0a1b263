"""CPU, world_size 2, gloo: the data-parallel plumbing of common/dist_util.py (the reference's
MpiAdamOptimizer mean-allreduce, mpi_adam_optimizer.py:21,39-40, and sync_from_root, mpi_util.py:15-26)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Store:
    def __init__(self, rank, n=1000):
        g = torch.Generator().manual_seed(rank)
        self.params = torch.randn(n, generator=g)
        self.m = torch.randn(n, generator=g)
        self.v = torch.rand(n, generator=g)
        self.grads = torch.randn(n, generator=g)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, weights, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from baselines_b200.common.dist_util import DataParallel
    st = _Store(rank)
    g_local = st.grads.clone()
    dp = DataParallel(None, rank_weight=weights[rank])
    assert dp.active and dp.world == world and dp.rank == rank
    synced_before = dp.check_synced(st)
    dp.sync_from_root(st)
    synced_after = dp.check_synced(st)
    dp.average_gradients(st)
    q.put((rank, g_local.numpy(), st.grads.numpy(), st.params.numpy(), st.m.numpy(), synced_before, synced_after))
    dist.destroy_process_group()


def _run(weights):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, weights, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_gradient_mean_and_root_sync_world2():
    (r0, g0, a0, p0, m0, sb0, sa0), (r1, g1, a1, p1, m1, sb1, sa1) = _run([1.0, 1.0])
    assert np.allclose(a0, (g0 + g1) / 2, atol=1e-6) and np.array_equal(a0, a1)
    root = _Store(0)
    assert np.array_equal(p0, root.params.numpy()) and np.array_equal(p1, root.params.numpy())
    assert np.array_equal(m1, root.m.numpy())
    assert not sb1 and sa0 and sa1


def test_rank_weighted_mean_world2():
    """mpi_rank_weight (mpi_adam_optimizer.py:21,26,40): sum(w_r g_r) / sum(w_r)."""
    (_, g0, a0, *_), (_, g1, a1, *_) = _run([1.0, 3.0])
    assert np.allclose(a0, (1.0 * g0 + 3.0 * g1) / 4.0, atol=1e-6) and np.array_equal(a0, a1)
