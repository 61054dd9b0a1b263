"""Data-parallel plumbing: the reference's only multi-process strategy (collective table).

  MpiAdamOptimizer.compute_gradients  common/mpi_adam_optimizer.py:18-51  -> one all-reduce(mean) of the flat
                                                                           fp32 gradient buffer per minibatch
  sync_from_root                      common/mpi_util.py:15-26            -> broadcast of params + Adam slots

One process per GPU; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is the transport.
`comm` may be None (use the default process group if it is initialised), False (force single process) or a
torch.distributed ProcessGroup.
"""
import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, comm=None, rank_weight=1):
        self.group = None
        self.world = 1
        self.rank = 0
        self.rank_weight = float(rank_weight)
        if comm is False:
            return
        if dist.is_available() and dist.is_initialized():
            self.group = comm if (comm is not None and comm is not True) else dist.group.WORLD
            self.world = dist.get_world_size(self.group)
            self.rank = dist.get_rank(self.group)
        self.total_weight = None

    @property
    def active(self):
        return self.world > 1

    def _total_weight(self, device):
        if self.total_weight is None:
            t = torch.tensor([self.rank_weight], dtype=torch.float32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)      # mpi_adam_optimizer.py:26
            self.total_weight = float(t.item())
        return self.total_weight

    def average_gradients(self, store):
        """flat_grad * rank_weight -> Allreduce(SUM) -> / total_weight (mpi_adam_optimizer.py:21,39-40)."""
        if not self.active:
            return
        g = store.grads
        tw = self._total_weight(g.device)
        if self.rank_weight != 1.0:
            g.mul_(self.rank_weight)
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        g.mul_(1.0 / tw)

    def sync_from_root(self, store):
        if not self.active:
            return
        for buf in (store.params, store.m, store.v):
            dist.broadcast(buf, src=dist.get_global_rank(self.group, 0) if self.group is not dist.group.WORLD else 0,
                           group=self.group)

    def check_synced(self, store):
        """mpi_adam_optimizer.py:53-68: parameters must be identical on every rank."""
        if not self.active:
            return True
        s = store.params.double().sum().reshape(1)
        lo, hi = s.clone(), s.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool((lo == hi).item())
