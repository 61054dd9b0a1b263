"""MicrobatchedModel -- same constructor contract as baselines/ppo2/microbatched_model.py:5-33: the minibatch is
processed in microbatches of `microbatch_size` samples, advantages are normalised over the FULL minibatch (:43)
and the gradients of all microbatches are averaged into a single optimizer step (:70-75).  On the GPU this is the
ordinary chunked accumulation of `Model.train_rollout`."""
from .model import Model


class MicrobatchedModel(Model):
    def __init__(self, *, policy, ob_space, ac_space, nbatch_act, nbatch_train, nsteps, ent_coef, vf_coef,
                 max_grad_norm, mpi_rank_weight=1, comm=None, microbatch_size=None, **kw):
        self.nmicrobatches = nbatch_train // microbatch_size
        self.microbatch_size = microbatch_size
        assert nbatch_train % microbatch_size == 0, \
            'microbatch_size ({}) should divide nbatch_train ({}) evenly'.format(microbatch_size, nbatch_train)
        super().__init__(policy=policy, ob_space=ob_space, ac_space=ac_space, nbatch_act=nbatch_act,
                         nbatch_train=nbatch_train, nsteps=nsteps, ent_coef=ent_coef, vf_coef=vf_coef,
                         max_grad_norm=max_grad_norm, mpi_rank_weight=mpi_rank_weight, comm=comm,
                         microbatch_size=microbatch_size, **kw)
