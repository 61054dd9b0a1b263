"""PPO2 learner object on B200 kernels -- drop-in for the reference's baselines/ppo2/model.py Model.

Same constructor keywords (model.py:27-28) and the same duck-typed protocol the reference's Runner / learn /
run.py consume: step, value, train, initial_state, loss_names, save, load.  Added device
entry points (`step_device`, `value_device`, `train_rollout`) let this repo's Runner / learn keep the
rollout resident in HBM instead of round-tripping numpy.
"""
import math
import os

import numpy as np
import torch

from .. import graphs, ops
from ..common.policies import PolicyNet
from ..common import dist_util


class Model(object):
    def __init__(self, *, policy, ob_space, ac_space, nbatch_act, nbatch_train, nsteps, ent_coef, vf_coef,
                 max_grad_norm, mpi_rank_weight=1, comm=None, microbatch_size=None, device=None,
                 train_chunk=None):
        if not torch.cuda.is_available():
            raise RuntimeError("baselines_b200.ppo2.Model needs a CUDA device: the learner hot path is "
                               "hand-written sm_100a CUDA and has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.ent_coef, self.vf_coef, self.max_grad_norm = float(ent_coef), float(vf_coef), max_grad_norm
        self.nbatch_act, self.nbatch_train, self.nsteps = nbatch_act, nbatch_train, nsteps
        chunk = int(train_chunk or os.environ.get("B200RL_TRAIN_CHUNK", 131072))
        if microbatch_size is not None:                       # microbatched_model.py:5: same maths, smaller launches
            chunk = int(microbatch_size)
        self.chunk = max(1, min(chunk, max(1, nbatch_train)))
        cap = max(nbatch_act, self.chunk)
        with torch.cuda.device(self.device):
            # ortho_init consumes the GLOBAL numpy stream seeded by set_global_seeds (ppo2.py:80), like the reference
            self.net = PolicyNet(policy, cap, self.device, rng=np.random)
            self.opt = _make_optimizer(self.net.store, max_grad_norm)
            self._act_a = torch.zeros((nbatch_act,) if self.net.discrete else (nbatch_act, self.net.nout),
                                      dtype=torch.int64 if self.net.discrete else torch.float32, device=self.device)
            self._act_v = torch.zeros(nbatch_act, dtype=torch.float32, device=self.device)
            self._act_nlp = torch.zeros(nbatch_act, dtype=torch.float32, device=self.device)
        self.loss_names = ['policy_loss', 'value_loss', 'policy_entropy', 'approxkl', 'clipfrac']   # model.py:115
        self.initial_state = None
        self.act_model = self.train_model = self
        self._rng_seed = int(np.random.randint(0, 2 ** 31 - 1))
        self.graphs = graphs.GraphCache()
        self._mb_idx = None                       # fixed home of the current minibatch's indices (graph replays)
        self._stats_out = torch.zeros(5, dtype=torch.float64, device=self.device)
        self.comm = comm
        self.dist = dist_util.DataParallel(comm, mpi_rank_weight)
        self._train_calls = 0
        self.dist.sync_from_root(self.net.store)             # model.py:131 sync_from_root
        self.net.refresh()

    # ------------------------------------------------------------------------------------ act path
    def step_device(self, obs_dev, actions, values, neglogp, noise=None, persistent=False):
        """PolicyWithValue.step (policies.py:77-96) on device tensors; obs_dev as produced by net.encode_obs.
        persistent=True: the caller passes the same buffers on every call (the Runner's rollout slots), so the launch
        sequence is captured once per slot and replayed."""
        B = obs_dev.shape[0]
        if persistent and noise is None:
            key = ("act", obs_dev.data_ptr(), actions.data_ptr(), values.data_ptr(), neglogp.data_ptr(), B)
            self.graphs.run(key, lambda: self.net.act(obs_dev, B, actions, values, neglogp, seed=self._rng_seed))
        else:
            self.net.act(obs_dev, B, actions, values, neglogp, noise=noise, seed=self._rng_seed)

    def value_device(self, obs_dev, values, persistent=False):
        B = obs_dev.shape[0]

        def body():
            self.net.forward(obs_dev, B)
            values.copy_(self.net.v_out[:B, 0] if self.net.v_out.dim() == 2 else self.net.v_out[:B])
        if persistent:
            self.graphs.run(("value", obs_dev.data_ptr(), values.data_ptr(), B), body)
        else:
            body()

    def step(self, observation, S=None, M=None, noise=None, **_):
        """numpy in / numpy out, like the reference: (actions, values, states=None, neglogpacs)."""
        with torch.cuda.device(self.device):
            x = self.net.encode_obs(np.asarray(observation))
            B = x.shape[0]
            a, v, n = self._bufs(B)
            nz = None if noise is None else torch.as_tensor(np.ascontiguousarray(noise), dtype=torch.float32).to(self.device)
            self.step_device(x, a, v, n, noise=nz)
            return a.cpu().numpy(), v.cpu().numpy(), None, n.cpu().numpy()

    def value(self, ob, *args, **kwargs):
        with torch.cuda.device(self.device):
            x = self.net.encode_obs(np.asarray(ob))
            B = x.shape[0]
            _, v, _ = self._bufs(B)
            self.value_device(x, v)
            return v.cpu().numpy()

    def _bufs(self, B):
        if B <= self._act_v.shape[0]:
            return self._act_a[:B], self._act_v[:B], self._act_nlp[:B]
        if B > self.net.cap:
            raise ValueError(f"batch {B} exceeds the workspace capacity {self.net.cap}")
        dev = self.device
        a = torch.zeros((B,) if self.net.discrete else (B, self.net.nout),
                        dtype=torch.int64 if self.net.discrete else torch.float32, device=dev)
        return a, torch.zeros(B, device=dev), torch.zeros(B, device=dev)

    # ------------------------------------------------------------------------------------ train path
    def train_rollout(self, lr, cliprange, obs, actions, returns, values, neglogpacs, src_idx):
        """One minibatch of ppo2/model.py:133-158 on device-resident rollout arrays.

        obs/actions/returns/values/neglogpacs: flat device buffers in buffer order; src_idx: int64 device
        tensor of the M buffer offsets forming this minibatch (the shuffled `mbinds` of ppo2.py:164 mapped to
        buffer order), or None for "all rows in order".  Returns a device float64[5] of the loss statistics."""
        net, store = self.net, self.net.store
        M = int(src_idx.numel()) if src_idx is not None else int(returns.numel())
        with torch.cuda.device(self.device):
            # scalars that change from call to call go to device memory first; everything after that is a fixed launch
            # sequence for a given (rollout buffers, M), captured once and replayed (graphs.py)
            ops.set_scalars(net.clip_dev, cliprange)
            self.opt.begin_step(lr)
            idx_all = None
            if src_idx is not None:
                if self._mb_idx is None or self._mb_idx.numel() < M:
                    self._mb_idx = torch.empty(max(M, self.nbatch_train), dtype=torch.int64, device=self.device)
                self._mb_idx[:M].copy_(src_idx)
                idx_all = self._mb_idx
            inv_M = 1.0 / M

            def grads():
                store.grads.zero_()
                net.stats.zero_()
                ops.adv_stats(returns, values, None if idx_all is None else idx_all[:M], M, net.adv_st)  # model.py:139
                for s in range(0, M, self.chunk):
                    B = min(self.chunk, M - s)
                    if idx_all is not None:
                        net.loss_backward(obs, B, idx_all[s:s + B], actions, returns, values, neglogpacs, None,
                                          self.ent_coef, self.vf_coef, inv_M)
                    else:
                        sl = slice(s, s + B)
                        net.loss_backward(obs[sl], B, None, actions[sl], returns[sl], values[sl], neglogpacs[sl],
                                          None, self.ent_coef, self.vf_coef, inv_M)
                net.freeze_identity()

            def update():
                self.opt.apply()                                         # model.py:107 clip -> :114 Adam
                net.refresh()
                self._stats_out.copy_(net.stats)

            if idx_all is None:                                          # caller-owned temporaries: run eagerly
                grads()
                self.dist.average_gradients(store)
                update()
            else:
                key = ("train", M, obs.data_ptr(), actions.data_ptr(), returns.data_ptr(), values.data_ptr(),
                       neglogpacs.data_ptr())
                if self.dist.active and os.environ.get("B200RL_GRAPH_NCCL", "0") != "1":
                    self.graphs.run(key + ("grads",), grads)
                    self.dist.average_gradients(store)                   # mpi_adam_optimizer.py:39-40, BEFORE the clip
                    self.graphs.run(key + ("update",), update)
                else:
                    # single process: one graph per minibatch.  (B200RL_GRAPH_NCCL=1 also captures the NCCL all-reduce;
                    # measured on 2 GPUs it is no faster than the split form -- 236.7 vs 236.6 ms -- and the process
                    # group then hangs at teardown, so the split form is the default.)
                    self.graphs.run(key, lambda: (grads(), self.dist.average_gradients(store), update()),
                                    allow_fallback=self.dist.active)
            self._after_train_call()
            return self._stats_out / M

    def _after_train_call(self):
        """mpi_adam_optimizer.py:41-42: every 100th compute_gradients call checks that the ranks still hold identical
        parameters (check_synced :53-68); a mismatch is a hard error there (assert) and here."""
        self._train_calls += 1
        if self.dist.active and self._train_calls % 100 == 0:
            if not self.dist.check_synced(self.net.store):
                raise AssertionError("parameters are not synchronised across ranks (check_synced, "
                                     "mpi_adam_optimizer.py:53-68) after {} train calls".format(self._train_calls))

    def train(self, lr, cliprange, obs, returns, masks, actions, values, neglogpacs, states=None):
        """Reference signature (model.py:133); numpy minibatch in, list of 5 python floats out."""
        if states is not None:
            raise NotImplementedError("recurrent policies are outside the hot-path scope (SURVEY.md 2, #4)")
        with torch.cuda.device(self.device):
            dev = self.device
            x = self.net.encode_obs(np.asarray(obs))
            if self.net.discrete:
                a = torch.as_tensor(np.ascontiguousarray(actions), dtype=torch.int64).to(dev)
            else:
                a = torch.as_tensor(np.ascontiguousarray(actions), dtype=torch.float32).to(dev).contiguous()
            f = lambda z: torch.as_tensor(np.ascontiguousarray(z), dtype=torch.float32).to(dev)
            st = self.train_rollout(float(lr), float(cliprange), x, a, f(returns), f(values), f(neglogpacs), None)
            return [float(s) for s in st.cpu().numpy()]

    # ------------------------------------------------------------------------------------ checkpoints
    def save(self, save_path):
        """tf_util.save_variables (tf_util.py:345-355): joblib dict {tf variable name: ndarray}.  Adam slots
        are stored under the TF slot names '<var>/Adam:0', '<var>/Adam_1:0' like the reference's global
        variables."""
        import joblib
        d = dict(self.net.store.export_tf("params"))
        for k, v in self.net.store.export_tf("m").items():
            d[k.replace(":0", "/Adam:0")] = v
        for k, v in self.net.store.export_tf("v").items():
            d[k.replace(":0", "/Adam_1:0")] = v
        d["beta1_power:0"] = np.float32(self.opt.beta1 ** (self.opt.t + 1))
        d["beta2_power:0"] = np.float32(self.opt.beta2 ** (self.opt.t + 1))
        # float32 beta1_power underflows to 0 after ~1000 Adam steps (62 PPO2 updates at the defaults), so the step
        # count cannot be recovered from it; it is stored explicitly under a key no TF variable uses
        d["b200rl/adam_t"] = np.int64(self.opt.t)
        if self.net.obs_rms is not None:                     # RunningMeanStd variables (mpi_running_mean_std.py:11-26)
            for k, name in self.net.rms_names.items():
                d[name] = np.array(self.net.obs_rms[k], dtype=np.float64)
        dirname = os.path.dirname(save_path)
        if dirname:
            os.makedirs(dirname, exist_ok=True)
        joblib.dump(d, save_path)

    def load(self, load_path):
        import joblib
        d = joblib.load(os.path.expanduser(load_path))
        store = self.net.store
        store.import_tf({k: v for k, v in d.items() if k in store.tf_map}, "params")
        store.import_tf({k.replace("/Adam:0", ":0"): v for k, v in d.items() if k.endswith("/Adam:0")}, "m")
        store.import_tf({k.replace("/Adam_1:0", ":0"): v for k, v in d.items() if k.endswith("/Adam_1:0")}, "v")
        self.opt.t = _adam_step_from_checkpoint(d, self.opt.beta1, self.opt.beta2, self.opt.t)
        if self.net.obs_rms is not None:
            self.net.set_obs_rms({k: d[name] for k, name in self.net.rms_names.items() if name in d})
        self.net.refresh()

    # parameters in the reference's TF naming / layout (used by the parity tests)
    def get_params(self):
        return self.net.store.export_tf("params")

    def set_params(self, params):
        self.net.store.import_tf(params, "params")
        self.net.refresh()


def _adam_step_from_checkpoint(d, beta1, beta2, default):
    """Adam step count of a checkpoint.  Our own files carry it as an integer; a reference (TF) checkpoint only has
    the float32 accumulators beta1_power = beta1^(t+1), beta2_power = beta2^(t+1) (tf.train.AdamOptimizer slots saved
    by tf_util.save_variables, tf_util.py:345-355).  beta1_power is denormal / zero after ~800 steps, so beta2_power
    (usable up to ~9e4 steps) is preferred; once both have underflowed the bias correction is 1 to fp32 precision and
    any large t gives the same update."""
    if "b200rl/adam_t" in d:
        return int(d["b200rl/adam_t"])
    tiny = float(np.finfo(np.float32).tiny)
    for key, beta in (("beta2_power:0", beta2), ("beta1_power:0", beta1)):
        if key in d:
            p = float(d[key])
            if p >= 1.0:
                return 0
            if p > tiny * 1e3:                       # well inside the normal range: log() is accurate
                return max(0, int(round(math.log(p) / math.log(beta))) - 1)
    if "beta1_power:0" in d or "beta2_power:0" in d:
        return 10 ** 6                               # both underflowed: sqrt(1-b2^t)/(1-b1^t) == 1
    return default


def _make_optimizer(store, max_grad_norm):
    from ..nn import Optimizer
    return Optimizer(store, eps=1e-5, max_grad_norm=max_grad_norm)       # model.py:100 epsilon=1e-5
