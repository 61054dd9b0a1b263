"""Generate tests/golden/*.npz by EXECUTING the reference's own numpy / pure-python code.

Run in the build container only (needs /root/reference, read-only):
    python oracle/gen_golden.py
The fixtures are committed; tests and the GPU box never touch /root/reference.

What is executed unmodified from the reference:
* baselines/ppo2/runner.py  Runner.run  (rollout bookkeeping, GAE :53-65, sf01 :69-74)
* baselines/common/segment_tree.py  SumSegmentTree / MinSegmentTree
* baselines/deepq/replay_buffer.py  PrioritizedReplayBuffer (loaded by file path because
  deepq/__init__.py imports TensorFlow)
`gym` is absent from the image, so an empty module is stubbed into sys.modules before
importing.
"""
import importlib.util
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.modules.setdefault("gym", types.ModuleType("gym"))
    sys.path.insert(0, REF)
    from baselines.ppo2.runner import Runner                     # noqa
    from baselines.common.segment_tree import SumSegmentTree, MinSegmentTree  # noqa
    spec = importlib.util.spec_from_file_location(
        "ref_replay_buffer", os.path.join(REF, "baselines/deepq/replay_buffer.py"))
    rb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rb)
    return Runner, SumSegmentTree, MinSegmentTree, rb


class _Space:
    def __init__(self, shape, dtype):
        self.shape = shape
        self.dtype = np.dtype(dtype)


class FakeEnv:
    """Replays pre-drawn rewards / dones; obs encode (t, env) so sf01 ordering is visible."""

    def __init__(self, rew, done, ob_shape=(2,)):
        self.rew, self.done = rew, done
        self.num_envs = rew.shape[1]
        self.observation_space = _Space(ob_shape, np.float32)
        self.t = 0

    def _obs(self):
        o = np.zeros((self.num_envs,) + self.observation_space.shape, np.float32)
        o[:, 0] = self.t
        o[:, 1] = np.arange(self.num_envs)
        return o

    def reset(self):
        return self._obs()

    def step(self, actions):
        r, d = self.rew[self.t], self.done[self.t]
        self.t += 1
        return self._obs(), r, d, [{} for _ in range(self.num_envs)]


class FakeModel:
    initial_state = None

    def __init__(self, val):
        self.val = val
        self.t = 0

    def step(self, obs, S=None, M=None):
        n = obs.shape[0]
        v = self.val[self.t]
        self.t += 1
        return np.full(n, self.t, np.int64), v, None, (v * 0.5).astype(np.float32)

    def value(self, obs, S=None, M=None):
        return self.val[self.t]


def gen_gae(Runner, name, T, N, seed, p_done, nrollouts=1):
    rng = np.random.RandomState(seed)
    REW = rng.randn(T * nrollouts, N).astype(np.float32)
    VAL = rng.randn(T * nrollouts + 1, N).astype(np.float32)
    DONE = rng.rand(T * nrollouts, N) < p_done
    env, model = FakeEnv(REW, DONE), FakeModel(VAL)
    runner = Runner(env=env, model=model, nsteps=T, gamma=0.99, lam=0.95)
    out = {}
    for k in range(nrollouts):
        first_dones = np.asarray(runner.dones, dtype=np.bool_).copy()
        # model.value() is asked for VAL[(k+1)*T]; FakeModel.t points there after T steps
        obs, returns, masks, actions, values, neglogpacs, states, epinfos = runner.run()
        model.t = (k + 1) * T  # value() did not advance t; next rollout's first step reads VAL[(k+1)T]
        out[f"obs{k}"], out[f"returns{k}"], out[f"masks{k}"] = obs, returns, masks
        out[f"actions{k}"], out[f"values{k}"], out[f"neglogpacs{k}"] = actions, values, neglogpacs
        out[f"first_dones{k}"] = first_dones
        out[f"last_dones{k}"] = np.asarray(runner.dones, dtype=np.bool_).copy()
    np.savez_compressed(os.path.join(OUT, name), REW=REW, VAL=VAL, DONE=DONE, T=T, N=N,
                        gamma=0.99, lam=0.95, nrollouts=nrollouts, **out)
    return out


def _import_reference_frame_stack():
    """vec_frame_stack.py needs `gym.spaces.Box` (third-party, absent here): a 4-attribute stand-in is put into the
    gym stub.  The package __init__ (which drags in gym.core via VecMonitor) is bypassed by loading the two files
    under a synthetic package so the relative import `.vec_env` still resolves."""
    gym = sys.modules.setdefault("gym", types.ModuleType("gym"))
    sp = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), np.dtype(dtype)
            self.shape = self.low.shape

    sp.Box = Box
    gym.spaces = sp
    sys.modules["gym.spaces"] = sp
    d = os.path.join(REF, "baselines/common/vec_env")
    pkg = types.ModuleType("refvec")
    pkg.__path__ = [d]
    sys.modules["refvec"] = pkg
    mods = {}
    for name in ("vec_env", "vec_frame_stack"):
        spec = importlib.util.spec_from_file_location("refvec." + name, os.path.join(d, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refvec." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["vec_frame_stack"].VecFrameStack, Box


def gen_frame_stack():
    """Drive the reference VecFrameStack with a scripted venv: random frames, scripted episode ends."""
    VecFrameStack, Box = _import_reference_frame_stack()
    for name, N, hw, c, nstack, T, p_done, seed in (("frame_stack_c1.npz", 5, (6, 4), 1, 4, 12, 0.2, 11),
                                                    ("frame_stack_c2.npz", 3, (3, 5), 2, 3, 9, 0.3, 12)):
        rng = np.random.RandomState(seed)
        frames = rng.randint(0, 256, size=(T + 1, N) + hw + (c,)).astype(np.uint8)
        news = rng.rand(T, N) < p_done

        class Venv:
            num_envs = N
            observation_space = Box(np.zeros(hw + (c,), np.uint8), np.full(hw + (c,), 255, np.uint8), dtype=np.uint8)
            action_space = None
            t = 0

            def reset(self):
                self.t = 0
                return frames[0]

            def step_async(self, actions):
                pass

            def step_wait(self):
                self.t += 1
                return frames[self.t], np.zeros(N, np.float32), news[self.t - 1], [{}] * N

        env = VecFrameStack(Venv(), nstack)
        out = [env.reset().copy()]
        for t in range(T):
            env.step_async(None)
            o, _, _, _ = env.step_wait()
            out.append(o.copy())
        np.savez_compressed(os.path.join(OUT, name), frames=frames, news=news, stacked=np.stack(out),
                            nstack=nstack, c=c)


def gen_vec_normalize():
    """Reference VecNormalize (vec_normalize.py, numpy RunningMeanStd) driven by a scripted venv.  `tensorflow` and
    `baselines.common.tf_util` are stubbed (running_mean_std.py imports them at module level; the numpy class does
    not use them)."""
    _import_reference_frame_stack()                      # installs the synthetic package `refvec`
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    tfu = types.ModuleType("baselines.common.tf_util")
    tfu.get_session = lambda *a, **k: None
    sys.modules.setdefault("baselines.common.tf_util", tfu)
    pkg = sys.modules["refvec"]
    pkg.VecEnvWrapper = sys.modules["refvec.vec_env"].VecEnvWrapper
    d = os.path.join(REF, "baselines/common/vec_env")
    spec = importlib.util.spec_from_file_location("refvec.vec_normalize", os.path.join(d, "vec_normalize.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["refvec.vec_normalize"] = m
    spec.loader.exec_module(m)
    N, D, T = 4, 3, 40
    rng = np.random.RandomState(21)
    obs = (rng.randn(T + 1, N, D) * np.array([1.0, 5.0, 0.1]) + np.array([0.0, 2.0, -1.0])).astype(np.float32)
    rews = rng.randn(T, N).astype(np.float32) * 3.0
    news = rng.rand(T, N) < 0.1

    class Venv:
        num_envs = N
        observation_space = _Space((D,), np.float32)
        action_space = None
        t = 0

        def reset(self):
            self.t = 0
            return obs[0]

        def step_async(self, a):
            pass

        def step_wait(self):
            self.t += 1
            return obs[self.t], rews[self.t - 1], news[self.t - 1], [{}] * N

    env = m.VecNormalize(Venv())
    out_o, out_r = [np.asarray(env.reset()).copy()], []
    for t in range(T):
        env.step_async(None)
        o, r, _, _ = env.step_wait()
        out_o.append(np.asarray(o).copy())
        out_r.append(np.asarray(r).copy())
    np.savez_compressed(os.path.join(OUT, "vec_normalize_trace.npz"), obs=obs, rews=rews, news=news,
                        norm_obs=np.stack(out_o), norm_rews=np.stack(out_r), ob_mean=env.ob_rms.mean,
                        ob_var=env.ob_rms.var, ob_count=env.ob_rms.count, ret_var=env.ret_rms.var)


def gen_host_misc():
    """Small host helpers the learn loops call: schedules (deepq.py:222-231) and explained_variance (ppo2.py:195)."""
    sys.path.insert(0, REF)
    from baselines.common.schedules import LinearSchedule, PiecewiseSchedule, ConstantSchedule
    from baselines.common.math_util import explained_variance
    ts = np.array([0, 1, 7, 99, 100, 101, 5000, 10000, 12345, 10 ** 6], np.int64)
    lin = LinearSchedule(schedule_timesteps=int(0.1 * 100000), initial_p=1.0, final_p=0.02)
    beta = LinearSchedule(100000, initial_p=0.4, final_p=1.0)
    pw = PiecewiseSchedule([(0, 1.0), (100, 0.5), (10000, 0.1)], outside_value=0.05)
    rng = np.random.RandomState(4)
    y = rng.randn(257).astype(np.float32)
    yp = (y + 0.3 * rng.randn(257)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "host_misc.npz"), ts=ts,
                        lin=np.array([lin.value(int(t)) for t in ts]), beta=np.array([beta.value(int(t)) for t in ts]),
                        pw=np.array([pw.value(int(t)) for t in ts]), const=ConstantSchedule(0.7).value(3),
                        y=y, yp=yp, ev=explained_variance(yp, y), ev_perfect=explained_variance(y, y),
                        ev_const=explained_variance(yp, np.ones_like(y)))


class _TfStub(types.ModuleType):
    """Any attribute / call resolves to another stub: lets reference modules that `import tensorflow as tf` be
    imported so that their PURE-NUMPY functions can be executed (nothing TF-backed is ever called)."""

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _TfStub(k)

    def __call__(self, *a, **k):
        return _TfStub('call')


def gen_init_and_adam():
    """(1) a2c/utils.py:20-35 ortho_init is numpy + SVD under a TF signature -> executed for the layer shapes of
    nature_cnn / mlp / heads under fixed seeds.  (2) common/mpi_adam.py:25-42 MpiAdam.update is the reference's
    numpy statement of TF-Adam -> executed on a bare instance (no TF variables: getflat/setfromflat are closures over
    a numpy vector, comm=None) for 5 steps."""
    saved_tf = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = _TfStub("tensorflow")
    try:
        spec = importlib.util.spec_from_file_location("ref_a2c_utils", os.path.join(REF, "baselines/a2c/utils.py"))
        u = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(u)
        out = {}
        shapes = [((8, 8, 4, 32), np.sqrt(2)), ((4, 4, 32, 64), np.sqrt(2)), ((3, 3, 64, 64), np.sqrt(2)),
                  ((64, 17), 0.01), ((376, 64), np.sqrt(2)), ((64, 1), 1.0), ((5, 9), 1.0), ((9, 5), 1.0)]
        np.random.seed(1234)
        for i, (shp, sc) in enumerate(shapes):
            out[f"w{i}"] = u.ortho_init(sc)(shp, np.float32)
            out[f"shape{i}"] = np.array(shp)
            out[f"scale{i}"] = sc
        sys.modules["baselines.common.tf_util"] = _TfStub("baselines.common.tf_util")
        spec = importlib.util.spec_from_file_location("ref_mpi_adam", os.path.join(REF, "baselines/common/mpi_adam.py"))
        ma = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ma)
        rng = np.random.RandomState(7)
        theta = rng.randn(257).astype(np.float32)
        out["adam_theta0"] = theta.copy()
        opt = object.__new__(ma.MpiAdam)
        opt.beta1, opt.beta2, opt.epsilon, opt.scale_grad_by_procs, opt.comm, opt.t = 0.9, 0.999, 1e-5, True, None, 0
        opt.m, opt.v = np.zeros(257, 'float32'), np.zeros(257, 'float32')
        state = {"theta": theta}
        opt.getflat = lambda: state["theta"]
        opt.setfromflat = lambda x: state.__setitem__("theta", x)
        grads, thetas = [], []
        for k in range(5):
            g = (rng.randn(257) * (10.0 ** rng.randint(-3, 2))).astype(np.float32)
            opt.update(g, 2.5e-4 * (1 - 0.1 * k))
            grads.append(g)
            thetas.append(np.asarray(state["theta"]).copy())
        out.update(adam_grads=np.stack(grads), adam_thetas=np.stack(thetas), adam_m=opt.m, adam_v=opt.v)
        np.savez_compressed(os.path.join(OUT, "init_adam.npz"), **out)
    finally:
        if saved_tf is None:
            sys.modules.pop("tensorflow", None)
        else:
            sys.modules["tensorflow"] = saved_tf
        sys.modules.pop("baselines.common.tf_util", None)


def gen_segment_tree(Sum, Min):
    rng = np.random.RandomState(7)
    cap = 64
    s, m = Sum(cap), Min(cap)
    ops = []          # (kind, a, b, result)
    for step in range(600):
        kind = rng.randint(0, 4)
        if kind == 0 or step < 40:
            i, v = int(rng.randint(0, cap)), float(rng.rand() * 3 + 1e-3)
            s[i] = v
            m[i] = v
            ops.append((0, i, v, 0.0))
        elif kind == 1:
            a = int(rng.randint(0, cap - 1))
            b = int(rng.randint(a + 1, cap + 1))
            ops.append((1, a, b, float(s.sum(a, b))))
            ops.append((2, a, b, float(m.min(a, b))))
        elif kind == 2:
            a = int(rng.randint(0, cap - 2))
            b = -int(rng.randint(1, cap - a - 1))           # negative `end` wraps (segment_tree.py:71-72)
            ops.append((1, a, b, float(s.sum(a, b))))
        else:
            p = float(rng.rand() * s.sum())
            ops.append((3, 0, p, float(s.find_prefixsum_idx(p))))
    arr = np.array(ops, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "segment_tree_trace.npz"), ops=arr, capacity=cap,
                        final_sum=np.array(s._value, np.float64), final_min=np.array(m._value, np.float64))


def gen_per(rb):
    """PrioritizedReplayBuffer trace: adds (with ring wrap), samples with python's `random`
    seeded so the uniforms can be regenerated, priority updates, more samples."""
    size, alpha, batch = 100, 0.6, 16            # capacity rounds to 128; ring wraps at 100
    buf = rb.PrioritizedReplayBuffer(size, alpha)
    rng = np.random.RandomState(11)
    trace = {}
    nadd1 = 70
    for i in range(nadd1):
        buf.add(np.full((2,), i, np.float32), np.array(i % 3), float(i), np.full((2,), i + 1, np.float32), float(i % 7 == 0))
    rounds = []
    uni_all, idx_all, w_all, prio_all, beta_all, nstored = [], [], [], [], [], []
    added = nadd1
    for r in range(6):
        beta = 0.4 + 0.1 * r
        random.seed(1000 + r)
        uniforms = [random.random() for _ in range(batch)]
        random.seed(1000 + r)
        out = buf.sample(batch, beta)
        weights, idxes = out[5], out[6]
        prios = (np.abs(rng.randn(batch)) + 1e-6) * (3.0 if r == 2 else 1.0)
        buf.update_priorities(idxes, prios)
        uni_all.append(uniforms); idx_all.append(idxes); w_all.append(weights)
        prio_all.append(prios); beta_all.append(beta); nstored.append(len(buf))
        nadd = 25
        for i in range(nadd):                   # keeps adding -> wraps the ring on later rounds
            buf.add(np.zeros(2, np.float32), np.array(0), 0.0, np.zeros(2, np.float32), 0.0)
        added += nadd
        rounds.append(nadd)
    np.savez_compressed(os.path.join(OUT, "per_trace.npz"), size=size, alpha=alpha, batch=batch, nadd1=nadd1,
                        adds_after_round=np.array(rounds), uniforms=np.array(uni_all, np.float64),
                        idxes=np.array(idx_all, np.int64), weights=np.array(w_all, np.float64),
                        priorities=np.array(prio_all, np.float64), betas=np.array(beta_all, np.float64),
                        nstored=np.array(nstored), final_sum=np.array(buf._it_sum._value, np.float64),
                        final_min=np.array(buf._it_min._value, np.float64), max_priority=buf._max_priority)

    # the p_total quirk (replay_buffer.py:109): sum(0, len-1) drops the last stored element
    b2 = rb.PrioritizedReplayBuffer(4, 1.0)
    for i in range(4):
        b2.add(np.zeros(1), np.array(0), 0.0, np.zeros(1), 0.0)
    np.savez_compressed(os.path.join(OUT, "per_ptotal_quirk.npz"),
                        p_total_used=b2._it_sum.sum(0, len(b2._storage) - 1), full_sum=b2._it_sum.sum())


def gen_uniform_replay(rb):
    """Uniform ReplayBuffer (deepq/replay_buffer.py:7-68): ring writes with wrap-around, `sample` drawing
    random.randint positions (:67), `_encode_sample` return types (:33-43: rewards / dones become float64)."""
    size, batch = 50, 12
    buf = rb.ReplayBuffer(size)
    rng = np.random.RandomState(21)
    n_add = [30, 15, 40]                                  # second and third rounds wrap the ring
    adds, samples = [], []
    k = 0
    for r, na in enumerate(n_add):
        for _ in range(na):
            o, o1 = rng.randint(0, 256, (3, 3, 2)).astype(np.uint8), rng.randint(0, 256, (3, 3, 2)).astype(np.uint8)
            a, rew, d = int(rng.randint(4)), float(rng.randn()), float(rng.rand() < 0.2)
            buf.add(o, np.array(a), rew, o1, d)
            adds.append((o, a, rew, o1, d))
            k += 1
        random.seed(500 + r)
        obs_t, act, rews, obs_tp1, dones = buf.sample(batch)
        assert rews.dtype == np.float64 and dones.dtype == np.float64
        samples.append((obs_t, act, rews, obs_tp1, dones, len(buf)))
    np.savez_compressed(os.path.join(OUT, "replay_uniform_trace.npz"), size=size, batch=batch, n_add=np.array(n_add),
                        add_obs=np.stack([a[0] for a in adds]), add_act=np.array([a[1] for a in adds]),
                        add_rew=np.array([a[2] for a in adds]), add_obs1=np.stack([a[3] for a in adds]),
                        add_done=np.array([a[4] for a in adds]),
                        s_obs=np.stack([s[0] for s in samples]), s_act=np.stack([s[1] for s in samples]),
                        s_rew=np.stack([s[2] for s in samples]), s_obs1=np.stack([s[3] for s in samples]),
                        s_done=np.stack([s[4] for s in samples]), s_len=np.array([s[5] for s in samples]))


def main():
    os.makedirs(OUT, exist_ok=True)
    Runner, Sum, Min, rb = _import_reference()
    if "--only-uniform-replay" in sys.argv:
        gen_uniform_replay(rb)
        return
    small = gen_gae(Runner, "gae_small.npz", T=8, N=3, seed=1234, p_done=0.25)
    # SURVEY.md section 8c vector (regenerated here; assert it reproduces)
    adv0 = (small["returns0"] - small["values0"])[:8]
    ref_adv0 = np.array([-1.2307085, 1.5043753, 1.2574286, -2.3858485, 0.8286112, -1.1049898, 1.3270301, 0.9890473], np.float32)
    assert np.allclose(adv0, ref_adv0, atol=2e-6), (adv0, ref_adv0)
    gen_gae(Runner, "gae_medium.npz", T=128, N=64, seed=5, p_done=0.02)
    gen_gae(Runner, "gae_two_rollouts.npz", T=16, N=5, seed=9, p_done=0.2, nrollouts=2)
    gen_gae(Runner, "gae_alldone.npz", T=6, N=4, seed=3, p_done=1.1)
    gen_segment_tree(Sum, Min)
    gen_per(rb)
    gen_uniform_replay(rb)
    gen_frame_stack()
    gen_vec_normalize()
    gen_host_misc()
    gen_init_and_adam()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
