"""CPU: pin the oracle against golden vectors produced by EXECUTING the reference
(oracle/gen_golden.py) and against the reference's own known-answer tests."""
import os
import random

import numpy as np
import pytest

from oracle import gae as ogae
from oracle.segment_tree import SumTree, MinTree, PrioritizedSampler

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name", ["gae_small.npz", "gae_medium.npz", "gae_two_rollouts.npz", "gae_alldone.npz"])
def test_gae_oracle_bit_exact_vs_reference_runner(name):
    g = _load(name)
    T, N, K = int(g["T"]), int(g["N"]), int(g["nrollouts"])
    for k in range(K):
        rew = g["REW"][k * T:(k + 1) * T]
        val = g["VAL"][k * T:(k + 1) * T]
        last_val = g["VAL"][(k + 1) * T]
        # mb_dones[t] = done BEFORE step t (runner.py:34): first row = runner.dones at entry
        dones_before = np.concatenate([g[f"first_dones{k}"][None], g["DONE"][k * T:(k + 1) * T - 1]], 0)
        last_dones = g["DONE"][(k + 1) * T - 1]
        assert np.array_equal(last_dones, g[f"last_dones{k}"])
        adv, ret = ogae.gae_reference_order(rew, val, dones_before, last_val, last_dones,
                                            float(g["gamma"]), float(g["lam"]))
        assert np.array_equal(ogae.sf01(ret), g[f"returns{k}"])          # bit exact
        assert np.array_equal(ogae.sf01(val), g[f"values{k}"])
        assert np.array_equal(ogae.sf01(dones_before), g[f"masks{k}"])
        # env-major flat order i = e*T + t (runner.py:69-74): obs encode (t, env)
        obs = g[f"obs{k}"]
        assert np.array_equal(obs[:, 1], np.repeat(np.arange(N), T))
        assert np.array_equal(obs[:, 0] - k * T, np.tile(np.arange(T), N))


def test_gae_survey_vector():
    """env 0 of the (T,N)=(8,3) seed-1234 case."""
    g = _load("gae_small.npz")
    adv = (g["returns0"] - g["values0"])[:8]
    ref = np.array([-1.2307085, 1.5043753, 1.2574286, -2.3858485, 0.8286112, -1.1049898, 1.3270301, 0.9890473])
    assert np.allclose(adv, ref, atol=2e-6)
    assert np.array_equal(g["masks0"][:8].astype(int), [0, 0, 1, 1, 0, 1, 0, 0])
    assert abs(float((g["returns0"] - g["values0"]).sum()) - 3.0016775) < 1e-4


# ---- the reference's known-answer tests (baselines/common/tests/test_segment_tree.py) ----
def test_tree_set():                     # :6-18
    t = SumTree(4)
    t.set(2, 1.0); t.set(3, 3.0)
    assert np.isclose(t.sum(), 4.0)
    assert np.isclose(t.sum(0, 2), 0.0)
    assert np.isclose(t.sum(0, 3), 1.0)
    assert np.isclose(t.sum(2, 3), 1.0)
    assert np.isclose(t.sum(2, -1), 1.0)
    assert np.isclose(t.sum(2, 4), 4.0)


def test_tree_set_overlap():             # :21-32
    t = SumTree(4)
    t.set(2, 1.0); t.set(2, 3.0)
    assert np.isclose(t.sum(), 3.0)
    assert np.isclose(t.sum(2, 3), 3.0)
    assert np.isclose(t.sum(2, -1), 3.0)
    assert np.isclose(t.sum(2, 4), 3.0)
    assert np.isclose(t.sum(1, 2), 0.0)


def test_prefixsum_idx():                # :35-46
    t = SumTree(4)
    t.set(2, 1.0); t.set(3, 3.0)
    for p, want in [(0.0, 2), (0.5, 2), (0.99, 2), (1.01, 3), (3.00, 3), (4.00, 3)]:
        assert t.find_prefixsum_idx(p) == want


def test_prefixsum_idx2():               # :49-62
    t = SumTree(4)
    for i, v in enumerate([0.5, 1.0, 1.0, 3.0]):
        t.set(i, v)
    for p, want in [(0.00, 0), (0.55, 1), (0.99, 1), (1.51, 2), (3.00, 3), (5.50, 3)]:
        assert t.find_prefixsum_idx(p) == want


def test_max_interval_tree():            # :65-95
    t = MinTree(4)
    t.set(0, 1.0); t.set(2, 0.5); t.set(3, 3.0)
    chk = lambda vals: [np.isclose(t.min(*a), v) for a, v in vals]
    assert all(chk([((), 0.5), ((0, 2), 1.0), ((0, 3), 0.5), ((0, -1), 0.5), ((2, 4), 0.5), ((3, 4), 3.0)]))
    t.set(2, 0.7)
    assert all(chk([((), 0.7), ((0, 2), 1.0), ((0, 3), 0.7), ((0, -1), 0.7), ((2, 4), 0.7), ((3, 4), 3.0)]))
    t.set(2, 4.0)
    assert all(chk([((), 1.0), ((0, 2), 1.0), ((0, 3), 1.0), ((0, -1), 1.0), ((2, 4), 3.0), ((2, 3), 4.0),
                    ((2, -1), 4.0), ((3, 4), 3.0)]))


def test_segment_tree_trace_bit_exact():
    g = _load("segment_tree_trace.npz")
    cap = int(g["capacity"])
    s, m = SumTree(cap), MinTree(cap)
    for kind, a, b, res in g["ops"]:
        kind = int(kind)
        if kind == 0:
            s.set(int(a), b); m.set(int(a), b)
        elif kind == 1:
            assert s.sum(int(a), int(b)) == res
        elif kind == 2:
            assert m.min(int(a), int(b)) == res
        else:
            assert s.find_prefixsum_idx(b) == int(res)
    assert np.array_equal(s.value, g["final_sum"])
    assert np.array_equal(m.value, g["final_min"])


def test_per_trace_bit_exact():
    g = _load("per_trace.npz")
    per = PrioritizedSampler(int(g["size"]), float(g["alpha"]))
    for _ in range(int(g["nadd1"])):
        per.add()
    for r in range(len(g["betas"])):
        random.seed(1000 + r)                       # same python-RNG stream the reference consumed
        uniforms = [random.random() for _ in range(int(g["batch"]))]
        assert np.array_equal(uniforms, g["uniforms"][r])
        assert per.n == int(g["nstored"][r])
        idx = per.sample_idx(uniforms)
        assert np.array_equal(idx, g["idxes"][r])
        w = per.weights(idx, float(g["betas"][r]))
        assert np.array_equal(w, g["weights"][r])   # float64, bit exact
        per.update_priorities(idx, g["priorities"][r])
        for _ in range(int(g["adds_after_round"][r])):
            per.add()
    assert np.array_equal(per.sum_tree.value, g["final_sum"])
    assert np.array_equal(per.min_tree.value, g["final_min"])
    assert per.max_priority == float(g["max_priority"])


def test_per_ptotal_quirk():
    g = _load("per_ptotal_quirk.npz")
    per = PrioritizedSampler(4, 1.0)
    for _ in range(4):
        per.add()
    assert per.sum_tree.sum(0, per.n - 1) == float(g["p_total_used"]) == 3.0
    assert per.sum_tree.sum() == float(g["full_sum"]) == 4.0


@pytest.mark.parametrize("name", ["frame_stack_c1.npz", "frame_stack_c2.npz"])
def test_frame_stack_oracle_matches_reference(name):
    """oracle.frame_stack vs outputs of the reference VecFrameStack (vec_frame_stack.py) on a scripted venv."""
    from oracle import frame_stack as fs
    g = np.load(os.path.join(GOLDEN, name))
    frames, news, want = g["frames"], g["news"], g["stacked"]
    cur = fs.frame_stack_reset(frames[0], int(g["nstack"]))
    assert np.array_equal(cur, want[0])
    for t in range(news.shape[0]):
        cur = fs.frame_stack_step(cur, frames[t + 1], news[t])
        assert np.array_equal(cur, want[t + 1]), t


def test_host_vec_frame_stack_matches_reference():
    """baselines_b200.common.vec_env.VecFrameStack (host path) reproduces the reference outputs."""
    from baselines_b200.common import spaces
    from baselines_b200.common.vec_env import VecEnv, VecFrameStack
    g = np.load(os.path.join(GOLDEN, "frame_stack_c1.npz"))
    frames, news, want = g["frames"], g["news"], g["stacked"]

    class Scripted(VecEnv):
        def __init__(self):
            super().__init__(frames.shape[1], spaces.Box(0, 255, frames.shape[2:], np.uint8), spaces.Discrete(2))
            self.t = 0

        def reset(self):
            self.t = 0
            return frames[0]

        def step_async(self, actions):
            pass

        def step_wait(self):
            self.t += 1
            return frames[self.t], np.zeros(self.num_envs, np.float32), news[self.t - 1], [{}] * self.num_envs

    env = VecFrameStack(Scripted(), int(g["nstack"]))
    assert env.observation_space.shape == want.shape[2:]
    assert np.array_equal(env.reset(), want[0])
    for t in range(news.shape[0]):
        o, _, d, _ = env.step(None)
        assert np.array_equal(o, want[t + 1]) and np.array_equal(d, news[t])


def test_ortho_init_matches_executed_reference():
    """a2c/utils.py:20-35 executed (numpy + SVD under a TF stub) for the nature_cnn / mlp / head shapes: the oracle's
    and the product's initialisers consume the global RandomState identically and return the same matrices.  The
    fixture comparison allows 1e-6 (LAPACK kernels differ between host CPUs); oracle vs product is exact."""
    import torch  # noqa: F401  (baselines_b200.nn imports it)
    from baselines_b200 import nn
    from oracle import nets
    g = np.load(os.path.join(GOLDEN, "init_adam.npz"))
    shapes = [(tuple(int(x) for x in g[f"shape{i}"]), float(g[f"scale{i}"])) for i in range(8)]
    np.random.seed(1234)
    wo = [nets.ortho_init_np(s, sc) for s, sc in shapes]
    np.random.seed(1234)
    wp = [nn.ortho_init(s, sc) for s, sc in shapes]
    for i, (a, b) in enumerate(zip(wo, wp)):
        assert a.dtype == np.float32 and a.shape == shapes[i][0]
        assert np.array_equal(a, b), i
        assert np.allclose(a, g[f"w{i}"], atol=1e-6, rtol=0), (i, float(np.abs(a - g[f"w{i}"]).max()))


def test_adam_matches_executed_reference_numpy_adam():
    """common/mpi_adam.py:25-42 (the reference's numpy statement of TF-Adam, proven equal to
    tf.train.AdamOptimizer by its own test :64-99) executed for 5 steps.  m and v are float32 in both and must be
    identical; the reference's parameter vector is promoted to float64 under numpy >= 2 (np.float64 step size times
    a float32 array), ours stays float32 like TF's: agreement to float32 rounding."""
    import torch
    from oracle import nets
    g = np.load(os.path.join(GOLDEN, "init_adam.npz"))
    p = torch.from_numpy(g["adam_theta0"].copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for k in range(5):
        p, m, v = nets.adam_tf(p, torch.from_numpy(g["adam_grads"][k]), m, v, k + 1, 2.5e-4 * (1 - 0.1 * k), eps=1e-5)
        assert np.allclose(p.numpy(), g["adam_thetas"][k], atol=1e-6, rtol=0), k
    assert np.array_equal(m.numpy(), g["adam_m"]) and np.array_equal(v.numpy(), g["adam_v"])
