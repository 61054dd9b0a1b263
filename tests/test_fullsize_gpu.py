"""GPU: BASELINE.json full-size configurations checked through size-independent properties and, where the CPU
oracle is cheap enough, directly (SURVEY.md 8d sizes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gae_cfg3_full_size_bit_exact():
    """cfg-3: T=512, N=16384 (142.6 MB of algorithmic traffic) -- the numpy oracle finishes in about a second."""
    from baselines_b200 import ops
    from oracle.gae import gae_reference_order
    T, N = 512, 16384
    rng = np.random.RandomState(0)
    rew = rng.randn(T, N).astype(np.float32)
    val = rng.randn(T, N).astype(np.float32)
    done = rng.rand(T, N) < 0.01
    lv = rng.randn(N).astype(np.float32)
    ld = rng.rand(N) < 0.01
    c = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    adv = torch.empty(T, N, device="cuda")
    ret = torch.empty(T, N, device="cuda")
    ops.gae_scan(c(rew), c(val), c(done.astype(np.uint8)), c(lv), c(ld.astype(np.uint8)), adv, ret, 0.99, 0.95)
    adv_o, ret_o = gae_reference_order(rew, val, done, lv, ld, 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), adv_o) and np.array_equal(ret.cpu().numpy(), ret_o)
    # size-independent property: lambda = 0 reduces to the one-step TD error
    ops.gae_scan(c(rew), c(val), c(done.astype(np.uint8)), c(lv), c(ld.astype(np.uint8)), adv, ret, 0.99, 0.0)
    nv = np.concatenate([val[1:], lv[None]], 0)
    nnt = 1.0 - np.concatenate([done[1:], ld[None]], 0)
    td = (rew + (np.float32(0.99) * nv) * nnt - val).astype(np.float32)
    assert np.array_equal(adv.cpu().numpy(), td)


def test_per_cfg4_capacity_2pow20_properties():
    """cfg-4: prioritized replay with 1M slots (tree 2^20): tree invariants + sampling properties."""
    from baselines_b200 import ops
    cap, batch = 1 << 20, 512
    dev = "cuda"
    s = torch.zeros(2 * cap, dtype=torch.float64, device=dev)
    m = torch.full((2 * cap,), float("inf"), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    pri = (torch.randn(cap, device=dev, generator=g, dtype=torch.float64).abs() + 1e-6) ** 0.6
    idx_all = torch.arange(cap, device=dev)
    for o in range(0, cap, 1 << 16):                       # 16 launches-groups of 64 chunks
        ops.tree_set(s, m, cap, idx_all[o:o + (1 << 16)], pri[o:o + (1 << 16)])
    torch.cuda.synchronize()
    sv = s.cpu().numpy()
    mv = m.cpu().numpy()
    # every internal node is EXACTLY op(left, right) (same recomputation rule as segment_tree.py:76-86)
    assert np.array_equal(sv[1:cap], sv[2:2 * cap:2] + sv[3:2 * cap:2])
    assert np.array_equal(mv[1:cap], np.minimum(mv[2:2 * cap:2], mv[3:2 * cap:2]))
    assert np.array_equal(sv[cap:], pri.cpu().numpy())
    u = torch.rand(batch, device=dev, generator=g, dtype=torch.float64)
    idx = torch.empty(batch, dtype=torch.int64, device=dev)
    w = torch.empty(batch, dtype=torch.float64, device=dev)
    ops.per_sample(s, m, cap, cap, u, 0.4, idx, w)
    ii, ww = idx.cpu().numpy(), w.cpu().numpy()
    assert ii.min() >= 0 and ii.max() < cap
    assert np.all(np.diff(ii) >= 0)                        # stratified masses are increasing -> indices sorted
    assert ww.max() <= 1.0 + 1e-12 and ww.min() > 0
    # the prefix sum up to the sampled leaf brackets the stratum mass (find_prefixsum_idx contract)
    leaves = sv[cap:]
    csum = np.cumsum(leaves)
    p_total = csum[cap - 2]                                # reference quirk: sum(0, len-1) drops the last element
    mass = (u.cpu().numpy() + np.arange(batch)) * (p_total / batch)
    assert np.all(csum[ii] >= mass * (1 - 1e-9)) and np.all((csum[ii] - leaves[ii]) <= mass * (1 + 1e-9))


def test_ppo2_mlp_humanoid_shape_update():
    """cfg-3 network shape (obs 376, 17-d Gaussian, value_network='copy') through learn() on a device env."""
    from baselines_b200.common.vec_env import DeviceSyntheticVecEnv
    from baselines_b200.ppo2 import ppo2
    env = DeviceSyntheticVecEnv(1024, (376,), np.float32, act_dim=17, seed=0)
    model = ppo2.learn(network="mlp", env=env, total_timesteps=1024 * 64 * 2, seed=0, nsteps=64, nminibatches=32,
                       noptepochs=2, lr=lambda f: 3e-4 * f, cliprange=0.2, value_network="copy", log_interval=100,
                       comm=False)
    p = model.get_params()
    assert all(np.isfinite(v).all() for v in p.values())
    assert p["ppo2_model/pi/mlp_fc0/w:0"].shape == (376, 64) and p["ppo2_model/pi/logstd:0"].shape == (1, 17)
    assert p["ppo2_model/vf/mlp_fc1/w:0"].shape == (64, 64)
    assert sum(v.size for v in p.values()) == 57763          # 29 410 + 28 353 parameters
