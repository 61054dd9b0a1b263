"""CPU: anchors for oracle/nets.py that do not go through its own torch code.

The TF1 graph cannot be executed here (TensorFlow 1.x is absent), so `oracle/nets.py` is a restatement.  These tests
tie it to three things that are independent of it:
  1. definition-level numpy loops of the TF op semantics the reference calls (tf.nn.conv2d NHWC VALID / SAME with
     HWIO filters a2c/utils.py:50-56, tf.matmul :63, softmax_cross_entropy_with_logits_v2 distributions.py:181,
     clip_by_global_norm ppo2/model.py:105-107);
  2. float64 central finite differences of the scalar loss (ppo2/model.py:57-91, deepq/build_graph.py:388-413)
     against the autograd gradients the oracle hands to the parity tests;
  3. the statistical identities the reference's own test pins for its distributions
     (common/distributions.py:299-348: E[-log p] = entropy and KL[p,q] = -H[p] - E_p[log q], both within 3 sigma,
     N = 100 000, same parameter vectors).
"""
import math

import numpy as np
import torch

from oracle import nets


# ------------------------------------------------------------------------------------------------- 1. op definitions
def _conv2d_nhwc_loops(x, w, stride, pad):
    """tf.nn.conv2d(x, w, strides=[1,s,s,1], padding=pad, data_format='NHWC') from its definition:
    out[b,i,j,k] = sum_{di,dj,q} x[b, s*i+di-pt, s*j+dj-pl, q] * w[di,dj,q,k]; SAME pads
    total = max((ceil(in/s)-1)*s + rf - in, 0), top/left = total // 2."""
    B, H, W, C = x.shape
    rf, _, _, nf = w.shape
    if pad == "VALID":
        oh, ow, pt, pl = (H - rf) // stride + 1, (W - rf) // stride + 1, 0, 0
    else:
        oh, ow = -(-H // stride), -(-W // stride)
        pt = max((oh - 1) * stride + rf - H, 0) // 2
        pl = max((ow - 1) * stride + rf - W, 0) // 2
    out = np.zeros((B, oh, ow, nf), np.float64)
    for i in range(oh):
        for j in range(ow):
            for di in range(rf):
                for dj in range(rf):
                    y, xx = stride * i + di - pt, stride * j + dj - pl
                    if 0 <= y < H and 0 <= xx < W:
                        out[:, i, j, :] += x[:, y, xx, :].astype(np.float64) @ w[di, dj].astype(np.float64)
    return out


def test_conv_restatement_matches_tf_definition():
    rng = np.random.RandomState(0)
    for (H, C, rf, st, nf, pad) in [(12, 3, 4, 2, 5, "VALID"), (9, 4, 3, 1, 6, "VALID"), (11, 2, 8, 4, 3, "SAME"),
                                    (10, 3, 4, 2, 4, "SAME"), (7, 5, 3, 1, 2, "SAME")]:
        x = rng.randn(2, H, H, C)
        w = rng.randn(rf, rf, C, nf)
        b = rng.randn(1, nf, 1, 1)
        got = nets._conv_nhwc(torch.tensor(x), torch.tensor(w), torch.tensor(b), st, pad).numpy()
        want = _conv2d_nhwc_loops(x, w, st, pad) + b.reshape(1, 1, 1, nf)
        assert got.shape == want.shape, (H, rf, st, pad)
        assert np.allclose(got, want, atol=1e-10), (H, rf, st, pad)


def test_nature_cnn_shapes_and_flatten_order():
    """models.py:15-26: 84x84x4 -> 20x20x32 -> 9x9x64 -> 7x7x64 -> 3136 (H, W, C order, a2c/utils.py:142-145) -> 512."""
    np.random.seed(1)
    p = nets.init_policy_params("cnn", (84, 84, 4), "discrete", 6)
    tp = nets.to_torch(p, torch.float64)
    rng = np.random.RandomState(2)
    obs = rng.randint(0, 256, (2, 84, 84, 4)).astype(np.uint8)
    h = obs.astype(np.float64) / 255.0
    for name, _nf, _rf, st in nets.NATURE_CONVS:
        w, b = p[f"ppo2_model/pi/{name}/w:0"], p[f"ppo2_model/pi/{name}/b:0"]
        h = np.maximum(_conv2d_nhwc_loops(h, w, st, "VALID") + b.reshape(1, 1, 1, -1), 0.0)
    assert h.shape == (2, 7, 7, 64)
    lat = np.maximum(h.reshape(2, -1) @ p["ppo2_model/pi/fc1/w:0"].astype(np.float64) + p["ppo2_model/pi/fc1/b:0"], 0.0)
    got = nets.nature_cnn(tp, "ppo2_model/pi", torch.as_tensor(obs)).numpy()
    assert np.allclose(got, lat, atol=1e-9)


def test_softmax_xent_entropy_clip_definitions():
    rng = np.random.RandomState(3)
    logits = rng.randn(50, 7) * 3
    a = rng.randint(0, 7, 50)
    # softmax_cross_entropy_with_logits_v2(labels=onehot(a)) = -sum_k onehot_k * log softmax_k
    sm = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    want = -np.log(sm[np.arange(50), a])
    got = nets.cat_neglogp(torch.tensor(logits), torch.tensor(a)).numpy()
    assert np.allclose(got, want, atol=1e-12)
    assert np.allclose(nets.cat_entropy(torch.tensor(logits)).numpy(), -(sm * np.log(sm)).sum(1), atol=1e-12)
    # tf.clip_by_global_norm: t_i * clip / max(global_norm, clip)
    gs = [rng.randn(3, 4), rng.randn(5)]
    gn = math.sqrt(sum((g ** 2).sum() for g in gs))
    for clip in (0.5, 100.0):
        out, n = nets.clip_by_global_norm([torch.tensor(g) for g in gs], clip)
        assert abs(float(n) - gn) < 1e-12
        for o, g in zip(out, gs):
            assert np.allclose(o.numpy(), g * clip / max(gn, clip), atol=1e-12)
    # tf_util.huber_loss (tf_util.py:39-45)
    x = np.linspace(-3, 3, 13)
    want = np.where(np.abs(x) < 1.0, 0.5 * x * x, np.abs(x) - 0.5)
    assert np.allclose(nets.huber(torch.tensor(x)).numpy(), want)


def test_observation_encoding_restatement():
    """common/input.py:54-57 and policies.py:182-185 with mpi_running_mean_std.py:29-30 initial statistics."""
    oh = nets.encode_observation(np.array([2, 0, 3]), torch.float32, onehot_n=4).numpy()
    assert np.array_equal(oh, np.eye(4, dtype=np.float32)[[2, 0, 3]])
    x = np.array([[-7.5, 0.25, 9.0]], np.float32)
    rms = dict(runningsum=np.zeros(3), runningsumsq=np.full(3, 1e-2), count=1e-2)     # the never-updated initial state
    assert np.array_equal(nets.encode_observation(x, torch.float32, rms=rms).numpy(), np.clip(x, -5, 5))
    rms = dict(runningsum=np.array([10.0, 0.0, -4.0]), runningsumsq=np.array([60.0, 1e-4, 40.0]), count=10.0)
    mean = np.array([1.0, 0.0, -0.4], np.float32)
    std = np.sqrt(np.maximum(np.array([6.0, 1e-5, 4.0], np.float32) - mean ** 2, 1e-2))
    assert np.allclose(nets.encode_observation(x, torch.float32, rms=rms).numpy(), np.clip((x - mean) / std, -5, 5))


# ------------------------------------------------------------------------------------------------- 2. finite differences
def _fd_check(loss_fn, tp, n_probe, rng, h=1e-6):
    """max relative error between autograd and central differences over n_probe random coordinates per tensor."""
    for t in tp.values():
        t.requires_grad_(True)
    loss = loss_fn()
    grads = torch.autograd.grad(loss, list(tp.values()), allow_unused=True)
    for t in tp.values():
        t.requires_grad_(False)
    worst = 0.0
    for (k, t), g in zip(tp.items(), grads):
        g = torch.zeros_like(t) if g is None else g
        flat = t.view(-1)
        for i in rng.choice(flat.numel(), size=min(n_probe, flat.numel()), replace=False):
            old = float(flat[i])
            flat[i] = old + h
            lp = float(loss_fn())
            flat[i] = old - h
            lm = float(loss_fn())
            flat[i] = old
            fd = (lp - lm) / (2 * h)
            an = float(g.reshape(-1)[i])
            worst = max(worst, abs(fd - an) / max(1e-6, abs(fd) + abs(an)))
    return worst


def test_ppo_loss_gradient_vs_float64_finite_differences():
    rng = np.random.RandomState(4)
    for network, ob_shape, kind, nA, vn in [("mlp", (5,), "discrete", 3, None), ("mlp", (4,), "box", 2, "copy"),
                                            ("cnn", (84, 84, 4), "discrete", 4, None)]:
        np.random.seed(5)
        p = nets.init_policy_params(network, ob_shape, kind, nA, value_network=vn)
        tp = nets.to_torch(p, torch.float64)
        B = 6 if network == "cnn" else 16
        obs = rng.randint(0, 256, (B,) + ob_shape).astype(np.uint8) if network == "cnn" else rng.randn(B, *ob_shape)
        acts = torch.tensor(rng.randint(0, nA, B)) if kind == "discrete" else torch.tensor(rng.randn(B, nA))
        advs, rets = torch.tensor(rng.randn(B)), torch.tensor(rng.randn(B))
        oldv = torch.tensor(rng.randn(B))
        with torch.no_grad():
            pi, ls, _ = nets.policy_forward(tp, network, torch.as_tensor(obs), vn)
            nlp = nets.cat_neglogp(pi, acts) if kind == "discrete" else nets.gauss_neglogp(pi, ls, acts)
        oldnlp = nlp + torch.tensor(rng.randn(B) * 0.05)
        # cliprange wide enough that no sample sits on a clip kink within +-h
        fn = lambda: nets.ppo_loss(tp, network, torch.as_tensor(obs), acts, advs, rets, oldnlp, oldv, 0.2, 0.01, 0.5, vn)[0]
        worst = _fd_check(fn, tp, 4 if network == "cnn" else 12, rng)
        assert worst < 2e-5, (network, kind, worst)


def test_dqn_loss_gradient_vs_float64_finite_differences():
    rng = np.random.RandomState(6)
    for network, ob_shape, dueling in [("mlp", (6,), True), ("mlp", (6,), False)]:
        qp = nets.init_q_params(network, ob_shape, 4, hiddens=(16,), dueling=dueling, seed=1)
        o = nets.DQNOracle(qp, network, 0.99, n_hidden=1, dueling=dueling, dtype=torch.float64)
        B = 12
        args = (rng.randn(B, *ob_shape), rng.randint(0, 4, B), rng.randn(B), rng.randn(B, *ob_shape),
                (rng.rand(B) < 0.2).astype(np.float64), rng.rand(B) + 0.1)
        fn = lambda: o.td_and_loss(*args)[1]
        assert _fd_check(fn, o.tp, 12, rng) < 2e-5


# ------------------------------------------------------------------------------------------------- 3. distribution identities
def test_distribution_identities_like_the_reference_test_probtypes():
    """common/distributions.py:321-348 (validate_probtype) on the oracle's Categorical / DiagGaussian."""
    N = 100000
    g = torch.Generator().manual_seed(0)
    np.random.seed(0)
    # DiagGaussian, pdparam of distributions.py:303
    pd = np.array([-.2, .3, .4, -.5, .1, -.5, .1, 0.8])
    mean, logstd = torch.tensor(pd[:4]).repeat(N, 1), torch.tensor(pd[4:]).repeat(N, 1)
    x = nets.gauss_sample(mean, logstd, torch.randn(N, 4, generator=g, dtype=torch.float64))
    ll = -nets.gauss_neglogp(mean, logstd, x)
    ent = float(nets.gauss_entropy(mean, logstd).mean())
    assert abs(ent + float(ll.mean())) < 3 * float(ll.std()) / math.sqrt(N)
    q = pd + np.random.randn(pd.size) * 0.1
    mean2, logstd2 = torch.tensor(q[:4]).repeat(N, 1), torch.tensor(q[4:]).repeat(N, 1)
    kl = float(nets.gauss_kl(mean, logstd, mean2, logstd2).mean())
    ll2 = -nets.gauss_neglogp(mean2, logstd2, x)
    assert abs(kl - (-ent - float(ll2.mean()))) < 3 * float(ll2.std()) / math.sqrt(N)
    assert kl >= 0
    # Categorical, pdparam of distributions.py:307
    pc = np.array([-.2, .3, .5])
    logits = torch.tensor(pc).repeat(N, 1)
    u = torch.rand(N, 3, generator=g, dtype=torch.float64).clamp_(1e-12, 1 - 1e-12)
    a = nets.cat_sample(logits, u)
    ll = -nets.cat_neglogp(logits, a)
    ent = float(nets.cat_entropy(logits).mean())
    assert abs(ent + float(ll.mean())) < 3 * float(ll.std()) / math.sqrt(N)
    q = pc + np.random.randn(pc.size) * 0.1
    logits2 = torch.tensor(q).repeat(N, 1)
    kl = float(nets.cat_kl(logits, logits2).mean())
    ll2 = -nets.cat_neglogp(logits2, a)
    assert abs(kl - (-ent - float(ll2.mean()))) < 3 * float(ll2.std()) / math.sqrt(N)
    assert kl >= 0
    # sampling frequencies follow softmax(logits) (Gumbel-max, distributions.py:199-201)
    freq = np.bincount(a.numpy(), minlength=3) / N
    sm = np.exp(pc) / np.exp(pc).sum()
    assert np.all(np.abs(freq - sm) < 4 * np.sqrt(sm * (1 - sm) / N))
