"""CPU oracle (test infrastructure): the TF1 graph of the PPO2 / DQN learner restated on
torch-CPU tensors (float32 by default, float64 for finite-difference checks).

PARITY UNPINNED at the TensorFlow boundary: TF 1.x is absent (see oracle/__init__.py).
Every function cites the reference lines it follows.  Anchors that do not go through this
module's own torch code (tests/test_oracle_anchors.py): the convolution / matmul / softmax
cross-entropy primitives are compared with definition-level numpy loops of the TF op
semantics, autograd gradients with float64 central finite differences of the loss, and the
distributions with the statistical identities the reference itself tests
(common/distributions.py:299-348, entropy = -E[log p], KL = -H - E_p[log q], 3 sigma).  Parameters live in an ordered
``dict name -> np.ndarray(float32)`` using the reference's TF variable names and
layouts (conv HWIO ``[rf, rf, nin, nf]`` a2c/utils.py:50; fc ``[nin, nh]`` :61; conv bias
``[1, nf, 1, 1]`` :49), so reference checkpoints (tf_util.py:345-355) map 1:1.

Gradients come from torch.autograd -- allowed for the oracle only; the product has
hand-written CUDA backward kernels.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------
# initialisation  (a2c/utils.py:20-35 ortho_init; creation order = policies.py:121-179)
# ----------------------------------------------------------------------------------
def ortho_init_np(shape, scale, rng=np.random):
    """a2c/utils.py:20-35.  Draws from the GLOBAL numpy RandomState by default, exactly
    like the reference (seeded by set_global_seeds, misc_util.py:48-60)."""
    shape = tuple(shape)
    if len(shape) == 2:
        flat_shape = shape
    elif len(shape) == 4:
        flat_shape = (int(np.prod(shape[:-1])), shape[-1])
    else:
        raise NotImplementedError
    a = rng.normal(0.0, 1.0, flat_shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == flat_shape else v
    q = q.reshape(shape)
    return (scale * q[:shape[0], :shape[1]]).astype(np.float32)


NATURE_CONVS = [("c1", 32, 8, 4), ("c2", 64, 4, 2), ("c3", 64, 3, 1)]   # models.py:21-24


def _conv_out(h, rf, stride):
    return (h - rf) // stride + 1


def _init_network(params, prefix, network, ob_shape, rng, num_layers=2, num_hidden=64):
    """Creates the variables of one latent network under ``prefix`` in the reference's
    creation order; returns the latent width."""
    if network == "cnn":
        h, w, c = ob_shape
        nin = c
        for name, nf, rf, stride in NATURE_CONVS:
            params[f"{prefix}/{name}/w:0"] = ortho_init_np((rf, rf, nin, nf), math.sqrt(2), rng)
            params[f"{prefix}/{name}/b:0"] = np.zeros((1, nf, 1, 1), np.float32)
            h, w, nin = _conv_out(h, rf, stride), _conv_out(w, rf, stride), nf
        flat = h * w * nin
        params[f"{prefix}/fc1/w:0"] = ortho_init_np((flat, 512), math.sqrt(2), rng)
        params[f"{prefix}/fc1/b:0"] = np.zeros((512,), np.float32)
        return 512
    if network == "mlp":
        nin = int(np.prod(ob_shape))
        for i in range(num_layers):                                   # models.py:94-99
            params[f"{prefix}/mlp_fc{i}/w:0"] = ortho_init_np((nin, num_hidden), math.sqrt(2), rng)
            params[f"{prefix}/mlp_fc{i}/b:0"] = np.zeros((num_hidden,), np.float32)
            nin = num_hidden
        return nin
    raise ValueError(network)


def init_policy_params(network, ob_shape, ac_kind, ac_dim, value_network=None, seed=None,
                       scope="ppo2_model", **net_kwargs):
    """Variable creation order of ``policy_fn`` (policies.py:126-177) + PolicyWithValue
    (:41-64): pi-network, [vf-network if value_network='copy'], pi head (init_scale 0.01,
    :49), [pi/logstd], vf head.  ac_kind in {'discrete', 'box'}."""
    rng = np.random
    if seed is not None:
        rng = np.random.RandomState(seed)
    p = OrderedDict()
    nlat = _init_network(p, f"{scope}/pi", network, ob_shape, rng, **net_kwargs)
    nlat_v = nlat
    if value_network == "copy":
        nlat_v = _init_network(p, f"{scope}/vf", network, ob_shape, rng, **net_kwargs)
    if nlat != ac_dim:                                                # _matching_fc distributions.py:351-355
        p[f"{scope}/pi/w:0"] = ortho_init_np((nlat, ac_dim), 0.01, rng)
        p[f"{scope}/pi/b:0"] = np.zeros((ac_dim,), np.float32)
    if ac_kind == "box":
        p[f"{scope}/pi/logstd:0"] = np.zeros((1, ac_dim), np.float32)  # distributions.py:104
    p[f"{scope}/vf/w:0"] = ortho_init_np((nlat_v, 1), 1.0, rng)         # policies.py:63 (fc default scale 1.0)
    p[f"{scope}/vf/b:0"] = np.zeros((1,), np.float32)
    return p


# ----------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------
def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def _conv_nhwc(x, w_hwio, b, stride, pad="VALID"):
    """a2c/utils.py:37-56 (tf.nn.conv2d NHWC + bias)."""
    xin = x.permute(0, 3, 1, 2)
    w = w_hwio.permute(3, 2, 0, 1)
    if pad == "VALID":
        y = F.conv2d(xin, w, stride=stride)
    else:  # TF 'SAME'
        ih, iw = xin.shape[2], xin.shape[3]
        rf = w.shape[2]
        oh, ow = -(-ih // stride), -(-iw // stride)
        ph = max((oh - 1) * stride + rf - ih, 0)
        pw = max((ow - 1) * stride + rf - iw, 0)
        xin = F.pad(xin, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        y = F.conv2d(xin, w, stride=stride)
    y = y + b.reshape(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1)


def nature_cnn(tp, prefix, obs):
    """models.py:15-26.  obs: [B,84,84,4] uint8 (or float) tensor."""
    dtype = tp[f"{prefix}/c1/w:0"].dtype
    h = obs.to(dtype) / 255.0                                         # models.py:19
    for name, _nf, _rf, stride in NATURE_CONVS:
        h = torch.relu(_conv_nhwc(h, tp[f"{prefix}/{name}/w:0"], tp[f"{prefix}/{name}/b:0"], stride))
    h = h.reshape(h.shape[0], -1)                                     # conv_to_fc a2c/utils.py:142-145 (H,W,C order)
    return torch.relu(h @ tp[f"{prefix}/fc1/w:0"] + tp[f"{prefix}/fc1/b:0"])


def encode_observation(obs, dtype, onehot_n=0, rms=None, clip=(-5.0, 5.0)):
    """common/input.py:43-63 (Discrete -> to_float(one_hot), Box -> to_float) preceded, for float Box
    observations with normalize_observations, by policies.py:182-185
    clip_by_value((x - rms.mean) / rms.std, -5, 5) with the RunningMeanStd of mpi_running_mean_std.py:29-30.
    rms: dict(runningsum, runningsumsq, count) float64 or None."""
    obs = torch.as_tensor(obs)
    if onehot_n:
        return F.one_hot(obs.long().reshape(-1), onehot_n).to(dtype)
    x = obs
    if rms is not None:
        mean = torch.as_tensor(np.asarray(rms["runningsum"]) / rms["count"]).float()
        var = torch.as_tensor(np.asarray(rms["runningsumsq"]) / rms["count"]).float() - mean ** 2
        std = torch.sqrt(torch.clamp(var, min=1e-2))
        x = torch.clamp((x.float() - mean) / std, clip[0], clip[1])
    return x.to(dtype)


def mlp(tp, prefix, obs, num_layers=2):
    """models.py:74-103 (tanh, no layer_norm)."""
    dtype = tp[f"{prefix}/mlp_fc0/w:0"].dtype
    h = obs.to(dtype).reshape(obs.shape[0], -1)                       # input.py:56-57 to_float; flatten
    for i in range(num_layers):
        h = torch.tanh(h @ tp[f"{prefix}/mlp_fc{i}/w:0"] + tp[f"{prefix}/mlp_fc{i}/b:0"])
    return h


def latent_fn(network):
    return {"cnn": nature_cnn, "mlp": mlp}[network]


def policy_forward(tp, network, obs, value_network=None, scope="ppo2_model"):
    """policies.py:41-64.  Returns (pi [B,nA or d], logstd or None, vf [B])."""
    lat = latent_fn(network)(tp, f"{scope}/pi", obs)
    vlat = latent_fn(network)(tp, f"{scope}/vf", obs) if value_network == "copy" else lat
    if f"{scope}/pi/w:0" in tp:
        pi = lat @ tp[f"{scope}/pi/w:0"] + tp[f"{scope}/pi/b:0"]
    else:
        pi = lat
    vf = (vlat @ tp[f"{scope}/vf/w:0"] + tp[f"{scope}/vf/b:0"])[:, 0]
    logstd = tp.get(f"{scope}/pi/logstd:0")
    return pi, logstd, vf


# ----------------------------------------------------------------------------------
# distributions
# ----------------------------------------------------------------------------------
def cat_neglogp(logits, actions):
    """distributions.py:164-183: softmax_cross_entropy_with_logits_v2(logits, onehot(a))."""
    lse = torch.logsumexp(logits, dim=-1)
    return lse - logits.gather(1, actions.long().view(-1, 1))[:, 0]


def cat_entropy(logits):
    """distributions.py:193-198."""
    a0 = logits - logits.max(dim=-1, keepdim=True).values
    ea0 = torch.exp(a0)
    z0 = ea0.sum(dim=-1, keepdim=True)
    p0 = ea0 / z0
    return (p0 * (torch.log(z0) - a0)).sum(dim=-1)


def cat_kl(logits, other):
    """distributions.py:185-192."""
    a0 = logits - logits.max(dim=-1, keepdim=True).values
    a1 = other - other.max(dim=-1, keepdim=True).values
    ea0, ea1 = torch.exp(a0), torch.exp(a1)
    z0, z1 = ea0.sum(dim=-1, keepdim=True), ea1.sum(dim=-1, keepdim=True)
    p0 = ea0 / z0
    return (p0 * (a0 - torch.log(z0) - a1 + torch.log(z1))).sum(dim=-1)


def cat_sample(logits, uniforms):
    """distributions.py:199-201 with the uniform noise injected."""
    return torch.argmax(logits - torch.log(-torch.log(uniforms)), dim=-1)


def gauss_neglogp(mean, logstd, x):
    """distributions.py:238-241."""
    std = torch.exp(logstd)
    d = x.shape[-1]
    return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * d \
        + (mean * 0.0 + logstd).sum(dim=-1)


def gauss_entropy(mean, logstd):
    """distributions.py:245-246."""
    return (mean * 0.0 + logstd + 0.5 * math.log(2.0 * math.pi * math.e)).sum(dim=-1)


def gauss_kl(mean, logstd, mean2, logstd2):
    """distributions.py:242-244."""
    std, std2 = torch.exp(logstd), torch.exp(logstd2)
    return (logstd2 - logstd + (std ** 2 + (mean - mean2) ** 2) / (2.0 * std2 ** 2) - 0.5).sum(dim=-1)


def gauss_sample(mean, logstd, normals):
    """distributions.py:247-248 with the normal noise injected."""
    return mean + torch.exp(logstd) * normals


def policy_step(params, network, obs, noise, value_network=None, dtype=torch.float32):
    """PolicyWithValue.step (policies.py:77-96): (actions, values, neglogp)."""
    tp = to_torch(params, dtype)
    with torch.no_grad():
        pi, logstd, vf = policy_forward(tp, network, torch.as_tensor(obs), value_network)
        noise = torch.as_tensor(noise, dtype=dtype)
        if logstd is None:
            a = cat_sample(pi, noise)
            nlp = cat_neglogp(pi, a)
        else:
            a = gauss_sample(pi, logstd, noise)
            nlp = gauss_neglogp(pi, logstd, a)
    return a.numpy(), vf.numpy(), nlp.numpy(), pi.numpy()


# ----------------------------------------------------------------------------------
# PPO2 loss / train step
# ----------------------------------------------------------------------------------
LOSS_NAMES = ["policy_loss", "value_loss", "policy_entropy", "approxkl", "clipfrac"]   # ppo2/model.py:115


def ppo_loss(tp, network, obs, actions, advs, returns, oldneglogp, oldvpred, cliprange,
             ent_coef, vf_coef, value_network=None):
    """ppo2/model.py:57-91."""
    pi, logstd, vpred = policy_forward(tp, network, obs, value_network)
    if logstd is None:
        neglogpac = cat_neglogp(pi, actions)
        entropy = cat_entropy(pi).mean()
    else:
        neglogpac = gauss_neglogp(pi, logstd, actions)
        entropy = gauss_entropy(pi, logstd).mean()
    vpredclipped = oldvpred + torch.clamp(vpred - oldvpred, -cliprange, cliprange)
    vf_losses1 = (vpred - returns) ** 2
    vf_losses2 = (vpredclipped - returns) ** 2
    vf_loss = 0.5 * torch.maximum(vf_losses1, vf_losses2).mean()
    ratio = torch.exp(oldneglogp - neglogpac)
    pg_losses = -advs * ratio
    pg_losses2 = -advs * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    pg_loss = torch.maximum(pg_losses, pg_losses2).mean()
    approxkl = 0.5 * ((neglogpac - oldneglogp) ** 2).mean()
    clipfrac = ((ratio - 1.0).abs() > cliprange).to(ratio.dtype).mean()
    loss = pg_loss - entropy * ent_coef + vf_loss * vf_coef
    return loss, [pg_loss, vf_loss, entropy, approxkl, clipfrac]


def normalize_advs(returns, values):
    """ppo2/model.py:136-139 (numpy float32, population std, +1e-8)."""
    advs = np.asarray(returns, np.float32) - np.asarray(values, np.float32)
    return (advs - advs.mean()) / (advs.std() + 1e-8)


def clip_by_global_norm(grads, clip_norm):
    """tf.clip_by_global_norm (ppo2/model.py:105-107): g * clip / max(||g||, clip)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    scale = clip_norm / torch.maximum(gn, torch.tensor(clip_norm, dtype=gn.dtype))
    return [g * scale for g in grads], gn


def adam_tf(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-5):
    """TF-Adam as pinned by baselines/common/mpi_adam.py:37-42 (numpy statement that the
    reference's own test :64-99 proves equal to tf.train.AdamOptimizer):
    a = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; p += -a*m/(sqrt(v)+eps)."""
    a = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * (g * g)
    p = p + (-a) * m / (torch.sqrt(v) + eps)
    return p, m, v


class PPO2Oracle:
    """Restates ppo2/model.py Model (train/step/value) on CPU."""

    def __init__(self, params, network, ent_coef, vf_coef, max_grad_norm, value_network=None,
                 dtype=torch.float32, adam_eps=1e-5):
        self.network = network
        self.value_network = value_network
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.dtype = dtype
        self.adam_eps = adam_eps
        self.tp = to_torch(params, dtype)
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.tp.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.tp.items())
        self.t = 0
        self.last_grads = None
        self.last_gnorm = None

    def params_np(self):
        return OrderedDict((k, v.detach().numpy().copy()) for k, v in self.tp.items())

    def grads(self, cliprange, obs, returns, actions, values, neglogpacs, advs=None):
        dt = self.dtype
        if advs is None:
            advs = normalize_advs(returns, values)
        for t in self.tp.values():
            t.requires_grad_(True)
            t.grad = None
        act_t = torch.as_tensor(actions)
        if act_t.dtype.is_floating_point:
            act_t = act_t.to(dt)
        loss, stats = ppo_loss(self.tp, self.network, torch.as_tensor(obs), act_t,
                               torch.as_tensor(advs, dtype=dt), torch.as_tensor(returns, dtype=dt),
                               torch.as_tensor(neglogpacs, dtype=dt), torch.as_tensor(values, dtype=dt),
                               cliprange, self.ent_coef, self.vf_coef, self.value_network)
        grads = torch.autograd.grad(loss, list(self.tp.values()), allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, self.tp.values())]
        for t in self.tp.values():
            t.requires_grad_(False)
        return [float(s) for s in stats], grads

    def train(self, lr, cliprange, obs, returns, masks, actions, values, neglogpacs, states=None,
              grad_transform=None):
        """ppo2/model.py:133-158.  grad_transform (optional) models the MPI mean of
        mpi_adam_optimizer.py:39-40 applied BEFORE the global-norm clip."""
        stats, grads = self.grads(cliprange, obs, returns, actions, values, neglogpacs)
        if grad_transform is not None:
            grads = grad_transform(grads)
        self.last_grads = OrderedDict((k, g.numpy().copy()) for k, g in zip(self.tp.keys(), grads))
        if self.max_grad_norm is not None:
            grads, gn = clip_by_global_norm(grads, self.max_grad_norm)
            self.last_gnorm = float(gn)
        self.t += 1
        for (k, p), g in zip(list(self.tp.items()), grads):
            self.tp[k], self.m[k], self.v[k] = adam_tf(p, g, self.m[k], self.v[k], self.t, lr,
                                                       eps=self.adam_eps)
        return stats

    def train_microbatched(self, lr, cliprange, obs, returns, masks, actions, values, neglogpacs, microbatch_size):
        """ppo2/microbatched_model.py:35-75: normalise over the full minibatch (:43), per microbatch take the
        gradients AFTER clip_by_global_norm (`self.grads`, ppo2/model.py:105-108), average them (:70), one
        apply_gradients; statistics are the mean over microbatches (:75)."""
        advs = normalize_advs(returns, values)
        n = len(returns)
        assert n % microbatch_size == 0
        nmicro = n // microbatch_size
        total, stats_all = None, []
        for i in range(nmicro):
            sl = slice(i * microbatch_size, (i + 1) * microbatch_size)
            st, grads = self.grads(cliprange, obs[sl], returns[sl], actions[sl], values[sl], neglogpacs[sl],
                                   advs=advs[sl])
            if self.max_grad_norm is not None:
                grads, _ = clip_by_global_norm(grads, self.max_grad_norm)
            total = grads if total is None else [a + g for a, g in zip(total, grads)]
            stats_all.append(st)
        grads = [g / nmicro for g in total]
        self.last_grads = OrderedDict((k, g.numpy().copy()) for k, g in zip(self.tp.keys(), grads))
        self.t += 1
        for (k, p), g in zip(list(self.tp.items()), grads):
            self.tp[k], self.m[k], self.v[k] = adam_tf(p, g, self.m[k], self.v[k], self.t, lr, eps=self.adam_eps)
        return np.mean(np.array(stats_all), axis=0).tolist()

    def step(self, obs, noise):
        return policy_step(self.params_np(), self.network, obs, noise, self.value_network, self.dtype)

    def value(self, obs):
        with torch.no_grad():
            return policy_forward(self.tp, self.network, torch.as_tensor(obs), self.value_network)[2].numpy()


# ----------------------------------------------------------------------------------
# deepq
# ----------------------------------------------------------------------------------
def xavier_uniform_np(shape, rng):
    """tf.contrib.layers default weights_initializer (xavier, uniform=True):
    U(-sqrt(6/(fan_in+fan_out)), +...) with fan computed over receptive field."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_q_params(network, ob_shape, num_actions, hiddens=(256,), dueling=True, seed=0,
                  scope="deepq/q_func"):
    """Variables of build_q_func (deepq/models.py:5-45) for network in {'cnn','conv_only'}.
    'cnn' = nature_cnn (ortho init, VALID); 'conv_only' = contrib convolution2d (xavier,
    SAME, models.py:221-249).  Heads are contrib fully_connected (xavier, zero bias)."""
    rng = np.random.RandomState(seed)
    p = OrderedDict()
    if network == "cnn":
        nlat = _init_network(p, scope, "cnn", ob_shape, rng)
    elif network == "conv_only":
        h, w, c = ob_shape
        nin = c
        for i, (_n, nf, rf, stride) in enumerate(NATURE_CONVS):
            nm = "Conv" if i == 0 else f"Conv_{i}"
            p[f"{scope}/convnet/{nm}/weights:0"] = xavier_uniform_np((rf, rf, nin, nf), rng)
            p[f"{scope}/convnet/{nm}/biases:0"] = np.zeros((nf,), np.float32)
            h, w, nin = -(-h // stride), -(-w // stride), nf
        nlat = h * w * nin
    elif network == "mlp":
        nlat = _init_network(p, scope, "mlp", ob_shape, rng)
    else:
        raise ValueError(network)
    streams = [("action_value", num_actions)] + ([("state_value", 1)] if dueling else [])
    for sname, nout in streams:
        nin = nlat
        for j, hsz in enumerate(hiddens):
            nm = "fully_connected" if j == 0 else f"fully_connected_{j}"
            p[f"{scope}/{sname}/{nm}/weights:0"] = xavier_uniform_np((nin, hsz), rng)
            p[f"{scope}/{sname}/{nm}/biases:0"] = np.zeros((hsz,), np.float32)
            nin = hsz
        j = len(hiddens)
        nm = "fully_connected" if j == 0 else f"fully_connected_{j}"
        p[f"{scope}/{sname}/{nm}/weights:0"] = xavier_uniform_np((nin, nout), rng)
        p[f"{scope}/{sname}/{nm}/biases:0"] = np.zeros((nout,), np.float32)
    return p


def _fc_name(j):
    return "fully_connected" if j == 0 else f"fully_connected_{j}"


def q_forward(tp, network, obs, scope, n_hidden=1, dueling=True):
    """deepq/models.py:10-43."""
    if network == "cnn":
        lat = nature_cnn(tp, scope, obs)
    elif network == "mlp":
        lat = mlp(tp, scope, obs)
    else:
        dtype = tp[f"{scope}/convnet/Conv/weights:0"].dtype
        h = obs.to(dtype) / 255.0
        for i, (_n, _nf, _rf, stride) in enumerate(NATURE_CONVS):
            nm = "Conv" if i == 0 else f"Conv_{i}"
            h = torch.relu(_conv_nhwc(h, tp[f"{scope}/convnet/{nm}/weights:0"],
                                      tp[f"{scope}/convnet/{nm}/biases:0"], stride, pad="SAME"))
        lat = h.reshape(h.shape[0], -1)

    def stream(sname):
        x = lat
        for j in range(n_hidden):
            x = torch.relu(x @ tp[f"{scope}/{sname}/{_fc_name(j)}/weights:0"]
                           + tp[f"{scope}/{sname}/{_fc_name(j)}/biases:0"])
        return x @ tp[f"{scope}/{sname}/{_fc_name(n_hidden)}/weights:0"] \
            + tp[f"{scope}/{sname}/{_fc_name(n_hidden)}/biases:0"]

    a = stream("action_value")
    if not dueling:
        return a
    s = stream("state_value")
    return s + (a - a.mean(dim=1, keepdim=True))


def huber(x, delta=1.0):
    """tf_util.py:39-45."""
    return torch.where(x.abs() < delta, 0.5 * x * x, delta * (x.abs() - 0.5 * delta))


class DQNOracle:
    """Restates deepq/build_graph.py:380-444 train / update_target with Adam(eps=1e-8)."""

    def __init__(self, q_params, network, gamma, n_hidden=1, dueling=True, double_q=True,
                 grad_norm_clipping=10.0, dtype=torch.float32, scope="deepq/q_func",
                 tscope="deepq/target_q_func"):
        self.network, self.gamma, self.n_hidden, self.dueling = network, gamma, n_hidden, dueling
        self.double_q, self.clip = double_q, grad_norm_clipping
        self.scope, self.tscope, self.dtype = scope, tscope, dtype
        self.tp = to_torch(q_params, dtype)
        self.tt = OrderedDict((k.replace(scope, tscope, 1), v.clone()) for k, v in self.tp.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.tp.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.tp.items())
        self.t = 0
        self.last_grads = None

    def q_values(self, obs):
        with torch.no_grad():
            return q_forward(self.tp, self.network, torch.as_tensor(obs), self.scope, self.n_hidden,
                             self.dueling).numpy()

    def td_and_loss(self, obs_t, actions, rewards, obs_tp1, dones, weights):
        dt = self.dtype
        q_t = q_forward(self.tp, self.network, torch.as_tensor(obs_t), self.scope, self.n_hidden, self.dueling)
        with torch.no_grad():
            q_tp1 = q_forward(self.tt, self.network, torch.as_tensor(obs_tp1), self.tscope, self.n_hidden,
                              self.dueling)
            if self.double_q:                                         # build_graph.py:399-402
                q_on = q_forward(self.tp, self.network, torch.as_tensor(obs_tp1), self.scope, self.n_hidden,
                                 self.dueling)
                best = q_tp1.gather(1, q_on.argmax(dim=1, keepdim=True))[:, 0]
            else:
                best = q_tp1.max(dim=1).values
            target = torch.as_tensor(rewards, dtype=dt) + self.gamma * (1.0 - torch.as_tensor(dones, dtype=dt)) * best
        q_sel = q_t.gather(1, torch.as_tensor(actions).long().view(-1, 1))[:, 0]
        td = q_sel - target
        loss = (torch.as_tensor(weights, dtype=dt) * huber(td)).mean()
        return td, loss

    def train(self, lr, obs_t, actions, rewards, obs_tp1, dones, weights, adam_eps=1e-8):
        for t in self.tp.values():
            t.requires_grad_(True)
        td, loss = self.td_and_loss(obs_t, actions, rewards, obs_tp1, dones, weights)
        grads = torch.autograd.grad(loss, list(self.tp.values()), allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, self.tp.values())]
        for t in self.tp.values():
            t.requires_grad_(False)
        self.last_grads = OrderedDict((k, g.numpy().copy()) for k, g in zip(self.tp.keys(), grads))   # pre-clip
        if self.clip is not None:                                     # per-variable tf.clip_by_norm :416-421
            out = []
            for g in grads:
                n = torch.sqrt((g.double() ** 2).sum()).to(g.dtype)
                out.append(g * self.clip / torch.maximum(n, torch.tensor(self.clip, dtype=g.dtype)))
            grads = out
        self.t += 1
        for (k, p), g in zip(list(self.tp.items()), grads):
            self.tp[k], self.m[k], self.v[k] = adam_tf(p, g, self.m[k], self.v[k], self.t, lr, eps=adam_eps)
        return td.detach().numpy()

    def update_target(self):
        self.tt = OrderedDict((k.replace(self.scope, self.tscope, 1), v.clone()) for k, v in self.tp.items())
