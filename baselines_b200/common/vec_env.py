"""VecEnv boundary (host side).  Env stepping stays on host CPU (north star); this module only restates the
interface the learner consumes -- reference: baselines/common/vec_env/vec_env.py:29-138 (VecEnv ABC),
dummy_vec_env.py:5-81 (auto-reset on done, obs buffers), plus synthetic envs for benchmarks.

Any object with the same attributes works with Runner (the reference's own SubprocVecEnv / DummyVecEnv do).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import spaces


class VecEnv(ABC):
    """vec_env.py:29-138: batched env; reset() -> obs[N,...]; step(a[N]) -> (obs, rews f32[N], dones bool[N], infos)."""
    closed = False

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.closed = True


class DummyVecEnv(VecEnv):
    """Sequential in-process VecEnv (dummy_vec_env.py:5-81): envs are created from thunks, stepped in a loop,
    and reset automatically when done (the returned obs is then the first obs of the next episode)."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        env = self.envs[0]
        super().__init__(len(self.envs), env.observation_space, env.action_space)
        shp, dt = tuple(env.observation_space.shape), env.observation_space.dtype
        self.buf_obs = np.zeros((self.num_envs,) + shp, dtype=dt)
        self.buf_dones = np.zeros((self.num_envs,), dtype=np.bool_)
        self.buf_rews = np.zeros((self.num_envs,), dtype=np.float32)
        self.buf_infos = [{} for _ in range(self.num_envs)]
        self.actions = None

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        for e in range(self.num_envs):
            a = self.actions[e]
            obs, self.buf_rews[e], self.buf_dones[e], self.buf_infos[e] = self.envs[e].step(a)
            if self.buf_dones[e]:
                obs = self.envs[e].reset()
            self.buf_obs[e] = obs
        return self.buf_obs.copy(), self.buf_rews.copy(), self.buf_dones.copy(), list(self.buf_infos)

    def reset(self):
        for e in range(self.num_envs):
            self.buf_obs[e] = self.envs[e].reset()
        return self.buf_obs.copy()


class VecEnvWrapper(VecEnv):
    """vec_env.py:140-175: a wrapper over a whole batch of envs; unknown public attributes fall through to venv."""

    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(venv.num_envs, observation_space or venv.observation_space,
                         action_space or venv.action_space)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def close(self):
        return self.venv.close()

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.venv, name)


class VecFrameStack(VecEnvWrapper):
    """vec_frame_stack.py:6-31: stack the last `nstack` observations along the channel axis; the stack of an env is
    cleared when its episode ends.

    reset()/step_wait() are the reference's host implementation (np.roll on a [N, H, W, nstack*c] array).  The
    B200 Runner does not call them: it sees `frame_stack_device`, pulls the UNSTACKED frames with step_frames()
    (1/nstack of the bytes over PCIe) and applies the same update to the HBM-resident rollout buffer with
    b200rl_frame_stack -- the previous stacked observation is already there as rollout.obs[t-1]."""
    frame_stack_device = True

    def __init__(self, venv, nstack):
        self.nstack = nstack
        wos = venv.observation_space
        low = np.repeat(wos.low, nstack, axis=-1)
        high = np.repeat(wos.high, nstack, axis=-1)
        self.frame_channels = int(wos.shape[-1])
        self.stackedobs = np.zeros((venv.num_envs,) + low.shape, low.dtype)
        ob_space = spaces.Box(low=low, high=high, dtype=wos.dtype)
        super().__init__(venv, observation_space=ob_space)

    def step_wait(self):
        obs, rews, news, infos = self.venv.step_wait()
        self.stackedobs = np.roll(self.stackedobs, shift=-1, axis=-1)
        self.stackedobs[np.asarray(news, dtype=np.bool_)] = 0
        self.stackedobs[..., -obs.shape[-1]:] = obs
        return self.stackedobs, rews, news, infos

    def reset(self):
        obs = self.venv.reset()
        self.stackedobs[...] = 0
        self.stackedobs[..., -obs.shape[-1]:] = obs
        return self.stackedobs

    # ---- unstacked access for the device-side stack
    def reset_frames(self):
        return self.venv.reset()

    def step_frames(self, actions):
        """(new frames [N, ..., c], rews, news, infos): the wrapped env's step; stacking is left to the caller."""
        self.venv.step_async(actions)
        return self.venv.step_wait()


class EpisodeStats:
    """What bench.Monitor contributes to the learner (bench/monitor.py:58-75): info['episode'] = {r, l, t}."""

    def __init__(self, env):
        import time
        self.env, self._t0, self._time = env, time.time(), time
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.r, self.l = 0.0, 0

    def reset(self):
        self.r, self.l = 0.0, 0
        return self.env.reset()

    def step(self, a):
        ob, rew, done, info = self.env.step(a)
        self.r += float(rew)
        self.l += 1
        if done:
            info = dict(info)
            info['episode'] = {"r": round(self.r, 6), "l": self.l, "t": round(self._time.time() - self._t0, 6)}
        return ob, rew, done, info


class SyntheticVecEnv(VecEnv):
    """Zero-cost host VecEnv for throughput measurement: observations cycle through a pool of
    pre-generated batches held in PINNED host memory, rewards ~ N(0,1), dones ~ Bernoulli(p_done)."""

    def __init__(self, num_envs, ob_shape=(84, 84, 4), ob_dtype=np.uint8, n_actions=6, act_dim=None, pool=8,
                 p_done=0.01, seed=0):
        ob_space = spaces.Box(0, 255, ob_shape, ob_dtype) if np.dtype(ob_dtype) == np.uint8 else \
            spaces.Box(-10.0, 10.0, ob_shape, ob_dtype)
        ac_space = spaces.Discrete(n_actions) if act_dim is None else spaces.Box(-1.0, 1.0, (act_dim,), np.float32)
        super().__init__(num_envs, ob_space, ac_space)
        rng = np.random.RandomState(seed)
        self.pool = []
        for _ in range(pool):
            if np.dtype(ob_dtype) == np.uint8:
                a = rng.randint(0, 256, size=(num_envs,) + tuple(ob_shape), dtype=np.uint8)
            else:
                a = np.clip(rng.randn(num_envs, *ob_shape), -10, 10).astype(ob_dtype)    # vec_normalize.py:10,39 clip
            t = torch.from_numpy(a)
            if torch.cuda.is_available():
                t = t.pin_memory()
            self.pool.append(t)
        self.rews = rng.randn(64, num_envs).astype(np.float32)
        self.dones = rng.rand(64, num_envs) < p_done
        self.t = 0

    def reset(self):
        self.t = 0
        return self.pool[0].numpy()

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        self.t += 1
        return (self.pool[self.t % len(self.pool)].numpy(), self.rews[self.t % 64], self.dones[self.t % 64],
                _EMPTY_INFOS[:self.num_envs] if self.num_envs <= len(_EMPTY_INFOS) else [{}] * self.num_envs)


_EMPTY_INFOS = [{} for _ in range(65536)]


class DeviceSyntheticVecEnv(VecEnv):
    """Same synthetic process with every array resident in HBM (inputs already on device when the timed
    region starts: bench.py's `value`).  Exposes reset_device / step_device, which Runner detects."""

    def __init__(self, num_envs, ob_shape=(84, 84, 4), ob_dtype=np.uint8, n_actions=6, act_dim=None, pool=8,
                 p_done=0.01, seed=0, device=None):
        host = SyntheticVecEnv(num_envs, ob_shape, ob_dtype, n_actions, act_dim, pool, p_done, seed)
        super().__init__(num_envs, host.observation_space, host.action_space)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.pool = [p.to(self.device) for p in host.pool]
        self.rews = torch.from_numpy(host.rews).to(self.device)
        self.dones = torch.from_numpy(host.dones.astype(np.uint8)).to(self.device)
        self.t = 0

    def reset_device(self):
        self.t = 0
        return self.pool[0]

    def step_device(self, actions):
        self.t += 1
        return self.pool[self.t % len(self.pool)], self.rews[self.t % 64], self.dones[self.t % 64]

    def reset(self):
        return self.reset_device().cpu().numpy()

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        o, r, d = self.step_device(None)
        return o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy().astype(np.bool_), [{}] * self.num_envs
