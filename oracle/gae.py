"""CPU oracle (test infrastructure): GAE(lambda) backward recurrence.

Restates ``baselines/ppo2/runner.py:53-65`` (reference) including its dtype
behaviour, which decides the last bit of every output:

* ``nextnonterminal = 1.0 - dones`` is float64 (bool array promoted by the python
  float), runner.py:58,61;
* ``self.gamma * nextvalues`` is a python float times a float32 array -> float32
  product (numpy scalar casting), THEN multiplied by the float64 mask;
* ``delta`` and ``lastgaelam`` are therefore float64 and only the store into
  ``mb_advs[t]`` (float32) rounds, runner.py:64;
* ``mb_returns = mb_advs + mb_values`` is a float32 add, runner.py:65.
"""
import numpy as np


def gae_reference_order(rewards, values, dones, last_values, last_dones, gamma, lam):
    """rewards, values: float32 [T, N]; dones: bool [T, N] (done BEFORE step t, i.e. the
    reference's ``mb_dones``); last_values float32 [N]; last_dones bool [N] (``self.dones``
    after the last env.step).  Returns (advs, returns) float32 [T, N]."""
    rewards = np.asarray(rewards, dtype=np.float32)
    values = np.asarray(values, dtype=np.float32)
    dones = np.asarray(dones, dtype=np.bool_)
    last_values = np.asarray(last_values, dtype=np.float32)
    last_dones = np.asarray(last_dones, dtype=np.bool_)
    T = rewards.shape[0]
    advs = np.zeros_like(rewards)
    lastgaelam = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nextnonterminal = 1.0 - last_dones          # float64
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - dones[t + 1]        # float64
            nextvalues = values[t + 1]
        # float32 product first (python float is a weak scalar), then float64
        delta = rewards[t] + (np.float32(gamma) * nextvalues) * nextnonterminal - values[t]
        advs[t] = lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam
    returns = advs + values
    return advs, returns


def sf01(arr):
    """runner.py:69-74: swap axes 0,1 and flatten -> env-major flat index i = e*T + t."""
    s = arr.shape
    return arr.swapaxes(0, 1).reshape(s[0] * s[1], *s[2:])
