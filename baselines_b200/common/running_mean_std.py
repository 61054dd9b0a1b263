"""Streaming mean / variance (host side, float64) -- restates the numpy class of the reference's
baselines/common/running_mean_std.py:5-35 (Chan et al. parallel-variance combination).  Pinned, through VecNormalize,
by tests/golden/vec_normalize_trace.npz (outputs of the reference class)."""
import numpy as np


def combine_moments(mean, var, count, b_mean, b_var, b_count):
    """running_mean_std.py:22-34: merge (mean, var, count) with a batch's moments; same operation order."""
    delta = b_mean - mean
    tot = count + b_count
    new_mean = mean + delta * b_count / tot
    m2 = var * count + b_var * b_count + np.square(delta) * count * b_count / tot
    return new_mean, m2 / tot, tot


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, b_mean, b_var, b_count):
        self.mean, self.var, self.count = combine_moments(self.mean, self.var, self.count, b_mean, b_var, b_count)
