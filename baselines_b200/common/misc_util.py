"""Host helpers used by learn(): reference common/misc_util.py:48-60 (set_global_seeds) and
common/math_util.py:25-42 (explained_variance)."""
import random

import numpy as np


def set_global_seeds(i):
    """misc_util.py:48-60.  The reference's rank offset is always 0 (its `import MPI` fails), so
    every rank seeds identically; kept."""
    if i is None:
        return
    myseed = int(i)
    np.random.seed(myseed)
    random.seed(myseed)
    try:
        import torch
        torch.manual_seed(myseed)
    except ImportError:
        pass


def explained_variance(ypred, y):
    """math_util.py:25-42: 1 - Var[y - ypred] / Var[y]; nan if Var[y] == 0."""
    assert y.ndim == 1 and ypred.ndim == 1
    vary = np.var(y)
    return np.nan if vary == 0 else 1 - np.var(y - ypred) / vary


def safemean(xs):
    return np.nan if len(xs) == 0 else np.mean(xs)


def constfn(val):
    def f(_):
        return val
    return f
