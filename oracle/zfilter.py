"""Oracle (test infrastructure): Welford running statistics and the ZFilter.

Restates running_stat.py:4-33 and filters.py:17-40, sample by sample in float64, plus
a batch form that returns what N successive ``ZFilter.__call__`` invocations return.
"""
from __future__ import annotations

import numpy as np


class WelfordState:
    """n, M (mean), S (sum of squared deviations) - running_stat.py:5-8."""

    def __init__(self, shape=()):
        self.n = 0
        self.M = np.zeros(shape, np.float64)
        self.S = np.zeros(shape, np.float64)

    def push(self, x):  # running_stat.py:9-18
        x = np.asarray(x, np.float64)
        assert x.shape == self.M.shape
        self.n += 1
        if self.n == 1:
            self.M[...] = x
        else:
            old = self.M.copy()
            self.M[...] = old + (x - old) / self.n
            self.S[...] = self.S + (x - old) * (x - self.M)

    @property
    def var(self):  # running_stat.py:26-27 : mean**2 at n == 1
        return self.S / (self.n - 1) if self.n > 1 else np.square(self.M)

    @property
    def std(self):
        return np.sqrt(self.var)


def zfilter_apply(state: WelfordState, x, demean=True, destd=True, clip=10.0, update=True):
    """One ZFilter.__call__ (filters.py:30-38)."""
    x = np.asarray(x, np.float64)
    if update:
        state.push(x)
    if demean:
        x = x - state.M
    if destd:
        x = x / (state.std + 1e-8)
    if clip:
        x = np.clip(x, -clip, clip)
    return x


def zfilter_batch(state: WelfordState, X, demean=True, destd=True, clip=10.0):
    """Rows of X pushed in order; row t is normalised with statistics that include
    rows 0..t and whatever `state` held before.  Mutates `state`."""
    X = np.asarray(X, np.float64)
    out = np.empty_like(X)
    for t in range(X.shape[0]):
        out[t] = zfilter_apply(state, X[t], demean, destd, clip, True)
    return out
