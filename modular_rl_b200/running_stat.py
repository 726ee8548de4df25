"""Running mean / variance with the reference's interface (running_stat.py:4-33).

The per-sample ``push`` is host-side float64 numpy, exactly the reference's Welford update
(it is called once per environment step, next to the simulator).  ``push_batch`` hands a
whole block of samples to the device scan (mrl_zfilter_scan) and leaves the same state behind.
"""
import numpy as np


class RunningStat(object):
    def __init__(self, shape):
        self._n = 0
        self._M = np.zeros(shape)
        self._S = np.zeros(shape)

    def push(self, x):
        x = np.asarray(x)
        assert x.shape == self._M.shape
        self._n += 1
        if self._n == 1:
            self._M[...] = x
            return
        delta = x - self._M
        self._M[...] = self._M + delta / self._n
        self._S[...] = self._S + delta * (x - self._M)

    def state(self):
        """(n, M, S) - what mrl_zfilter_scan consumes and returns."""
        return float(self._n), self._M.reshape(-1).copy(), self._S.reshape(-1).copy()

    def set_state(self, n, M, S):
        self._n = int(n)
        self._M[...] = np.asarray(M).reshape(self._M.shape)
        self._S[...] = np.asarray(S).reshape(self._S.shape)

    @property
    def n(self):
        return self._n

    @property
    def mean(self):
        return self._M

    @property
    def var(self):
        if self._n > 1:
            return self._S / (self._n - 1)
        return np.square(self._M)          # the reference's n == 1 rule (running_stat.py:27)

    @property
    def std(self):
        return np.sqrt(self.var)

    @property
    def shape(self):
        return self._M.shape
