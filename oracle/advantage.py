"""Oracle (test infrastructure): discounted reverse scan and GAE.

``discount`` uses the identical third-party routine the reference calls
(scipy.signal.lfilter, misc_utils.py:27); ``discount_recurrence`` is the scalar
recurrence the fork asserts it is bit-equal to (a.py:15-23, x.py:496-504).
``compute_advantage`` follows core.py:63-75,100-105 (fork lines 79-96 excluded).
"""
from __future__ import annotations

import numpy as np
import scipy.signal


def discount(x, gamma):
    x = np.asarray(x)
    assert x.ndim >= 1
    return scipy.signal.lfilter([1], [1, -gamma], x[::-1], axis=0)[::-1]


def discount_recurrence(x, gamma):
    """v = v*gamma + x[t], t = T-1 .. 0, float64 (a.py:18-21)."""
    x = np.asarray(x, np.float64)
    out = np.empty_like(x)
    v = np.zeros(x.shape[1:], np.float64)
    for t in range(len(x) - 1, -1, -1):
        v = v * gamma + x[t]
        out[t] = v
    return out


def compute_advantage(predict, paths, gamma, lam):
    """In-place on `paths` like the reference: writes return / baseline / advantage.
    `predict(path) -> [T]` is NnVf.predict (core.py:648-650)."""
    for path in paths:
        path["return"] = discount(path["reward"], gamma)
        b = path["baseline"] = predict(path)
        b1 = np.append(b, 0 if path["terminated"] else b[-1])
        deltas = path["reward"] + gamma * b1[1:] - b1[:-1]
        path["advantage"] = discount(deltas, gamma * lam)
    alladv = np.concatenate([path["advantage"] for path in paths])
    std = alladv.std()
    mean = alladv.mean()
    for path in paths:
        path["advantage"] = (path["advantage"] - mean) / std


def gae_flat(reward, baseline, offsets, terminated, gamma, lam):
    """Same arithmetic on the flat CSR layout the C-ABI uses (include/mrl_b200.h):
    reward/baseline [N], offsets int64[n_paths+1], terminated uint8[n_paths].
    Returns (returns, advantages_unstandardised) as float64."""
    reward = np.asarray(reward, np.float64)
    baseline = np.asarray(baseline)          # keep float32 if the VF net produced float32:
    ret = np.empty_like(reward)              # np.append(f32, f32) stays f32, so gamma*b1 is
    adv = np.empty_like(reward)              # rounded to f32 on unterminated paths (core.py:73-74)
    for p in range(len(offsets) - 1):
        a, b = int(offsets[p]), int(offsets[p + 1])
        r, v = reward[a:b], baseline[a:b]
        v1 = np.append(v, 0 if terminated[p] else v[-1])
        ret[a:b] = discount(r, gamma)
        adv[a:b] = discount(r + gamma * v1[1:] - v1[:-1], gamma * lam)
    return ret, adv


def standardize(adv):
    adv = np.asarray(adv, np.float64)
    return (adv - adv.mean()) / adv.std()


def time_index(offsets):
    """Within-path step index for every flat timestep (NnVf.preproc's arange,
    core.py:659-660) - part of the bit-exact integer contract."""
    offsets = np.asarray(offsets, np.int64)
    n = int(offsets[-1])
    t = np.arange(n, dtype=np.int64)
    pid = np.searchsorted(offsets, t, side="right") - 1
    return t - offsets[pid], pid
