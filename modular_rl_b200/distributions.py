"""Host-side numpy helpers with the reference's names (distributions.py).  categorical_sample is
rollout-side (one row per environment step); categorical_kl / categorical_entropy are float64
diagnostics no updater calls."""
import numpy as np


def categorical_sample(prob_nk):
    """Inverse-CDF sampling, one draw per row: argmax(cumsum(p) > U)  (distributions.py:3-13)."""
    prob_nk = np.asarray(prob_nk)
    assert prob_nk.ndim == 2
    u = np.random.rand(prob_nk.shape[0], 1)
    return np.argmax(np.cumsum(prob_nk, axis=1) > u, axis=1)


TINY = np.finfo(np.float64).tiny


def categorical_kl(p_nk, q_nk):
    p = np.asarray(p_nk, dtype=np.float64)
    q = np.asarray(q_nk, dtype=np.float64)
    ratio = p / (q + TINY)
    ratio[p == 0] = 1                       # 0 * log(0/q) := 0
    ratio[(q == 0) & (p != 0)] = np.inf     # p * log(p/0) := inf
    return (p * np.log(ratio)).sum(axis=1)


def categorical_entropy(p_nk):
    p = np.array(p_nk, dtype=np.float64)
    p[p == 0] = 1
    return (-p * np.log(p)).sum(axis=1)
