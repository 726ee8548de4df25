"""Synthetic batches in the shape of BASELINE.json's configs.

Weights: glorot-uniform kernels, zero biases, last policy kernel x0.1, logstd = 0
(agentzoo.py:34-48, core.py:716).  Observations: N(0,1) clipped to +-5, i.e. what
ZFilter(clip=5) emits (agentzoo.py:89).  oldprob = the policy's own output at theta
(rounded to float32, as `act` stores it in path["prob"], core.py:261-267); actions are
sampled from it; trajectories have i.i.d. uniform lengths in [50, T_max], truncated so that
they sum to N; 80 % terminate.  Everything is seeded.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np

GAUSS, CAT, VALUE = 0, 1, 2


@dataclass(frozen=True)
class Workload:
    name: str
    dims: Tuple[int, ...]
    head: int
    N: int
    t_max: int
    seed: int


WORKLOADS = {
    "cartpole": Workload("cartpole", (4, 64, 64, 2), CAT, 5000, 200, 0),
    "hopper": Workload("hopper", (11, 64, 64, 3), GAUSS, 50_000, 1000, 1),
    "humanoid": Workload("humanoid", (376, 100, 50, 25, 17), GAUSS, 1_000_000, 1000, 2),
    "walker2d": Workload("walker2d", (17, 64, 64, 6), GAUSS, 200_000, 1000, 3),
    "cat128": Workload("cat128", (128, 64, 64, 18), CAT, 4_000_000, 5000, 4),
}


def num_params(dims, head) -> int:
    p = sum(dims[l] * dims[l + 1] + dims[l + 1] for l in range(len(dims) - 1))
    return p + (dims[-1] if head == GAUSS else 0)


def init_params(dims, head, rng: np.random.Generator, last_scale: float = 0.1) -> np.ndarray:
    chunks = []
    L = len(dims) - 1
    for l in range(L):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        W = rng.uniform(-lim, lim, size=(dims[l], dims[l + 1])).astype(np.float32)
        if l == L - 1:
            W *= np.float32(last_scale)
        chunks += [W.ravel(), np.zeros(dims[l + 1], np.float32)]
    if head == GAUSS:
        chunks.append(np.zeros(dims[-1], np.float32))
    return np.concatenate(chunks)


def make_paths(N: int, t_max: int, rng: np.random.Generator, t_min: int = 50, p_term: float = 0.8):
    """-> (offsets int64[n_paths+1], terminated uint8[n_paths])"""
    t_min = min(t_min, t_max)
    lens = []
    tot = 0
    while tot < N:
        T = int(rng.integers(t_min, t_max + 1))
        T = min(T, N - tot)
        lens.append(T)
        tot += T
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    terminated = (rng.random(len(lens)) < p_term).astype(np.uint8)
    return offsets, terminated


def make_obs(N: int, d0: int, rng: np.random.Generator, dtype=np.float32) -> np.ndarray:
    ob = rng.standard_normal((N, d0), dtype=np.float32)
    np.clip(ob, -5, 5, out=ob)
    return ob.astype(dtype, copy=False)


def sample_actions(head: int, prob: np.ndarray, rng: np.random.Generator):
    """DiagGauss.sample (core.py:432-435) / categorical_sample (distributions.py:3-13)."""
    if head == GAUSS:
        d = prob.shape[1] // 2
        eps = rng.standard_normal((prob.shape[0], d), dtype=np.float32)
        return (eps * prob[:, d:] + prob[:, :d]).astype(np.float32)
    cs = np.cumsum(prob, axis=1)
    u = rng.random((prob.shape[0], 1), dtype=np.float32)
    return np.argmax(cs > u, axis=1).astype(np.int32)


def policy_batch(wl: Workload, forward_fn: Callable[[np.ndarray, np.ndarray], np.ndarray],
                 N: Optional[int] = None, ob_dtype=np.float32):
    """forward_fn(theta, ob) -> net output [N, dout] (means | probabilities).
    Returns dict(theta, ob, act, adv, oldprob, offsets, terminated, reward)."""
    N = int(N or wl.N)
    rng = np.random.default_rng(wl.seed)
    theta = init_params(wl.dims, wl.head, rng)
    ob = make_obs(N, wl.dims[0], rng, ob_dtype)
    out = np.asarray(forward_fn(theta, ob), np.float32)
    if wl.head == GAUSS:
        d = wl.dims[-1]
        std = np.broadcast_to(np.exp(theta[-d:])[None, :], out.shape)
        oldprob = np.concatenate([out, std], axis=1).astype(np.float32)
    else:
        oldprob = out
    act = sample_actions(wl.head, oldprob, rng)
    adv = rng.standard_normal(N)
    adv = ((adv - adv.mean()) / adv.std()).astype(np.float32)
    offsets, terminated = make_paths(N, wl.t_max, rng)
    reward = rng.standard_normal(N)
    return dict(theta=theta, ob=ob, act=act, adv=adv, oldprob=oldprob, offsets=offsets,
                terminated=terminated, reward=reward)


def perturb(theta: np.ndarray, scale: float, seed: int) -> np.ndarray:
    """theta + scale * N(0,1): moves the policy off theta_old so kl/ratio are non-trivial in tests."""
    rng = np.random.default_rng(seed)
    return (theta + scale * rng.standard_normal(theta.size)).astype(np.float32)
