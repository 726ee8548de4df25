"""Oracle (test infrastructure): policy MLP, DiagGauss / Categorical math,
surrogate losses, policy gradient and Fisher-vector product.

Restates, in numpy, what the reference builds symbolically in Theano:

* network architecture and parameter order  - agentzoo.py:25-61, core.py:311-313,
  core.py:518-557 (flat vector = per Dense layer [kernel(in,out) C-order, bias],
  ConcatFixedStd's logstd last - core.py:708-725);
* DiagGauss - core.py:402-438;  Categorical - core.py:339-365;
* surr / kl / ent - trpo.py:37-42,60-63;  policy gradient - trpo.py:43;
* Fisher-vector product - trpo.py:45-58 (closed form  (1/N) sum_n J^T M J v, which
  equals the reverse-over-reverse derivative the reference takes because
  KL(stopgrad(pi) || pi) is stationary at pi; `oracle.autodiff_check` differentiates
  the reference's own expressions to pin this);
* PPO penalised surrogate - ppo.py:35-49.

All functions take a ``dtype`` (float64 = the oracle proper, float32 = emulation of
the fork's floatX=float32 run, used as the CPU baseline / noise floor).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import math

import numpy as np

GAUSS, CAT, VALUE = "gauss", "cat", "value"
_ACTS = ("tanh", "relu", "sigmoid")


@dataclass(frozen=True)
class NetSpec:
    """dims = [d0, h1, ..., h_{L-1}, d_out]; head in {gauss, cat, value}."""
    dims: Tuple[int, ...]
    head: str = GAUSS
    activation: str = "tanh"

    def __post_init__(self):
        assert self.head in (GAUSS, CAT, VALUE)
        assert self.activation in _ACTS
        assert len(self.dims) >= 2

    @property
    def n_layers(self) -> int:
        return len(self.dims) - 1


def param_slices(spec: NetSpec):
    """[(name, shape, start, stop)] in trainable_weights order."""
    out, pos = [], 0
    for l in range(spec.n_layers):
        din, dout = spec.dims[l], spec.dims[l + 1]
        out.append((f"W{l + 1}", (din, dout), pos, pos + din * dout)); pos += din * dout
        out.append((f"b{l + 1}", (dout,), pos, pos + dout)); pos += dout
    if spec.head == GAUSS:
        d = spec.dims[-1]
        out.append(("logstd", (d,), pos, pos + d)); pos += d
    return out


def num_params(spec: NetSpec) -> int:
    return param_slices(spec)[-1][3]


def split_params(theta, spec: NetSpec):
    """-> (Ws, bs, logstd|None) views of the flat vector (core.py:531-535)."""
    Ws, bs, logstd = [], [], None
    for name, shape, a, b in param_slices(spec):
        v = theta[a:b].reshape(shape)
        if name[0] == "W":
            Ws.append(v)
        elif name[0] == "b":
            bs.append(v)
        else:
            logstd = v
    return Ws, bs, logstd


def init_params(spec: NetSpec, rng: np.random.Generator, last_scale: float = 0.1) -> np.ndarray:
    """Keras-2.0.2 Dense defaults: glorot-uniform kernel, zero bias; the last kernel is
    scaled by 0.1 (agentzoo.py:39-48) for policies (the VF net is NOT scaled,
    agentzoo.py:53-59 -> pass last_scale=1); logstd zeros (core.py:716)."""
    th = np.zeros(num_params(spec), np.float32)
    Ws, _, _ = split_params(th, spec)
    for l, W in enumerate(Ws):
        lim = np.sqrt(6.0 / (W.shape[0] + W.shape[1]))
        W[...] = rng.uniform(-lim, lim, size=W.shape).astype(np.float32)
        if l == len(Ws) - 1:
            W *= np.float32(last_scale)
    return th


# ---------------------------------------------------------------- activations
def _act(z, kind):
    if kind == "tanh":
        return np.tanh(z)
    if kind == "relu":
        return np.maximum(z, 0)
    return 1.0 / (1.0 + np.exp(-z))


def _dact_from_h(h, kind):
    if kind == "tanh":
        return 1.0 - h * h
    if kind == "relu":
        return (h > 0).astype(h.dtype)
    return h * (1.0 - h)


def forward(theta, spec: NetSpec, x, dtype=np.float64):
    """-> (hs, z) with hs = [x, h1, ..., h_{L-1}], z = last pre-head output [N, d_out]."""
    th = np.asarray(theta, dtype)
    Ws, bs, _ = split_params(th, spec)
    h = np.asarray(x, dtype)
    hs = [h]
    for l in range(spec.n_layers - 1):
        h = _act(h @ Ws[l] + bs[l], spec.activation)
        hs.append(h)
    z = h @ Ws[-1] + bs[-1]
    return hs, z


def softmax(z):
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def head_prob(theta, spec: NetSpec, z, dtype=np.float64):
    """The `prob` row the reference's net outputs: [mean, std] (core.py:722-725) or
    softmax probabilities (agentzoo.py:44); value nets return z."""
    if spec.head == GAUSS:
        _, _, logstd = split_params(np.asarray(theta, dtype), spec)
        std = np.broadcast_to(np.exp(logstd)[None, :], z.shape)
        return np.concatenate([z, std], axis=1)
    if spec.head == CAT:
        return softmax(z)
    return z


# ---------------------------------------------------------------- distributions
def gauss_loglik(a, prob, d):  # core.py:412-416
    mean, std = prob[:, :d], prob[:, d:]
    return (-0.5 * np.square((a - mean) / std).sum(axis=1)
            - 0.5 * math.log(2.0 * math.pi) * d - np.log(std).sum(axis=1))


def gauss_kl(prob0, prob1, d):  # core.py:421-426
    m0, s0, m1, s1 = prob0[:, :d], prob0[:, d:], prob1[:, :d], prob1[:, d:]
    return (np.log(s1 / s0).sum(axis=1)
            + ((np.square(s0) + np.square(m0 - m1)) / (2.0 * np.square(s1))).sum(axis=1) - 0.5 * d)


def gauss_entropy(prob, d):  # core.py:428-430
    return np.log(prob[:, d:]).sum(axis=1) + 0.5 * math.log(2 * math.pi * math.e) * d


def cat_lik(a, prob):  # core.py:349-350
    return prob[np.arange(prob.shape[0]), np.asarray(a).astype(np.int64)]


def cat_kl(prob0, prob1):  # core.py:355-356
    return (prob0 * np.log(prob0 / prob1)).sum(axis=1)


def cat_entropy(prob):  # core.py:358-359
    return -(prob * np.log(prob)).sum(axis=1)


def loglik(spec, a, prob):
    if spec.head == GAUSS:
        return gauss_loglik(a, prob, spec.dims[-1])
    return np.log(cat_lik(a, prob))


def kl_rows(spec, prob0, prob1):
    if spec.head == GAUSS:
        return gauss_kl(prob0, prob1, spec.dims[-1])
    return cat_kl(prob0, prob1)


def entropy_rows(spec, prob):
    if spec.head == GAUSS:
        return gauss_entropy(prob, spec.dims[-1])
    return cat_entropy(prob)


# ---------------------------------------------------------------- losses
def losses(theta, spec, ob, act, adv, oldprob, dtype=np.float64, ratio="logdiff"):
    """[surr, kl, ent] - trpo.py:37-42,60-63 (ratio='logdiff') or ppo.py:35-47
    (ratio='lik': p/oldp through `likelihood`)."""
    _, z = forward(theta, spec, ob, dtype)
    prob = head_prob(theta, spec, z, dtype)
    oldprob = np.asarray(oldprob, dtype)
    adv = np.asarray(adv, dtype)
    a = np.asarray(act, dtype) if spec.head == GAUSS else np.asarray(act)
    N = ob.shape[0]
    if ratio == "logdiff":
        rho = np.exp(loglik(spec, a, prob) - loglik(spec, a, oldprob))
    else:
        if spec.head == GAUSS:
            rho = np.exp(loglik(spec, a, prob)) / np.exp(loglik(spec, a, oldprob))
        else:
            rho = cat_lik(a, prob) / cat_lik(a, oldprob)
    surr = dtype(-1.0 / N) * rho.dot(adv)
    kl = kl_rows(spec, oldprob, prob).mean()
    ent = entropy_rows(spec, prob).mean()
    return np.array([surr, kl, ent], dtype)


def _backprop(spec, Ws, hs, dz, logstd_grad=None):
    """Reverse sweep shared by gradient and Fvp: dz = dLoss/dz_L [N, d_out] ->
    flat gradient in param order."""
    grads = []
    delta = dz
    for l in range(spec.n_layers - 1, -1, -1):
        gW = hs[l].T @ delta
        gb = delta.sum(axis=0)
        grads.append((gW, gb))
        if l > 0:
            delta = (delta @ Ws[l].T) * _dact_from_h(hs[l], spec.activation)
    flat = []
    for gW, gb in reversed(grads):
        flat += [gW.ravel(), gb]
    if logstd_grad is not None:
        flat.append(logstd_grad)
    return np.concatenate(flat)


def surr_kl_grads(theta, spec, ob, act, adv, oldprob, dtype=np.float64, ratio="logdiff",
                  reverse_kl=False):
    """-> (losses[3], grad_surr[P], grad_kl[P]) ; grad_kl is of mean KL(old||new)
    (or KL(new||old) when reverse_kl, ppo.py:40-43).  Closed-form derivatives of
    SURVEY A.2."""
    th = np.asarray(theta, dtype)
    Ws, bs, logstd = split_params(th, spec)
    hs, z = forward(th, spec, ob, dtype)
    prob = head_prob(th, spec, z, dtype)
    oldprob = np.asarray(oldprob, dtype)
    adv = np.asarray(adv, dtype)
    N = ob.shape[0]
    d = spec.dims[-1]
    if spec.head == GAUSS:
        a = np.asarray(act, dtype)
        std = prob[:, d:]
        mu = z
        lp = gauss_loglik(a, prob, d)
        olp = gauss_loglik(a, oldprob, d)
        rho = np.exp(lp - olp) if ratio == "logdiff" else np.exp(lp) / np.exp(olp)
        w = (-rho * adv / N)[:, None]
        dz_s = w * (a - mu) / np.square(std)
        dls_s = (w * (np.square((a - mu) / std) - 1.0)).sum(axis=0)
        m0, s0 = oldprob[:, :d], oldprob[:, d:]
        if not reverse_kl:
            klr = gauss_kl(oldprob, prob, d)
            dz_k = (mu - m0) / np.square(std) / N
            dls_k = (1.0 - (np.square(s0) + np.square(m0 - mu)) / np.square(std)).sum(axis=0) / N
        else:
            klr = gauss_kl(prob, oldprob, d)
            dz_k = (mu - m0) / np.square(s0) / N
            dls_k = (-1.0 + np.square(std) / np.square(s0)).sum(axis=0) / N
        ent = gauss_entropy(prob, d).mean()
        g_s = _backprop(spec, Ws, hs, dz_s, dls_s)
        g_k = _backprop(spec, Ws, hs, dz_k, dls_k)
    else:
        a = np.asarray(act).astype(np.int64)
        p = prob
        p0 = oldprob
        rho = (np.exp(np.log(cat_lik(a, p)) - np.log(cat_lik(a, p0))) if ratio == "logdiff"
               else cat_lik(a, p) / cat_lik(a, p0))
        w = (-rho * adv / N)[:, None]
        onehot = np.zeros_like(p)
        onehot[np.arange(N), a] = 1.0
        dz_s = w * (onehot - p)
        if not reverse_kl:
            klr = cat_kl(p0, p)
            dz_k = (p - p0) / N
        else:
            klr = cat_kl(p, p0)
            dz_k = p * (np.log(p / p0) - klr[:, None]) / N
        ent = cat_entropy(p).mean()
        g_s = _backprop(spec, Ws, hs, dz_s)
        g_k = _backprop(spec, Ws, hs, dz_k)
    ls = np.array([dtype(-1.0 / N) * rho.dot(adv), klr.mean(), ent], dtype)
    return ls, g_s, g_k


def policy_gradient(theta, spec, ob, act, adv, oldprob, dtype=np.float64):
    """flatgrad(surr, params) - trpo.py:43."""
    return surr_kl_grads(theta, spec, ob, act, adv, oldprob, dtype)[1]


def fisher_vector_product(theta, spec, ob, v, dtype=np.float64):
    """trpo.py:45-58 in closed form.  `v` is the flat tangent (the
    reference declares it T.fvector, trpo.py:48, i.e. float32 on entry)."""
    th = np.asarray(theta, dtype)
    v = np.asarray(np.asarray(v, np.float32), dtype)  # fvector downcast
    Ws, bs, logstd = split_params(th, spec)
    Vs, vbs, vls = split_params(v, spec)
    hs, z = forward(th, spec, ob, dtype)
    N = ob.shape[0]
    Rh = np.zeros_like(hs[0])
    Rz = None
    for l in range(spec.n_layers):
        Rz = Rh @ Ws[l] + hs[l] @ Vs[l] + vbs[l]
        if l < spec.n_layers - 1:
            Rh = _dact_from_h(hs[l + 1], spec.activation) * Rz
    if spec.head == GAUSS:
        var = np.exp(2.0 * logstd)
        dz = Rz / var / N
        return _backprop(spec, Ws, hs, dz, 2.0 * vls)
    p = softmax(z)
    dz = (p * Rz - p * (p * Rz).sum(axis=1, keepdims=True)) / N
    return _backprop(spec, Ws, hs, dz)


def ppo_lossgrad(theta, spec, ob, act, adv, oldprob, kl_coeff, kl_cutoff, dtype=np.float64,
                 reverse_kl=False):
    """(pensurr, flatgrad) - ppo.py:47-49,56.  pensurr = surr + kl_coeff*kl +
    1000*(kl>cut)*(kl-cut)^2."""
    ls, g_s, g_k = surr_kl_grads(theta, spec, ob, act, adv, oldprob, dtype, ratio="lik",
                                 reverse_kl=reverse_kl)
    surr, kl = ls[0], ls[1]
    over = kl > kl_cutoff
    pen = surr + kl_coeff * kl + 1000.0 * over * (kl - kl_cutoff) ** 2
    coef = kl_coeff + 2000.0 * over * (kl - kl_cutoff)
    return dtype(pen), g_s + dtype(coef) * g_k


# ---------------------------------------------------------------- sampling
def categorical_sample(prob_nk, u):
    """distributions.py:3-13 with the uniform draws passed in: argmax(cumsum(p) > u)."""
    cs = np.cumsum(np.asarray(prob_nk), axis=1)
    return np.argmax(cs > np.asarray(u).reshape(-1, 1), axis=1)


def gauss_sample(prob, d, eps):
    """core.py:432-435 with the normal draws passed in."""
    return eps * prob[:, d:] + prob[:, :d]
