"""Oracle (test infrastructure): PpoSgdUpdater.__call__ and adam_updates (ppo.py:115-258) on an
already-concatenated batch.  Minibatches of 128 in the order of np.random.permutation (the caller
seeds numpy's global generator, as the reference relies on it); the old policy's probabilities come
from a forward pass at the parameters the update starts from (ppo.py:149, update_old_net ppo.py:174);
every `train` call returns the losses BEFORE its own Adam step (Theano evaluates outputs, then applies
updates)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import policy_math as pm


class Adam:
    """adam_updates (ppo.py:231-258): one shared step counter, a_t = lr*sqrt(1-b2^t)/(1-b1^t),
    step = a_t*m/(sqrt(v)+eps) - note eps is NOT bias-corrected, unlike most library Adams."""

    def __init__(self, n, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, dtype=np.float64):
        self.lr, self.b1, self.b2, self.eps, self.f = learning_rate, beta1, beta2, epsilon, dtype
        self.t = 0
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)

    def step(self, theta, g):
        f = self.f
        g = np.asarray(g, f)
        self.t += 1
        a_t = f(self.lr) * np.sqrt(f(1) - f(self.b2) ** self.t) / (f(1) - f(self.b1) ** self.t)
        self.m = f(self.b1) * self.m + f(1 - self.b1) * g
        self.v = f(self.b2) * self.v + f(1 - self.b2) * g * g
        return np.asarray(theta, f) - a_t * self.m / (np.sqrt(self.v) + f(self.eps))


def _oldprob(theta, spec, ob, dtype):
    _, z = pm.forward(np.asarray(theta, dtype), spec, ob, dtype)
    if spec.head == pm.GAUSS:
        d = spec.dims[-1]
        std = np.exp(np.asarray(theta, dtype)[-d:])
        return np.concatenate([z, np.broadcast_to(std[None, :], z.shape)], axis=1)
    return pm.softmax(z)


def ppo_sgd_update(theta, spec, ob, act, adv, adam: Adam, kl_coeff=1.0, kl_target=1e-2, epochs=10,
                   do_split=False, batchsize=128, dtype=np.float64):
    """Returns (info OrderedDict, theta_new, new_kl_coeff, n_minibatches)."""
    f = dtype
    N = ob.shape[0]
    cutoff = kl_target * 2.0
    theta = np.asarray(theta, f)
    oldprob = _oldprob(theta, spec, ob, f)                     # old net = parameters at entry
    train_stop = (int(.75 * N) // batchsize) * batchsize if do_split else N

    def losses3(th, sl):
        ls, _, _ = pm.surr_kl_grads(th, spec, ob[sl], act[sl], adv[sl], oldprob[sl], dtype=f, ratio="lik")
        return np.asarray(ls, np.float64)

    tr, te = slice(0, train_stop), slice(train_stop, None)
    if do_split:
        test_before = losses3(theta, te)
    before = losses3(theta, tr)
    n_mb = 0
    train_losses = before
    for _ in range(epochs):
        sortinds = np.random.permutation(train_stop)
        losses = []
        for istart in range(0, train_stop, batchsize):
            idx = sortinds[istart:istart + batchsize]
            ls, g_s, g_k = pm.surr_kl_grads(theta, spec, ob[idx], act[idx], adv[idx], oldprob[idx], dtype=f,
                                            ratio="lik")
            over = ls[1] > cutoff
            g = g_s + f(kl_coeff + 2000.0 * over * (ls[1] - cutoff)) * g_k
            losses.append(np.asarray(ls, np.float64))
            theta = adam.step(theta, g)
            n_mb += 1
        train_losses = np.mean(losses, axis=0)
        if do_split:
            test_losses = losses3(theta, te)
    klafter = train_losses[1]
    if klafter > 1.3 * kl_target:
        kl_coeff *= 1.5
    elif klafter < 0.7 * kl_target:
        kl_coeff /= 1.5
    info = OrderedDict()
    for name, lb, la in zip(("surr", "kl", "ent"), before, train_losses):
        info[name + "_before"] = lb
        info[name + "_after"] = la
        info[name + "_change"] = la - lb
    if do_split:
        for name, lb, la in zip(("surr", "kl", "ent"), test_before, test_losses):
            info["test_" + name + "_before"] = lb
            info["test_" + name + "_after"] = la
            info["test_" + name + "_change"] = la - lb
    return info, theta, kl_coeff, n_mb
