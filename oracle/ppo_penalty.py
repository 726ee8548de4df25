"""Oracle (test infrastructure): PpoLbfgsUpdater.__call__ (ppo.py:59-112) on an
already-concatenated batch, using scipy's fmin_l_bfgs_b like the reference."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import scipy.optimize

from . import policy_math as pm


def ppo_lbfgs_update(theta, spec, ob, act, adv, oldprob, kl_coeff=1.0, kl_target=1e-2,
                     maxiter=25, reverse_kl=False, do_split=False, dtype=np.float64):
    """Returns (info OrderedDict, theta_new, new_kl_coeff, n_evals)."""
    f = dtype
    N = ob.shape[0]
    stop = int(0.75 * N) if do_split else N
    tr = (ob[:stop], act[:stop], adv[:stop], oldprob[:stop])
    te = (ob[stop:], act[stop:], adv[stop:], oldprob[stop:])
    cutoff = kl_target * 2.0
    thprev = np.asarray(theta, f)
    evals = [0]

    def losses3(th, args):
        ls, _, _ = pm.surr_kl_grads(np.asarray(th, f), spec, *args, dtype=f, ratio="lik",
                                    reverse_kl=reverse_kl)
        return ls

    def lossandgrad(th):
        evals[0] += 1
        l, g = pm.ppo_lossgrad(np.asarray(th, f), spec, *tr, kl_coeff, cutoff, f, reverse_kl)
        return float(l), g.astype(np.float64)

    before = losses3(thprev, tr)
    if do_split:
        tbefore = losses3(thprev, te)
    theta_new, _, _ = scipy.optimize.fmin_l_bfgs_b(lossandgrad, thprev.astype(np.float64),
                                                   maxiter=maxiter)
    theta_new = np.asarray(theta_new, f)
    after = losses3(theta_new, tr)
    klafter = after[1]
    if klafter > 1.3 * kl_target:
        kl_coeff *= 1.5
    elif klafter < 0.7 * kl_target:
        kl_coeff /= 1.5
    info = OrderedDict()
    for name, lb, la in zip(("surr", "kl", "ent"), before, after):
        info[name + "_before"] = lb
        info[name + "_after"] = la
        info[name + "_change"] = la - lb
    if do_split:
        tafter = losses3(theta_new, te)
        for name, lb, la in zip(("surr", "kl", "ent"), tbefore, tafter):
            info["test_" + name + "_before"] = lb
            info["test_" + name + "_after"] = la
            info["test_" + name + "_change"] = la - lb
    return info, theta_new, kl_coeff, evals[0]
