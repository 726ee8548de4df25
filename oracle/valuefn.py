"""Oracle (test infrastructure): neural value function (NnVf / NnRegression /
LbfgsOptimizer), core.py:595-697, and misc_utils.explained_variance_2d
(misc_utils.py:44-49).

The optimiser is the same third-party routine the reference calls
(scipy.optimize.fmin_l_bfgs_b, core.py:687) with the reference's defaults.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import scipy.optimize

from . import policy_math as pm


def preproc(ob_no, timestep_limit):
    """core.py:659-660: append the column arange(T)/timestep_limit."""
    ob_no = np.asarray(ob_no)
    t = np.arange(len(ob_no)).reshape(-1, 1) / float(timestep_limit)
    return np.concatenate([ob_no, t], axis=1)


def vf_forward(theta, spec, x, dtype=np.float64):
    _, z = pm.forward(theta, spec, x, dtype)
    return z  # [N, 1]


def vf_losses(theta, spec, x, target, dtype=np.float64, l2coeff=1e-3):
    """[loss, mse, l2] - core.py:613-617 (l2 over ALL trainable weights, biases too)."""
    th = np.asarray(theta, dtype)
    y = vf_forward(th, spec, x, dtype)
    N = x.shape[0]
    mse = np.square(np.asarray(target, dtype) - y).sum() / N
    l2 = dtype(l2coeff) * np.square(th).sum()
    return np.array([mse + l2, mse, l2], dtype)


def vf_lossgrad(theta, spec, x, target, dtype=np.float64, l2coeff=1e-3):
    th = np.asarray(theta, dtype)
    Ws, bs, _ = pm.split_params(th, spec)
    hs, y = pm.forward(th, spec, x, dtype)
    N = x.shape[0]
    diff = y - np.asarray(target, dtype)
    loss = np.square(diff).sum() / N + dtype(l2coeff) * np.square(th).sum()
    g = pm._backprop(spec, Ws, hs, 2.0 * diff / N) + dtype(2.0 * l2coeff) * th
    return dtype(loss), g


def explained_variance_2d(ypred, y):
    assert y.ndim == 2 and ypred.ndim == 2
    vary = np.var(y, axis=0)
    out = 1 - np.var(y - ypred) / vary      # un-axised numerator, as the reference
    out[vary < 1e-10] = 0
    return out


def regression_fit(theta, spec, x, ytarg, mixfrac=1.0, maxiter=25, dtype=np.float64):
    """NnRegression.fit + LbfgsOptimizer.update (core.py:619-637,674-697).
    Returns (stats OrderedDict, theta_new, n_evals)."""
    f = dtype
    thprev = np.asarray(theta, f)
    ytarg = np.asarray(ytarg)
    ypredold = vf_forward(thprev, spec, x, f)
    target = ytarg * mixfrac + ypredold * (1 - mixfrac)
    evals = [0]

    def lossandgrad(th):
        evals[0] += 1
        l, g = vf_lossgrad(np.asarray(th, f), spec, x, target, f)   # set_params casts to floatX
        return float(l), g.astype(np.float64)

    before = vf_losses(thprev, spec, x, target, f)
    theta_new, _, _ = scipy.optimize.fmin_l_bfgs_b(lossandgrad, thprev.astype(np.float64),
                                                   maxiter=maxiter)
    theta_new = np.asarray(theta_new, f)
    after = vf_losses(theta_new, spec, x, target, f)
    out = OrderedDict()
    for name, lb, la in zip(("loss", "mse", "l2"), before, after):
        out[name + "_before"] = lb
        out[name + "_after"] = la
    yprednew = vf_forward(theta_new, spec, x, f)
    out["PredStdevBefore"] = ypredold.std()
    out["PredStdevAfter"] = yprednew.std()
    out["TargStdev"] = ytarg.std()
    out["EV_before"] = explained_variance_2d(ypredold, ytarg)[0]
    out["EV_after"] = explained_variance_2d(yprednew, ytarg)[0]
    return out, theta_new, evals[0]
