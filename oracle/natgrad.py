"""Oracle (test infrastructure): conjugate gradient, backtracking line search and the
TRPO natural-gradient step.

Follows trpo.py:165-200 (cg), trpo.py:143-159 (linesearch) and the canonical part of
TrpoUpdater.__call__ (trpo.py:72-140 minus the fork's TensorFlow cross-check lines
7,9,82-84,90-91,97-100,106-117,131-132).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import policy_math as pm


def conjugate_gradient(f_Ax, b, cg_iters=10, residual_tol=1e-10):
    """Demmel p.312 as the reference writes it (trpo.py:165-200).
    Returns (x, iterations_run, final_rdotr)."""
    p = b.copy()
    r = b.copy()
    x = np.zeros_like(b)
    rdotr = r.dot(r)
    it = 0
    for it in range(1, cg_iters + 1):
        z = f_Ax(p)
        alpha = rdotr / p.dot(z)
        x += alpha * p
        r -= alpha * z
        new_rdotr = r.dot(r)
        beta = new_rdotr / rdotr
        p = r + beta * p
        rdotr = new_rdotr
        if rdotr < residual_tol:      # break AFTER the update (trpo.py:192-193)
            break
    return x, it, rdotr


def backtracking_linesearch(f, x, fullstep, expected_improve_rate, max_backtracks=10,
                            accept_ratio=0.1):
    """trpo.py:143-159.  Returns (success, x_out, n_evals_of_f, accepted_index|-1)."""
    fval = f(x)
    evals = 1
    for k, stepfrac in enumerate(0.5 ** np.arange(max_backtracks)):
        xnew = x + stepfrac * fullstep
        newfval = f(xnew)
        evals += 1
        actual = fval - newfval
        expected = expected_improve_rate * stepfrac
        ratio = actual / expected
        if ratio > accept_ratio and actual > 0:
            return True, xnew, evals, k
    return False, x, evals, -1


def trpo_update(theta, spec, ob, act, adv, oldprob, cg_damping=1e-3, max_kl=1e-2,
                cg_iters=10, dtype=np.float64):
    """One TrpoUpdater.__call__ on an already-concatenated batch.

    dtype=float64: the oracle proper.  dtype=float32: emulates the fork's
    floatX=float32 run (all Theano outputs float32, CG in float32 numpy,
    theta cast to floatX on every set_params_flat - core.py:540).
    Returns (stats OrderedDict, info dict with g, stepdir, fullstep, theta_new, ...).
    """
    f = dtype
    thprev = np.asarray(theta, f).copy()
    args = (ob, act, adv, oldprob)

    def fvp(p):
        return (pm.fisher_vector_product(thprev, spec, ob, p, f).astype(f)
                + f(cg_damping) * p)

    g = pm.policy_gradient(thprev, spec, *args, dtype=f).astype(f)
    losses_before = pm.losses(thprev, spec, *args, dtype=f)
    info = dict(g=g, n_fvp=0, n_loss_evals=1)
    theta_new = thprev
    if np.allclose(g, 0):                                  # trpo.py:102-103
        info.update(skipped=True, success=False, accepted_index=-1, cg_iters_run=0)
    else:
        stepdir, iters, rdotr = conjugate_gradient(fvp, -g, cg_iters)
        shs = 0.5 * stepdir.dot(fvp(stepdir))              # trpo.py:119
        lm = np.sqrt(shs / max_kl)
        fullstep = stepdir / lm
        neggdotstepdir = -g.dot(stepdir)

        def loss(th):                                      # trpo.py:126-128
            return pm.losses(np.asarray(th, f), spec, *args, dtype=f)[0]

        success, theta_new, evals, kacc = backtracking_linesearch(
            loss, thprev, fullstep, neggdotstepdir / lm)
        theta_new = np.asarray(theta_new, f)
        info.update(skipped=False, stepdir=stepdir, fullstep=fullstep, shs=shs, lm=lm,
                    expected_improve_rate=neggdotstepdir / lm, success=success,
                    accepted_index=kacc, cg_iters_run=iters, cg_rdotr=rdotr,
                    n_fvp=iters + 1, n_loss_evals=1 + evals + 1)
    losses_after = pm.losses(theta_new, spec, *args, dtype=f)
    out = OrderedDict()
    for name, lb, la in zip(("surr", "kl", "ent"), losses_before, losses_after):
        out[name + "_before"] = lb
        out[name + "_after"] = la
    info["theta_new"] = theta_new
    return out, info
