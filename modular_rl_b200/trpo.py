"""Trust Region Policy Optimization - the reference's trpo.py on the B200 path.

``TrpoUpdater(stochpol, usercfg)(paths)`` keeps its signature and its result (an OrderedDict with
surr/kl/ent _before/_after, trpo.py:136-140); the whole natural-gradient step - policy gradient,
conjugate gradient over Fisher-vector products, step scaling, backtracking line search, parameter
update or rollback - is one call into the library (mrl_net_trpo_step), with the CG scalars resident
on the device.  ``cg`` and ``linesearch`` keep the reference's generic callable-based signatures.
"""
from collections import OrderedDict

import numpy as np

from .core import EzFlat, batch_for_paths, concat
from .misc_utils import EzPickle, update_default_config


class TrpoUpdater(EzFlat, EzPickle):
    options = [
        ("cg_damping", float, 1e-3, "Add multiple of the identity to Fisher matrix during CG"),
        ("max_kl", float, 1e-2, "KL divergence between old and new policy (averaged over state-space)"),
    ]

    def __init__(self, stochpol, usercfg):
        EzPickle.__init__(self, stochpol, usercfg)
        self.cfg = update_default_config(self.options, usercfg)
        self.stochpol = stochpol
        EzFlat.__init__(self, stochpol.net)
        self.loss_names = ["surr", "kl", "ent"]
        self.comm = None                      # set by a data-parallel driver (see parallel.py)
        self.last_info = None

    # the three compiled functions of trpo.py:68-70, on an already-bound batch
    def compute_policy_gradient(self, batch):
        return self.stochpol.net.policy_gradient(batch)[0]

    def compute_losses(self, batch):
        return self.stochpol.net.losses(batch)

    def compute_fisher_vector_product(self, p, batch):
        return self.stochpol.net.fvp(batch, p)

    def bind(self, paths):
        """trpo.py:74-77: concatenate the paths and hand (ob, action, advantage, prob) to the device."""
        batch, _ = batch_for_paths(paths, None)   # any cached limit: the policy ignores the time feature
        probtype = self.stochpol.probtype
        prob_np = concat([path["prob"] for path in paths])
        action_na = concat([path["action"] for path in paths])
        advantage_n = concat([path["advantage"] for path in paths])
        dout = self.stochpol.dims[-1]
        batch.set_policy_inputs(probtype.head, dout, action_na, advantage_n, prob_np)
        return batch

    def __call__(self, paths):
        cfg = self.cfg
        batch = self.bind(paths)
        stats, info = self.stochpol.net.trpo_step(batch, cg_damping=cfg["cg_damping"], max_kl=cfg["max_kl"])
        self.last_info = info
        if info["skipped"]:
            print("got zero gradient. not updating")       # trpo.py:102-103
        out = OrderedDict()
        for i, lname in enumerate(self.loss_names):
            out[lname + "_before"] = stats[2 * i]
            out[lname + "_after"] = stats[2 * i + 1]
        return out


def linesearch(f, x, fullstep, expected_improve_rate, max_backtracks=10, accept_ratio=.1):
    """Backtracking line search on a host callable (trpo.py:143-159); expected_improve_rate is the
    slope dy/dx at x.  Returns (True, xnew) for the first step fraction .5**k whose actual/expected
    improvement exceeds accept_ratio (and is positive), else (False, x).  No KL check."""
    fval = f(x)
    for stepfrac in .5 ** np.arange(max_backtracks):
        xnew = x + stepfrac * fullstep
        actual_improve = fval - f(xnew)
        ratio = actual_improve / (expected_improve_rate * stepfrac)
        if ratio > accept_ratio and actual_improve > 0:
            return True, xnew
    return False, x


def cg(f_Ax, b, cg_iters=10, callback=None, verbose=False, residual_tol=1e-10):
    """Conjugate gradient on a host callable f_Ax (Demmel p 312; trpo.py:165-200): x0 = 0, at most
    cg_iters iterations, stops after the update that brings r.r below residual_tol."""
    p, r = b.copy(), b.copy()
    x = np.zeros_like(b)
    rdotr = r.dot(r)
    fmtstr, titlestr = "%10i %10.3g %10.3g", "%10s %10s %10s"
    if verbose:
        print(titlestr % ("iter", "residual norm", "soln norm"))
    i = -1
    for i in range(cg_iters):
        if callback is not None:
            callback(x)
        if verbose:
            print(fmtstr % (i, rdotr, np.linalg.norm(x)))
        z = f_Ax(p)
        v = rdotr / p.dot(z)
        x += v * p
        r -= v * z
        newrdotr = r.dot(r)
        p = r + (newrdotr / rdotr) * p
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    if callback is not None:
        callback(x)
    if verbose:
        print(fmtstr % (i + 1, rdotr, np.linalg.norm(x)))
    return x
