#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE.

Runs only in the authoring container (needs /root/reference; the GPU box has
neither).  Theano/Keras/TF cannot be imported here, so:

* numpy-only reference modules (misc_utils, running_stat, filters, distributions) are
  imported as they are, under a synthetic package name;
* the reference's symbolic classes/functions (DiagGauss, Categorical, ConcatFixedStd,
  compute_advantage in core.py; cg, linesearch in trpo.py) are cut out of their files
  with `ast` and compiled UNMODIFIED in a namespace where the Theano module `T` is a
  thin numpy or torch shim - so the formulas evaluated are the reference's lines;
* the three Theano graphs of TrpoUpdater.__init__ (trpo.py:37-61) and the PPO /
  value-function losses (ppo.py:35-49, core.py:613-617) are rebuilt on those classes
  with torch.float64 and differentiated by torch autograd, reverse-over-reverse for
  the Fisher-vector product exactly as trpo.py:45-58 does.

Usage:  python tests/golden/make_golden.py     (writes next to this file)
"""
import ast
import contextlib
import importlib
import io
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

REF = "/root/reference/modular_rl"
OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)


# ------------------------------------------------------------------ loaders
def ref_package():
    pkg = types.ModuleType("refmrl")
    pkg.__path__ = [REF]
    sys.modules["refmrl"] = pkg
    mods = {}
    for name in ("misc_utils", "running_stat", "filters", "distributions"):
        mods[name] = importlib.import_module("refmrl." + name)
    import scipy.signal  # noqa: F401  (misc_utils does `import scipy` only)
    return mods


def cut(path, names):
    """Source nodes (ClassDef/FunctionDef) named `names`, compiled as written."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    assert {n.name for n in keep} == set(names), (names, [n.name for n in keep])
    mod = ast.Module(body=keep, type_ignores=[])
    return compile(mod, path, "exec")


class NumpyT:
    square, log, exp, arange, concatenate = np.square, np.log, np.exp, np.arange, np.concatenate

    @staticmethod
    def repeat(x, n, axis=0):
        return np.repeat(x, n, axis=axis)


class TorchT:
    @staticmethod
    def _t(x):
        return x if torch.is_tensor(x) else torch.as_tensor(x, dtype=torch.float64)

    @staticmethod
    def square(x):
        return torch.square(TorchT._t(x))

    @staticmethod
    def log(x):
        return torch.log(TorchT._t(x))

    @staticmethod
    def exp(x):
        return torch.exp(TorchT._t(x))

    @staticmethod
    def arange(n):
        return torch.arange(int(n))

    @staticmethod
    def concatenate(xs, axis=0):
        return torch.cat(list(xs), dim=axis)

    @staticmethod
    def repeat(x, n, axis=0):
        return torch.repeat_interleave(x, int(n), dim=axis)


def load_probtypes(Tshim, mods):
    ns = dict(T=Tshim, np=np, distributions=mods["distributions"], floatX="float64")
    ns["Layer"] = type("Layer", (), {"__init__": lambda self, **kw: None})
    exec(cut(os.path.join(REF, "core.py"),
             ["ProbType", "DiagGauss", "Categorical", "ConcatFixedStd"]), ns)
    return ns


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ------------------------------------------------------------------ pieces
def gen_discount(mods, out):
    disc = mods["misc_utils"].discount
    rng = np.random.default_rng(101)
    kat = np.array([1, 10, 100, 1e3, 1e4, 1e5, 1e6])
    out["disc_kat_x"] = kat
    out["disc_kat_y"] = disc(kat, 0.99)                       # x.py:762
    for i, (T, g) in enumerate([(1, 0.99), (7, 0.5), (300, 0.995), (1000, 0.97 * 0.995)]):
        x = rng.standard_normal(T)
        out[f"disc{i}_x"], out[f"disc{i}_g"], out[f"disc{i}_y"] = x, np.float64(g), disc(x, g)
    x2 = rng.standard_normal((20, 3))
    out["disc2d_x"], out["disc2d_y"] = x2, disc(x2, 0.9)


def gen_advantage(mods, out):
    """Runs core.compute_advantage as written (fork cross-check lines included: the
    stub `gae` recomputes the advantage by the mask/bootstrapped recurrence of
    a.py:178-186, so the reference's own 1e-4 RuntimeError guard is live)."""
    rng = np.random.default_rng(202)
    w = rng.standard_normal(4)
    gamma, lam = 0.98, 0.93

    class Reg:
        class ez_for_net:
            @staticmethod
            def gf():
                return None

    class VF:
        reg = Reg

        @staticmethod
        def predict(path):
            return np.tanh(path["observation"] @ w).astype(np.float32)

        @staticmethod
        def preproc(x):
            return x

    class Gae:
        class m:
            trainable_variables = None

        @staticmethod
        def A(S, R, M):
            V = np.tanh(S @ w).astype(np.float32).astype(np.float64)
            T = len(R)
            adv = np.zeros(T)
            run = 0.0
            for t in range(T - 1, -1, -1):
                delta = R[t, 0] + gamma * V[t + 1] * M[t, 0] - V[t]
                run = delta + gamma * lam * run
                adv[t] = run
            return adv

    ns = dict(np=np, discount=mods["misc_utils"].discount, gae=Gae, reshape=lambda *a, **k: None)
    exec(cut(os.path.join(REF, "core.py"), ["compute_advantage"]), ns)
    lens = [5, 1, 9, 30, 2]
    term = [True, False, False, True, True]
    paths = []
    for T, tm in zip(lens, term):
        paths.append(dict(observation=rng.standard_normal((T, 4)), reward=rng.standard_normal(T),
                          terminated=tm))
    with quiet():
        ns["compute_advantage"](VF, paths, gamma, lam)
    out["adv_w"] = w
    out["adv_gamma"], out["adv_lam"] = np.float64(gamma), np.float64(lam)
    out["adv_lens"] = np.array(lens, np.int64)
    out["adv_term"] = np.array(term, np.uint8)
    out["adv_obs"] = np.concatenate([p["observation"] for p in paths])
    out["adv_reward"] = np.concatenate([p["reward"] for p in paths])
    out["adv_baseline"] = np.concatenate([p["baseline"] for p in paths])
    out["adv_return"] = np.concatenate([p["return"] for p in paths])
    out["adv_advantage"] = np.concatenate([p["advantage"] for p in paths])


def gen_filters(mods, out):
    RS = mods["running_stat"].RunningStat
    ZF = mods["filters"].ZFilter
    np.random.seed(0)
    with quiet():
        mods["running_stat"].test_running_stat()            # the reference's own unit test
    rng = np.random.default_rng(303)
    X = rng.standard_normal((40, 3)) * np.array([1.0, 10.0, 0.01]) + np.array([0.0, 5.0, -2.0])
    rs = RS((3,))
    means, vars_ = [], []
    for x in X:
        rs.push(x)
        means.append(rs.mean.copy()); vars_.append(np.array(rs.var))
    out["rs_x"], out["rs_mean"], out["rs_var"] = X, np.array(means), np.array(vars_)
    zf = ZF((3,), clip=5)
    out["zf_ob_y"] = np.array([zf(x) for x in X])
    r = rng.standard_normal(40) * 3
    zr = ZF((), demean=False, clip=10)
    out["zf_rew_x"] = r
    out["zf_rew_y"] = np.array([zr(x) for x in r])
    out["zf_noupdate_y"] = np.array(zf(X[0], update=False))


def gen_cg_ls(out):
    ns = dict(np=np)
    exec(cut(os.path.join(REF, "trpo.py"), ["cg", "linesearch"]), ns)
    rng = np.random.default_rng(404)
    B = rng.standard_normal((12, 12))
    A = B @ B.T + 0.5 * np.eye(12)
    b = rng.standard_normal(12)
    with quiet():
        out["cg_x10"] = ns["cg"](lambda p: A @ p, b)
        out["cg_x3"] = ns["cg"](lambda p: A @ p, b, cg_iters=3)
        # early break: well-conditioned system converges below 1e-10 before 10 iterations
        A2 = np.eye(12) * 2.0 + 0.01 * (B + B.T)
        out["cg_x_early"] = ns["cg"](lambda p: A2 @ p, b)
    out["cg_A"], out["cg_b"], out["cg_A2"] = A, b, A2
    # line search on a quadratic bowl; three regimes: accept k=0, accept k>0, failure
    Q = np.diag(np.linspace(1, 4, 6))
    x0 = np.ones(6)
    f = lambda x: 0.5 * x @ Q @ x
    g0 = Q @ x0
    for name, scale in (("ls_a", 0.2), ("ls_b", 3.0), ("ls_c", -1.0)):
        full = -scale * g0
        rate = -g0 @ full
        with quiet():
            ok, xn = ns["linesearch"](f, x0, full, rate)
        out[name + "_ok"], out[name + "_x"], out[name + "_full"], out[name + "_rate"] = \
            np.bool_(ok), xn, full, np.float64(rate)
    out["ls_Q"], out["ls_x0"] = Q, x0
    return ns


def gen_dists(mods, out):
    ns = load_probtypes(NumpyT, mods)
    rng = np.random.default_rng(505)
    g = ns["DiagGauss"](4)
    p0 = np.concatenate([rng.standard_normal((16, 4)), np.exp(0.3 * rng.standard_normal((16, 4)))], 1)
    p1 = np.concatenate([rng.standard_normal((16, 4)), np.exp(0.3 * rng.standard_normal((16, 4)))], 1)
    p0[0] = [-.2, .3, .4, -.5, 1.1, 1.5, .1, 1.9]           # core.py:445
    a = rng.standard_normal((16, 4))
    out["g_p0"], out["g_p1"], out["g_a"] = p0, p1, a
    out["g_loglik"], out["g_kl"], out["g_ent"] = g.loglikelihood(a, p0), g.kl(p0, p1), g.entropy(p0)
    c = ns["Categorical"](3)
    q0 = rng.dirichlet(np.ones(3), 16); q0[0] = [.2, .3, .5]  # core.py:452
    q1 = rng.dirichlet(np.ones(3), 16)
    ai = rng.integers(0, 3, 16)
    out["c_p0"], out["c_p1"], out["c_a"] = q0, q1, ai
    out["c_lik"], out["c_loglik"] = c.likelihood(ai, q0), c.loglikelihood(ai, q0)
    out["c_kl"], out["c_ent"] = c.kl(q0, q1), c.entropy(q0)
    np.random.seed(7)
    out["c_sample_seed7"] = mods["distributions"].categorical_sample(q0)
    np.random.seed(7)
    out["c_sample_u"] = np.random.rand(16, 1)
    np.random.seed(9)
    out["g_sample_seed9"] = g.sample(p0)
    np.random.seed(9)
    out["g_sample_eps"] = np.random.randn(16, 4)
    yp, y = rng.standard_normal((30, 1)), rng.standard_normal((30, 1))
    out["ev_yp"], out["ev_y"] = yp, y
    out["ev"] = mods["misc_utils"].explained_variance_2d(yp, y)


# ------------------------------------------------------------------ Theano graphs in torch
def dense_forward(params, x, hid_act=torch.tanh):
    """Keras Dense: act(x @ kernel + bias) (agentzoo.py:34-48); params = [W1,b1,...]."""
    h = x
    n = len(params) // 2
    for l in range(n):
        h = h @ params[2 * l] + params[2 * l + 1]
        if l < n - 1:
            h = hid_act(h)
    return h


def flat(ts):
    return torch.cat([t.reshape(-1) for t in ts])


def build_policy(ns, dims, head, rng):
    params = []
    for l in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        W = rng.uniform(-lim, lim, (dims[l], dims[l + 1]))
        if l == len(dims) - 2:
            W = W * 0.1
        params += [torch.tensor(W, requires_grad=True),
                   torch.tensor(0.05 * rng.standard_normal(dims[l + 1]), requires_grad=True)]
    if head == "gauss":
        cfs = ns["ConcatFixedStd"]()
        cfs.logstd = torch.tensor(0.2 * rng.standard_normal(dims[-1]), requires_grad=True)
        params.append(cfs.logstd)
        probtype = ns["DiagGauss"](dims[-1])

        def net(x):
            return cfs.call(dense_forward(params[:-1], x), None)     # core.py:722-725
    else:
        probtype = ns["Categorical"](dims[-1])

        def net(x):
            return torch.softmax(dense_forward(params, x), dim=1)
    return params, probtype, net


def set_flat(params, th):
    pos = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(torch.as_tensor(th[pos:pos + n], dtype=torch.float64).reshape(p.shape))
            pos += n


def gen_trpo(mods, cgls, out):
    ns = load_probtypes(TorchT, mods)
    for tag, dims, head, N in (("tg", (5, 8, 6, 3), "gauss", 96), ("tc", (4, 8, 8, 3), "cat", 96)):
        rng = np.random.default_rng(606 if head == "gauss" else 707)
        params, probtype, net = build_policy(ns, dims, head, rng)
        ob = torch.tensor(np.clip(rng.standard_normal((N, dims[0])), -5, 5))
        with torch.no_grad():
            oldprob = net(ob).clone()
        # move the policy a little so that kl/ratio are non-trivial
        th0 = flat(params).detach().numpy().copy()
        theta = th0 + 0.05 * rng.standard_normal(th0.size)
        set_flat(params, theta)
        if head == "gauss":
            np.random.seed(11)
            act = torch.tensor(probtype.sample(oldprob.numpy()))
        else:
            np.random.seed(11)
            act = torch.tensor(probtype.sample(oldprob.numpy()))
        adv = rng.standard_normal(N)
        adv = torch.tensor((adv - adv.mean()) / adv.std())

        def graphs():
            prob = net(ob)
            logp = probtype.loglikelihood(act, prob)
            oldlogp = probtype.loglikelihood(act, oldprob)
            surr = (-1.0 / N) * torch.exp(logp - oldlogp).dot(adv)            # trpo.py:42
            kl_ff = probtype.kl(prob.detach(), prob).sum() / N                # trpo.py:45-46
            ent = probtype.entropy(prob).mean()
            kl = probtype.kl(oldprob, prob).mean()
            return surr, kl_ff, ent, kl

        def compute_losses():
            with torch.no_grad():
                s, _, e, k = graphs()
            return np.array([s.item(), k.item(), e.item()])

        def compute_pg():
            s, _, _, _ = graphs()
            return flat(torch.autograd.grad(s, params)).detach().numpy()

        def compute_fvp(v):
            _, kl_ff, _, _ = graphs()
            grads = torch.autograd.grad(kl_ff, params, create_graph=True)     # trpo.py:47
            v32 = torch.as_tensor(np.asarray(v, np.float32).astype(np.float64))  # T.fvector, trpo.py:48
            gvp = (flat(grads) * v32).sum()                                   # trpo.py:56
            return flat(torch.autograd.grad(gvp, params)).detach().numpy()    # trpo.py:58

        v = rng.standard_normal(theta.size).astype(np.float32).astype(np.float64)
        out[tag + "_dims"] = np.array(dims)
        out[tag + "_theta"], out[tag + "_ob"], out[tag + "_act"] = theta, ob.numpy(), act.numpy()
        out[tag + "_adv"], out[tag + "_oldprob"], out[tag + "_v"] = adv.numpy(), oldprob.numpy(), v
        out[tag + "_losses"] = compute_losses()
        out[tag + "_pg"] = compute_pg()
        out[tag + "_fvp"] = compute_fvp(v)

        # ---- the canonical TrpoUpdater.__call__ (trpo.py:72-140), reference cg/linesearch
        for cfgtag, damping, max_kl in (("d", 1e-3, 1e-2), ("b", 0.1, 0.01)):
            set_flat(params, theta)
            thprev = theta.copy()
            g = compute_pg()
            lb = compute_losses()
            with quiet():
                stepdir = cgls["cg"](lambda p: compute_fvp(p) + damping * p, -g)
            shs = .5 * stepdir.dot(compute_fvp(stepdir) + damping * stepdir)
            lm = np.sqrt(shs / max_kl)
            fullstep = stepdir / lm
            neggdotstepdir = -g.dot(stepdir)

            def loss(th):
                set_flat(params, th)
                return compute_losses()[0]
            with quiet():
                success, th_new = cgls["linesearch"](loss, thprev, fullstep, neggdotstepdir / lm)
            set_flat(params, th_new)
            la = compute_losses()
            key = f"{tag}_{cfgtag}_"
            out[key + "stepdir"], out[key + "fullstep"], out[key + "theta_new"] = stepdir, fullstep, th_new
            out[key + "success"] = np.bool_(success)
            out[key + "before"], out[key + "after"] = lb, la
            out[key + "cfg"] = np.array([damping, max_kl])

        # ---- PPO penalised surrogate (ppo.py:35-49) at kl_coeff 1.0 and with the cutoff active
        set_flat(params, theta)
        for ptag, klc, cutoff, rev in (("p0", 1.0, 0.02, 0), ("p1", 0.7, 1e-4, 0), ("p2", 1.3, 0.02, 1)):
            prob = net(ob)
            p_n = probtype.likelihood(act, prob)
            oldp_n = probtype.likelihood(act, oldprob)
            kl = (probtype.kl(prob, oldprob) if rev else probtype.kl(oldprob, prob)).mean()
            surr = (-1.0 / N) * (p_n / oldp_n).dot(adv)
            pens = surr + klc * kl + 1000 * (kl > cutoff) * torch.square(kl - cutoff)
            gp = flat(torch.autograd.grad(pens, params)).detach().numpy()
            out[f"{tag}_{ptag}_cfg"] = np.array([klc, cutoff, rev])
            out[f"{tag}_{ptag}_pen"] = np.float64(pens.item())
            out[f"{tag}_{ptag}_grad"] = gp
            out[f"{tag}_{ptag}_losses"] = np.array([surr.item(), kl.item(),
                                                    probtype.entropy(prob).mean().item()])


def gen_vf(out):
    """NnRegression loss (core.py:613-617) on a 4+1 -> 8 -> 8 -> 1 net; NnVf.preproc
    (core.py:659-660) restated inline (it is a one-liner on numpy)."""
    rng = np.random.default_rng(808)
    dims = (5, 8, 8, 1)
    params = []
    for l in range(3):
        lim = np.sqrt(6.0 / (dims[l] + dims[l + 1]))
        params += [torch.tensor(rng.uniform(-lim, lim, (dims[l], dims[l + 1])), requires_grad=True),
                   torch.tensor(0.05 * rng.standard_normal(dims[l + 1]), requires_grad=True)]
    lens = [6, 11, 3]
    obs = [rng.standard_normal((T, 4)) for T in lens]
    tl = 50
    x = np.concatenate([np.concatenate([o, np.arange(len(o)).reshape(-1, 1) / float(tl)], axis=1)
                        for o in obs], axis=0)
    y = rng.standard_normal((x.shape[0], 1))
    xt, yt = torch.tensor(x), torch.tensor(y)
    pred = dense_forward(params, xt)
    l2 = 1e-3 * sum(torch.square(p).sum() for p in params)
    mse = torch.sum(torch.square(yt - pred)) / x.shape[0]
    loss = mse + l2
    g = flat(torch.autograd.grad(loss, params)).detach().numpy()
    out["vf_dims"] = np.array(dims)
    out["vf_theta"] = flat(params).detach().numpy()
    out["vf_lens"], out["vf_tl"] = np.array(lens, np.int64), np.int64(tl)
    out["vf_obs"] = np.concatenate(obs)
    out["vf_x"], out["vf_y"] = x, y
    out["vf_pred"] = pred.detach().numpy()
    out["vf_losses"] = np.array([loss.item(), mse.item(), l2.item()])
    out["vf_grad"] = g


def gen_cem(out):
    """The reference's cem generator (cem.py:10-50), executed unmodified.  Its pool=None branch wraps a py2
    `map` in np.array and cannot run on Python 3, so the population is evaluated through the `pool.map`
    branch with an in-process stand-in for the pool."""
    ns = dict(np=np)
    exec(cut(os.path.join(REF, "cem.py"), ["cem"]), ns)

    class Pool(object):
        def map(self, f, xs):
            return list(map(f, xs))

    c = np.linspace(-2.0, 3.0, 7)
    f = lambda th: -np.sum((th - c) ** 2) + 0.1 * np.sin(th).sum()
    np.random.seed(77)
    with quiet():
        infos = list(ns["cem"](f, np.zeros(7, np.float32), 40, 6, 0.2, initial_std=1.5, extra_std=0.4,
                               std_decay_time=3.0, pool=Pool()))
    out["cem_center"] = c
    out["cem_th"] = np.array([i["th"] for i in infos])
    out["cem_ys"] = np.array([i["ys"] for i in infos])
    out["cem_std"] = np.array([i["std"] for i in infos])
    out["cem_ymean"] = np.array([i["ymean"] for i in infos])


def main_cem():
    out = OrderedDict()
    gen_cem(out)
    path = os.path.join(OUT, "cem_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


def main():
    if sys.argv[1:] == ["cem"]:       # the CEM fixture lives in its own file (added after the main one)
        return main_cem()
    mods = ref_package()
    out = OrderedDict()
    gen_discount(mods, out)
    gen_advantage(mods, out)
    gen_filters(mods, out)
    cgls = gen_cg_ls(out)
    gen_dists(mods, out)
    gen_trpo(mods, cgls, out)
    gen_vf(out)
    path = os.path.join(OUT, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
