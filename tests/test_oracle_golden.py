"""The oracle against the reference: golden vectors produced by executing the
reference's own code (tests/golden/make_golden.py) and the reference's in-file tests
(running_stat.py:35-46, core.py:441-483, x.py:762)."""
import numpy as np
import pytest

from oracle import advantage as adv_o
from oracle import natgrad, policy_math as pm, ppo_penalty, valuefn, zfilter
from conftest import relerr


# ----------------------------------------------------------------- discount / GAE
def test_discount_kat_and_recurrence(golden):
    y = adv_o.discount(golden["disc_kat_x"], 0.99)
    assert np.array_equal(y, golden["disc_kat_y"])
    assert np.array_equal(y, adv_o.discount_recurrence(golden["disc_kat_x"], 0.99))  # a.py:15-23
    expect = [1047264.323491, 1057841.7409, 1068516.91, 1079209., 1089100., 1090000., 1000000.]
    assert np.allclose(y, expect, rtol=1e-12)
    for i in range(4):
        x, g = golden[f"disc{i}_x"], float(golden[f"disc{i}_g"])
        assert np.array_equal(adv_o.discount(x, g), golden[f"disc{i}_y"])
        assert np.array_equal(adv_o.discount_recurrence(x, g), golden[f"disc{i}_y"])
    assert np.array_equal(adv_o.discount(golden["disc2d_x"], 0.9), golden["disc2d_y"])


def _golden_paths(golden):
    lens = golden["adv_lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    w = golden["adv_w"]
    paths = []
    for p, T in enumerate(lens):
        a, b = off[p], off[p + 1]
        paths.append(dict(observation=golden["adv_obs"][a:b], reward=golden["adv_reward"][a:b],
                          terminated=bool(golden["adv_term"][p])))
    predict = lambda path: np.tanh(path["observation"] @ w).astype(np.float32)
    return paths, predict, off


def test_compute_advantage_matches_reference(golden):
    paths, predict, off = _golden_paths(golden)
    adv_o.compute_advantage(predict, paths, float(golden["adv_gamma"]), float(golden["adv_lam"]))
    cat = lambda k: np.concatenate([p[k] for p in paths])
    assert np.array_equal(cat("return"), golden["adv_return"])
    assert np.array_equal(cat("baseline"), golden["adv_baseline"])
    assert np.array_equal(cat("advantage"), golden["adv_advantage"])


def test_gae_flat_matches_reference(golden):
    paths, predict, off = _golden_paths(golden)
    base = golden["adv_baseline"]
    ret, adv = adv_o.gae_flat(golden["adv_reward"], base, off, golden["adv_term"],
                              float(golden["adv_gamma"]), float(golden["adv_lam"]))
    assert np.array_equal(ret, golden["adv_return"])
    assert np.allclose(adv_o.standardize(adv), golden["adv_advantage"], rtol=1e-13, atol=1e-14)
    t, pid = adv_o.time_index(off)
    assert t.dtype == np.int64 and np.array_equal(t[off[:-1]], np.zeros(len(off) - 1))
    assert np.array_equal(np.bincount(pid), golden["adv_lens"])


# ----------------------------------------------------------------- RunningStat / ZFilter
def test_running_stat_reference_unit_test():
    """running_stat.py:35-46 restated against the oracle's Welford state."""
    rng = np.random.default_rng(0)
    for shp in ((), (3,), (3, 4)):
        li, rs = [], zfilter.WelfordState(shp)
        for _ in range(5):
            val = rng.standard_normal(shp)
            rs.push(val)
            li.append(val)
            m = np.mean(li, axis=0)
            assert np.allclose(rs.M, m)
            v = np.square(m) if len(li) == 1 else np.var(li, ddof=1, axis=0)
            assert np.allclose(rs.var, v)


def test_zfilter_matches_reference(golden):
    X = golden["rs_x"]
    st = zfilter.WelfordState((3,))
    for t, x in enumerate(X):
        st.push(x)
        assert np.array_equal(st.M, golden["rs_mean"][t])
        assert np.array_equal(st.var, golden["rs_var"][t])
    st = zfilter.WelfordState((3,))
    y = zfilter.zfilter_batch(st, X, clip=5)
    assert np.array_equal(y, golden["zf_ob_y"])
    assert np.all(y[0] == 0)                                   # n==1 rule, SURVEY A.5
    assert np.array_equal(zfilter.zfilter_apply(st, X[0], clip=5, update=False), golden["zf_noupdate_y"])
    st = zfilter.WelfordState(())
    yr = np.array([zfilter.zfilter_apply(st, r, demean=False, clip=10) for r in golden["zf_rew_x"]])
    assert np.array_equal(yr, golden["zf_rew_y"])


# ----------------------------------------------------------------- cg / linesearch
def test_cg_matches_reference(golden):
    A, b, A2 = golden["cg_A"], golden["cg_b"], golden["cg_A2"]
    x, it, _ = natgrad.conjugate_gradient(lambda p: A @ p, b)
    assert it == 10 and np.array_equal(x, golden["cg_x10"])
    x, it, _ = natgrad.conjugate_gradient(lambda p: A @ p, b, cg_iters=3)
    assert it == 3 and np.array_equal(x, golden["cg_x3"])
    x, it, rd = natgrad.conjugate_gradient(lambda p: A2 @ p, b)
    assert it < 10 and rd < 1e-10 and np.array_equal(x, golden["cg_x_early"])


def test_linesearch_matches_reference(golden):
    Q, x0 = golden["ls_Q"], golden["ls_x0"]
    f = lambda x: 0.5 * x @ Q @ x
    seen = set()
    for name in ("ls_a", "ls_b", "ls_c"):
        ok, xn, _, k = natgrad.backtracking_linesearch(f, x0, golden[name + "_full"],
                                                       float(golden[name + "_rate"]))
        assert ok == bool(golden[name + "_ok"])
        assert np.array_equal(xn, golden[name + "_x"])
        seen.add((ok, k > 0))
    assert (True, False) in seen and (True, True) in seen and (False, False) in seen


# ----------------------------------------------------------------- distributions
def test_distribution_formulas(golden):
    g = golden
    assert np.allclose(pm.gauss_loglik(g["g_a"], g["g_p0"], 4), g["g_loglik"], rtol=1e-14)
    assert np.allclose(pm.gauss_kl(g["g_p0"], g["g_p1"], 4), g["g_kl"], rtol=1e-14)
    assert np.allclose(pm.gauss_entropy(g["g_p0"], 4), g["g_ent"], rtol=1e-14)
    assert np.array_equal(pm.cat_lik(g["c_a"], g["c_p0"]), g["c_lik"])
    assert np.allclose(pm.cat_kl(g["c_p0"], g["c_p1"]), g["c_kl"], rtol=1e-14)
    assert np.allclose(pm.cat_entropy(g["c_p0"]), g["c_ent"], rtol=1e-14)
    assert np.array_equal(pm.categorical_sample(g["c_p0"], g["c_sample_u"]), g["c_sample_seed7"])
    assert np.allclose(pm.gauss_sample(g["g_p0"], 4, g["g_sample_eps"]), g["g_sample_seed9"], rtol=1e-15)
    assert np.allclose(valuefn.explained_variance_2d(g["ev_yp"], g["ev_y"]), g["ev"], rtol=1e-14)


@pytest.mark.parametrize("kind", ["gauss", "cat"])
def test_probtype_monte_carlo_identities(kind):
    """core.py:441-483: E[-log p] == entropy and KL[p,q] == -H[p] - E_p[log q], 3 sigma."""
    rng = np.random.default_rng(0)
    N = 100000
    if kind == "gauss":
        prob = np.array([-.2, .3, .4, -.5, 1.1, 1.5, .1, 1.9])
        spec = pm.NetSpec((1, 4), pm.GAUSS)
        M = np.repeat(prob[None], N, 0)
        X = pm.gauss_sample(M, 4, rng.standard_normal((N, 4)))
    else:
        prob = np.array([.2, .3, .5])
        spec = pm.NetSpec((1, 3), pm.CAT)
        M = np.repeat(prob[None], N, 0)
        X = pm.categorical_sample(M, rng.random((N, 1)))
    ll = pm.loglik(spec, X, M)
    ent = pm.entropy_rows(spec, M).mean()
    assert abs(ent + ll.mean()) < 3 * ll.std() / np.sqrt(N)
    q = prob + rng.standard_normal(prob.size) * 0.1
    if kind == "cat":
        q = np.abs(q) / np.abs(q).sum()
    else:
        q[4:] = np.abs(q[4:])
    M2 = np.repeat(q[None], N, 0)
    kl = pm.kl_rows(spec, M, M2).mean()
    ll2 = pm.loglik(spec, X, M2)
    assert abs(kl - (-ent - ll2.mean())) < 3 * ll2.std() / np.sqrt(N)


# ----------------------------------------------------------------- TRPO graphs
def _case(golden, tag):
    head = pm.GAUSS if tag == "tg" else pm.CAT
    spec = pm.NetSpec(tuple(int(d) for d in golden[tag + "_dims"]), head)
    return spec, golden[tag + "_theta"], golden[tag + "_ob"], golden[tag + "_act"], \
        golden[tag + "_adv"], golden[tag + "_oldprob"]


@pytest.mark.parametrize("tag", ["tg", "tc"])
def test_losses_gradient_fvp_vs_autodiff_of_reference(golden, tag):
    spec, th, ob, act, adv, oldp = _case(golden, tag)
    assert pm.num_params(spec) == th.size
    assert np.allclose(pm.losses(th, spec, ob, act, adv, oldp), golden[tag + "_losses"], rtol=1e-12, atol=1e-15)
    assert relerr(pm.policy_gradient(th, spec, ob, act, adv, oldp), golden[tag + "_pg"]) < 1e-12
    assert relerr(pm.fisher_vector_product(th, spec, ob, golden[tag + "_v"]), golden[tag + "_fvp"]) < 1e-12


@pytest.mark.parametrize("tag", ["tg", "tc"])
@pytest.mark.parametrize("cfg", ["d", "b"])
def test_trpo_step_vs_reference_pipeline(golden, tag, cfg):
    spec, th, ob, act, adv, oldp = _case(golden, tag)
    damping, max_kl = golden[f"{tag}_{cfg}_cfg"]
    stats, info = natgrad.trpo_update(th, spec, ob, act, adv, oldp, damping, max_kl)
    key = f"{tag}_{cfg}_"
    assert relerr(info["stepdir"], golden[key + "stepdir"]) < 1e-9
    assert relerr(info["fullstep"], golden[key + "fullstep"]) < 1e-9
    assert info["success"] == bool(golden[key + "success"])
    assert relerr(info["theta_new"], golden[key + "theta_new"]) < 1e-10
    before = [stats[k + "_before"] for k in ("surr", "kl", "ent")]
    after = [stats[k + "_after"] for k in ("surr", "kl", "ent")]
    assert np.allclose(before, golden[key + "before"], rtol=1e-11, atol=1e-14)
    assert np.allclose(after, golden[key + "after"], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("tag", ["tg", "tc"])
@pytest.mark.parametrize("ptag", ["p0", "p1", "p2"])
def test_ppo_lossgrad_vs_autodiff_of_reference(golden, tag, ptag):
    spec, th, ob, act, adv, oldp = _case(golden, tag)
    klc, cutoff, rev = golden[f"{tag}_{ptag}_cfg"]
    pen, g = pm.ppo_lossgrad(th, spec, ob, act, adv, oldp, klc, cutoff, reverse_kl=bool(rev))
    assert np.isclose(pen, golden[f"{tag}_{ptag}_pen"], rtol=1e-12)
    assert relerr(g, golden[f"{tag}_{ptag}_grad"]) < 1e-11
    ls, _, _ = pm.surr_kl_grads(th, spec, ob, act, adv, oldp, ratio="lik", reverse_kl=bool(rev))
    assert np.allclose(ls, golden[f"{tag}_{ptag}_losses"], rtol=1e-12)
    if ptag == "p1":
        assert ls[1] > cutoff      # the 1000*(kl-cut)^2 branch is live in this fixture


def test_ppo_update_runs_and_adapts(golden):
    spec, th, ob, act, adv, oldp = _case(golden, "tg")
    info, th_new, klc, evals = ppo_penalty.ppo_lbfgs_update(th, spec, ob, act, adv, oldp, maxiter=5)
    assert set(info) == {f"{n}_{s}" for n in ("surr", "kl", "ent") for s in ("before", "after", "change")}
    assert evals >= 2 and klc in (1.0, 1.5, 1 / 1.5) and th_new.shape == th.shape
    assert info["surr_after"] <= info["surr_before"] + 1e-6


# ----------------------------------------------------------------- value function
def test_vf_loss_grad_preproc(golden):
    spec = pm.NetSpec(tuple(int(d) for d in golden["vf_dims"]), pm.VALUE)
    lens = golden["vf_lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    x = np.concatenate([valuefn.preproc(golden["vf_obs"][off[i]:off[i + 1]], int(golden["vf_tl"]))
                        for i in range(len(lens))])
    assert np.array_equal(x, golden["vf_x"])
    t, _ = adv_o.time_index(off)
    assert np.array_equal(t / float(golden["vf_tl"]), x[:, -1])
    th = golden["vf_theta"]
    assert np.allclose(valuefn.vf_forward(th, spec, x), golden["vf_pred"], rtol=1e-13)
    assert np.allclose(valuefn.vf_losses(th, spec, x, golden["vf_y"]), golden["vf_losses"], rtol=1e-13)
    l, g = valuefn.vf_lossgrad(th, spec, x, golden["vf_y"])
    assert np.isclose(l, golden["vf_losses"][0], rtol=1e-13)
    assert relerr(g, golden["vf_grad"]) < 1e-12
    stats, th_new, evals = valuefn.regression_fit(th, spec, x, golden["vf_y"], mixfrac=0.1, maxiter=25)
    assert stats["loss_after"] < stats["loss_before"] and evals > 2
    assert set(stats) >= {"loss_before", "mse_after", "l2_after", "PredStdevBefore", "PredStdevAfter",
                          "TargStdev", "EV_before", "EV_after"}


def test_fp32_emulation_noise_floor(golden):
    """The float32 restatement (the fork's floatX) differs from float64 by ~1e-6: this is
    the reference-equivalent noise floor quoted in DESIGN.md."""
    spec, th, ob, act, adv, oldp = _case(golden, "tg")
    g64 = pm.policy_gradient(th, spec, ob, act, adv, oldp)
    g32 = pm.policy_gradient(th, spec, ob, act, adv, oldp, dtype=np.float32)
    assert g32.dtype == np.float32 and relerr(g32, g64) < 1e-5
    f64 = pm.fisher_vector_product(th, spec, ob, golden["tg_v"])
    f32 = pm.fisher_vector_product(th, spec, ob, golden["tg_v"], dtype=np.float32)
    assert f32.dtype == np.float32 and relerr(f32, f64) < 1e-5
