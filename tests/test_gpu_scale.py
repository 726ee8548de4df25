"""Parity at BASELINE.json's FULL sizes (hard part 1: float32 accumulation over N = 1e6 sums).

The CUDA path through the C ABI against the float64 oracle on
  * configs[2]: 1 M-timestep Humanoid batch - losses, policy gradient, one Fisher-vector product;
  * configs[3]: 200 k-timestep Walker2d batch - PpoLbfgs penalised surrogate and its gradient;
  * configs[4]: 4 M-timestep ragged Categorical batch - GAE / returns, NnVf loss and gradient;
each within the 1e-5 relative tolerance north_star states.  The measured errors are written to
gpurun_out/parity_fullsize.json (copied to profiles/ by the builder).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import relerr  # noqa: E402

TOL = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESULTS = {}


@pytest.fixture(scope="module", autouse=True)
def _dump():
    yield
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_fullsize.json"), "w") as f:
            json.dump(RESULTS, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _policy_case(name):
    from modular_rl_b200 import device, synth
    from oracle import policy_math as pm
    wl = synth.WORKLOADS[name]
    spec = pm.NetSpec(wl.dims, pm.GAUSS if wl.head == 0 else pm.CAT)
    rng = np.random.default_rng(wl.seed)
    theta0 = synth.init_params(wl.dims, wl.head, rng)
    ob = synth.make_obs(wl.N, wl.dims[0], rng)
    off, term = synth.make_paths(wl.N, wl.t_max, rng)
    net = device.DeviceNet(wl.dims, wl.head)
    batch = device.DeviceBatch(wl.dims[0], with_time_feature=True)
    batch.set_obs(ob).set_paths(off, term, float(wl.t_max))
    net.set_params(theta0)
    out = net.forward(batch)                      # oldprob = the policy's own float32 output (core.py:261-267)
    d = wl.dims[-1]
    if wl.head == synth.GAUSS:
        oldprob = np.concatenate([out, np.broadcast_to(np.exp(theta0[-d:])[None], out.shape)], 1).astype(np.float32)
    else:
        oldprob = out
    act = synth.sample_actions(wl.head, oldprob, rng)
    adv = rng.standard_normal(wl.N)
    adv = ((adv - adv.mean()) / adv.std()).astype(np.float32)
    batch.set_policy_inputs(wl.head, d, act, adv, oldprob)
    theta = synth.perturb(theta0, 0.01, wl.seed + 7)
    net.set_params(theta)
    return wl, spec, net, batch, theta, (ob, act, adv, oldprob)


def test_humanoid_1m_losses_gradient_fvp():
    from oracle import policy_math as pm
    wl, spec, net, batch, theta, args = _policy_case("humanoid")
    assert wl.N == 1_000_000
    ls = net.losses(batch)
    g, _ = net.policy_gradient(batch)
    v = np.random.default_rng(12345).standard_normal(net.P).astype(np.float32)
    f = net.fvp(batch, v)
    ols = pm.losses(theta, spec, *args)
    og = pm.policy_gradient(theta, spec, *args)
    of = pm.fisher_vector_product(theta, spec, args[0], v)
    og32 = pm.policy_gradient(theta, spec, *args, dtype=np.float32)
    RESULTS["humanoid_1M"] = dict(losses=[float(x) for x in ls], oracle_losses=[float(x) for x in ols],
                                  grad_rel_l2=relerr(g, og), fvp_rel_l2=relerr(f, of),
                                  reference_float32_grad_rel_l2=relerr(og32, og))
    assert np.allclose(ls, ols, rtol=TOL, atol=2e-7), (ls, ols)
    assert relerr(g, og) < TOL
    assert relerr(f, of) < TOL


def test_walker_200k_ppo_lossgrad():
    from oracle import policy_math as pm
    wl, spec, net, batch, theta, args = _policy_case("walker2d")
    assert wl.N == 200_000
    klc, cutoff = 0.7, 2e-5                      # cutoff below the batch KL: the 1000 (kl - cut)^2 term is live
    pen, g, ls = net.ppo_lossgrad(batch, klc, cutoff, False)
    open_, og = pm.ppo_lossgrad(theta, spec, *args, klc, cutoff)
    ols, _, _ = pm.surr_kl_grads(theta, spec, *args, ratio="lik")
    dpen_dkl = klc + 2000.0 * (ols[1] > cutoff) * (ols[1] - cutoff)
    RESULTS["walker2d_200k_ppo"] = dict(pensurr=float(pen), oracle_pensurr=float(open_), grad_rel_l2=relerr(g, og),
                                        kl=float(ls[1]), oracle_kl=float(ols[1]))
    assert ols[1] > cutoff
    assert abs(pen - open_) < TOL * (abs(ols[0]) + abs(dpen_dkl) * ols[1]) + 1e-7
    assert relerr(g, og) < 2 * TOL


def test_cat_4m_gae_and_vf_lossgrad():
    """configs[4]: GAE + NnVf value fit on the long-horizon ragged Categorical batch (obs 128, 4 M timesteps)."""
    from modular_rl_b200 import device, synth
    from oracle import advantage as oadv, policy_math as pm, valuefn as vfo
    wl = synth.WORKLOADS["cat128"]
    N = wl.N
    assert N == 4_000_000
    rng = np.random.default_rng(wl.seed)
    ob = synth.make_obs(N, wl.dims[0], rng)
    off, term = synth.make_paths(N, wl.t_max, rng)
    reward = rng.standard_normal(N)
    vdims = (wl.dims[0] + 1, 64, 64, 1)
    vtheta = synth.init_params(vdims, synth.VALUE, rng, last_scale=1.0)
    vf = device.DeviceNet(vdims, synth.VALUE)
    batch = device.DeviceBatch(wl.dims[0], with_time_feature=True)
    batch.set_obs(ob).set_paths(off, term, float(wl.t_max))
    vf.set_params(vtheta)
    # bit-exact integer contract at full size: within-path time index
    tidx, _ = oadv.time_index(off)
    assert np.array_equal(batch.time_index(), tidx.astype(np.int32))
    base = vf.forward(batch)[:, 0]
    vf.predict_into_baseline(batch)
    ret, adv = batch.gae(reward, None, 0.995, 0.97, standardize=True)
    oret, oad = oadv.gae_flat(reward, base.astype(np.float64), off, term, 0.995, 0.97)
    oads = oadv.standardize(oad)
    RESULTS["cat128_4M_gae"] = dict(returns_rel_l2=relerr(ret, oret), advantages_max_abs=float(np.max(np.abs(adv - oads))),
                                    n_paths=int(len(term)))
    assert np.allclose(ret, oret, rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(adv - oads)) < 1e-9
    # NnVf loss / gradient (core.py:613-617) against float64 on the full batch
    batch.set_vf_target(ret)
    ls, g = vf.vf_lossgrad(batch, 1e-3)
    spec = pm.NetSpec(vdims, pm.VALUE)
    x = np.empty((N, vdims[0]), np.float64)
    x[:, :-1] = ob
    x[:, -1] = tidx / float(wl.t_max)
    del ob
    pred = vfo.vf_forward(vtheta, spec, x)
    assert relerr(base, pred[:, 0]) < TOL
    ols = vfo.vf_losses(vtheta, spec, x, oret.reshape(-1, 1), l2coeff=1e-3)
    _, og = vfo.vf_lossgrad(vtheta, spec, x, oret.reshape(-1, 1), l2coeff=1e-3)
    RESULTS["cat128_4M_vf"] = dict(losses=[float(v) for v in ls], oracle_losses=[float(v) for v in ols],
                                   grad_rel_l2=relerr(g, og), predict_rel_l2=relerr(base, pred[:, 0]))
    assert np.allclose(ls, ols, rtol=TOL)
    assert relerr(g, og) < TOL
