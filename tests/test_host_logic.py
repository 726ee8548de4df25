"""CPU tests of the host-side logic: option system / flags, the filter and running-stat mirrors
against the reference's golden vectors, sharding + merge formulas, and the N>1 data-parallel
algebra with world_size-2 gloo (per-shard sums from the oracle stand in for the device kernels)."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_running_stat_and_zfilter_mirror_reference(golden):
    from modular_rl_b200.filters import ZFilter
    from modular_rl_b200.running_stat import RunningStat
    X = golden["rs_x"]
    rs = RunningStat((3,))
    for t, x in enumerate(X):
        rs.push(x)
        assert np.array_equal(rs.mean, golden["rs_mean"][t])
        assert np.array_equal(rs.var, golden["rs_var"][t])
    zf = ZFilter((3,), clip=5)
    assert np.array_equal(np.array([zf(x) for x in X]), golden["zf_ob_y"])
    assert np.array_equal(zf(X[0], update=False), golden["zf_noupdate_y"])
    zr = ZFilter((), demean=False, clip=10)
    assert np.array_equal(np.array([zr(x) for x in golden["zf_rew_x"]]), golden["zf_rew_y"])
    # the reference's own unit test (running_stat.py:35-46)
    for shp in ((), (3,), (3, 4)):
        li, rs = [], RunningStat(shp)
        for _ in range(5):
            val = np.random.randn(*shp)
            rs.push(val)
            li.append(val)
            m = np.mean(li, axis=0)
            assert np.allclose(rs.mean, m)
            v = np.square(m) if (len(li) == 1) else np.var(li, ddof=1, axis=0)
            assert np.allclose(rs.var, v)


def test_option_tables_and_flags_match_reference():
    from modular_rl_b200 import agentzoo, misc_utils
    names = [o[0] for o in agentzoo.TrpoAgent.options]
    assert names == ["hid_sizes", "activation", "timestep_limit", "n_iter", "parallel", "timesteps_per_batch",
                     "gamma", "lam", "cg_damping", "max_kl", "filter"]
    d = {o[0]: o[2] for o in agentzoo.TrpoAgent.options}
    assert (d["cg_damping"], d["max_kl"], d["gamma"], d["lam"], d["timesteps_per_batch"]) == (1e-3, 1e-2, 0.99, 1.0, 100)
    names = [o[0] for o in agentzoo.PpoLbfgsAgent.options]
    assert names[6:-1] == ["gamma", "lam", "kl_target", "maxiter", "reverse_kl", "do_split"]
    parser = argparse.ArgumentParser()
    misc_utils.update_argument_parser(parser, misc_utils.GENERAL_OPTIONS)
    misc_utils.update_argument_parser(parser, agentzoo.TrpoAgent.options)
    args = parser.parse_args(["--hid_sizes", "10,5", "--cg_damping", "0.1", "--lam", "0.97", "--seed", "3"])
    assert args.hid_sizes == [10, 5] and list(args.hid_sizes) == [10, 5]     # reusable, not a one-shot map
    assert args.cg_damping == 0.1 and args.seed == 3 and args.max_kl == 1e-2
    cfg = misc_utils.update_default_config(agentzoo.TrpoAgent.options, dict(max_kl=0.5, unknown=1))
    assert cfg["max_kl"] == 0.5 and "unknown" not in cfg and cfg.max_kl == 0.5
    with pytest.raises(ValueError):
        misc_utils.update_argument_parser(argparse.ArgumentParser(), [], bogus=1)


def test_misc_helpers(golden):
    from modular_rl_b200 import distributions, misc_utils
    assert np.allclose(misc_utils.explained_variance_2d(golden["ev_yp"], golden["ev_y"]), golden["ev"], rtol=1e-14)
    arrs = [np.arange(6.).reshape(2, 3), np.arange(4.)]
    flat = misc_utils.flatten(arrs)
    back = misc_utils.unflatten(flat, [a.shape for a in arrs])
    assert all(np.array_equal(a, b) for a, b in zip(arrs, back))
    np.random.seed(7)
    assert np.array_equal(distributions.categorical_sample(golden["c_p0"]), golden["c_sample_seed7"])
    assert np.allclose(distributions.categorical_kl(golden["c_p0"], golden["c_p1"]), golden["c_kl"])
    assert np.allclose(distributions.categorical_entropy(golden["c_p0"]), golden["c_ent"])


def test_probtype_array_methods(golden):
    from modular_rl_b200.core import Categorical, DiagGauss
    g = DiagGauss(4)
    assert np.allclose(g.loglikelihood(golden["g_a"], golden["g_p0"]), golden["g_loglik"], rtol=1e-14)
    assert np.allclose(g.kl(golden["g_p0"], golden["g_p1"]), golden["g_kl"], rtol=1e-14)
    assert np.allclose(g.entropy(golden["g_p0"]), golden["g_ent"], rtol=1e-14)
    c = Categorical(3)
    assert np.array_equal(c.likelihood(golden["c_a"], golden["c_p0"]), golden["c_lik"])
    assert np.allclose(c.kl(golden["c_p0"], golden["c_p1"]), golden["c_kl"], rtol=1e-14)
    assert np.array_equal(c.maxprob(golden["c_p0"]), golden["c_p0"].argmax(1))


def test_cg_and_linesearch_signatures_match_reference(golden):
    from modular_rl_b200.trpo import cg, linesearch
    A, b = golden["cg_A"], golden["cg_b"]
    assert np.array_equal(cg(lambda p: A @ p, b), golden["cg_x10"])
    assert np.array_equal(cg(lambda p: A @ p, b, cg_iters=3), golden["cg_x3"])
    assert np.array_equal(cg(lambda p: golden["cg_A2"] @ p, b), golden["cg_x_early"])
    Q, x0 = golden["ls_Q"], golden["ls_x0"]
    for name in ("ls_a", "ls_b", "ls_c"):
        ok, xn = linesearch(lambda x: 0.5 * x @ Q @ x, x0, golden[name + "_full"], float(golden[name + "_rate"]))
        assert ok == bool(golden[name + "_ok"]) and np.array_equal(xn, golden[name + "_x"])


def test_envs_and_spaces():
    from modular_rl_b200 import envs
    env = envs.make("CartPole-v0")
    np.random.seed(0)
    ob = env.reset()
    assert ob.shape == (4,) and env.action_space.n == 2 and env.spec.max_episode_steps == 200
    tot = 0
    for _ in range(500):
        ob, r, done, _ = env.step(env.action_space.sample())
        tot += r
        if done:
            break
    assert 5 <= tot < 200
    pend = envs.make("Pendulum-v0")
    ob = pend.reset()
    ob, r, done, _ = pend.step(np.array([0.3]))
    assert ob.shape == (3,) and r <= 0 and not done and pend.action_space.shape == (1,)
    with pytest.raises(KeyError):
        envs.make("NoSuchEnv-v9")


# ----------------------------------------------------------------------------- sharding
def test_shard_bounds_cover_whole_paths():
    from modular_rl_b200.parallel import shard_bounds
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for n_paths in (1, 3, 50, 1000):
            lens = rng.integers(1, 400, n_paths)
            b = shard_bounds(lens, world)
            assert b[0][0] == 0 and b[-1][1] == n_paths
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [int(lens[a:e].sum()) for a, e in b]
            assert sum(sizes) == lens.sum()
            if n_paths >= 50 * world:
                assert max(sizes) - min(sizes) <= 2 * lens.max()      # balanced to within a path or two


def test_merge_moments_and_zfilter_prefix():
    from modular_rl_b200.parallel import merge_moments, zfilter_prefix
    from oracle import zfilter as ozf
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1000) * 3 + 7
    parts = np.split(x, [100, 101, 640])
    tr = [(len(p), p.mean(), ((p - p.mean()) ** 2).sum()) for p in parts] + [(0, 0.0, 0.0)]
    n, mean, m2 = merge_moments(tr)
    assert n == 1000 and np.isclose(mean, x.mean(), rtol=1e-14) and np.isclose(np.sqrt(m2 / n), x.std(), rtol=1e-13)
    X = rng.standard_normal((300, 4)) * 2 + 1
    blocks = [X[:50], X[50:200], X[200:]]
    states = []
    for blk in blocks:
        st = ozf.WelfordState((4,))
        for row in blk:
            st.push(row)
        states.append((st.n, st.M.copy(), st.S.copy()))
    full = ozf.WelfordState((4,))
    for row in X[:200]:
        full.push(row)
    n, M, S = zfilter_prefix(states, 2)
    assert n == 200 and np.allclose(M, full.M, rtol=1e-13) and np.allclose(S, full.S, rtol=1e-12)


# ----------------------------------------------------------------------------- gloo, world_size 2
def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from modular_rl_b200 import synth
    from modular_rl_b200.parallel import merge_moments, shard_bounds
    from oracle import advantage as oadv, natgrad, policy_math as pm
    wl = synth.Workload("dp", (7, 16, 8, 3), synth.GAUSS, 900, 60, 3)
    spec = pm.NetSpec(wl.dims, pm.GAUSS)
    data = synth.policy_batch(wl, lambda th, ob: pm.forward(th, spec, ob)[1])
    theta = synth.perturb(data["theta"], 0.03, 2)
    off = data["offsets"]
    a, b = shard_bounds(np.diff(off), world)[rank]
    lo, hi = int(off[a]), int(off[b])
    N = int(off[-1])
    sl = slice(lo, hi)

    def allsum(x):
        t = torch.from_numpy(np.ascontiguousarray(x, np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    # GAE is per trajectory -> shards need no exchange; standardisation merges (n, mean, M2)
    base = np.tanh(data["ob"][:, 0])
    ret, adv = oadv.gae_flat(data["reward"][sl], base[sl], off[a:b + 1] - lo, data["terminated"][a:b], 0.99, 0.95)
    fret, fadv = oadv.gae_flat(data["reward"], base, off, data["terminated"], 0.99, 0.95)
    assert np.array_equal(ret, fret[sl]) and np.array_equal(adv, fadv[sl])
    slot = np.zeros(3 * world)
    slot[3 * rank:3 * rank + 3] = [adv.size, adv.mean(), ((adv - adv.mean()) ** 2).sum()]
    gathered = allsum(slot).reshape(world, 3)
    n, mean, m2 = merge_moments([tuple(r) for r in gathered])
    sadv = (adv - mean) / np.sqrt(m2 / n)
    assert np.allclose(sadv, oadv.standardize(fadv)[sl], rtol=1e-12, atol=1e-12)

    # gradient / Fvp / losses: local sums scaled by 1/N_global, then one sum over ranks
    args = (data["ob"][sl], data["act"][sl], data["adv"][sl], data["oldprob"][sl])
    n_loc = hi - lo
    g = allsum(pm.policy_gradient(theta, spec, *args) * n_loc / N)
    assert np.allclose(g, pm.policy_gradient(theta, spec, data["ob"], data["act"], data["adv"], data["oldprob"]),
                       rtol=1e-10, atol=1e-14)
    v = np.random.default_rng(0).standard_normal(theta.size).astype(np.float32)
    f_loc = pm.fisher_vector_product(theta, spec, data["ob"][sl], v) * n_loc / N
    d = wl.dims[-1]
    f_loc[-d:] = 2.0 * v[-d:] / world          # data-independent logstd block: split so the sum restores it
    f = allsum(f_loc)
    assert np.allclose(f, pm.fisher_vector_product(theta, spec, data["ob"], v), rtol=1e-10, atol=1e-14)
    ls = allsum(pm.losses(theta, spec, *args) * n_loc / N)
    assert np.allclose(ls, pm.losses(theta, spec, data["ob"], data["act"], data["adv"], data["oldprob"]), rtol=1e-11)

    # replicated CG on all-reduced Fvps gives every rank the same bits
    def fvp(p):
        fl = pm.fisher_vector_product(theta, spec, data["ob"][sl], p) * n_loc / N
        fl[-d:] = 2.0 * np.asarray(p, np.float32)[-d:] / world
        return allsum(fl) + 0.1 * p
    x, it, _ = natgrad.conjugate_gradient(fvp, -g)
    ref, rit, _ = natgrad.conjugate_gradient(
        lambda p: pm.fisher_vector_product(theta, spec, data["ob"], p) + 0.1 * p, -g)
    assert it == rit and np.allclose(x, ref, rtol=1e-8, atol=1e-12)
    np.save(os.path.join(tmp, f"x{rank}.npy"), x)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_algebra_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    x0, x1 = np.load(tmp_path / "x0.npy"), np.load(tmp_path / "x1.npy")
    assert np.array_equal(x0, x1)          # bit-identical replicas


# ----------------------------------------------------------------------------- PpoSgdUpdater oracle (ppo.py:115-258)
def test_oracle_adam_matches_independent_adam():
    """adam_updates' recurrences (ppo.py:231-258) against torch.optim.Adam; the two differ only in where epsilon
    sits, so they must agree to rounding with epsilon = 0."""
    import torch
    from oracle.ppo_sgd import Adam
    rng = np.random.default_rng(0)
    th = rng.standard_normal(40)
    A = rng.standard_normal((40, 40))
    A = A @ A.T / 40
    o = Adam(40, 1e-2, epsilon=0.0)
    t = torch.tensor(th.copy(), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([t], lr=1e-2, eps=0.0)
    x = th.copy()
    for _ in range(15):
        x = o.step(x, A @ x)
        opt.zero_grad()
        (0.5 * (t @ torch.tensor(A) @ t)).backward()
        opt.step()
    assert np.abs(x - t.detach().numpy()).max() < 1e-13
    # with the reference's epsilon the first step is lr * g / (|g| + eps * ...) ~ lr * sign(g)
    o2 = Adam(3, 1e-3)
    x2 = o2.step(np.zeros(3), np.array([2.0, -0.5, 1e-3]))
    assert np.allclose(x2, [-1e-3, 1e-3, -1e-3], rtol=1e-3)
    want = -1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * 1e-3) / (np.sqrt(0.001) * 1e-3 + 1e-8)   # eps is not bias-corrected
    assert np.isclose(x2[2], want, rtol=1e-12)


def test_oracle_ppo_sgd_update_runs_and_adapts_kl_coeff():
    from oracle import policy_math as pm
    from oracle.ppo_sgd import Adam, ppo_sgd_update
    from modular_rl_b200 import synth
    rng = np.random.default_rng(3)
    dims = (5, 8, 3)
    spec = pm.NetSpec(dims, pm.GAUSS)
    theta = synth.init_params(dims, synth.GAUSS, rng)
    ob = rng.standard_normal((300, 5))
    prob = np.concatenate([pm.forward(theta, spec, ob)[1], np.ones((300, 3))], 1)
    act = prob[:, :3] + rng.standard_normal((300, 3))
    adv = rng.standard_normal(300)
    np.random.seed(1)
    info, th_new, klc, n_mb = ppo_sgd_update(theta, spec, ob, act, adv, Adam(theta.size), epochs=3)
    assert n_mb == 3 * 3 and th_new.shape == theta.shape and not np.allclose(th_new, theta)
    assert list(info)[:3] == ["surr_before", "surr_after", "surr_change"]
    assert abs(info["kl_before"]) < 1e-12 and info["kl_after"] > 0          # old net == net at entry
    assert klc in (1.5, 1.0, 1.0 / 1.5)
    np.random.seed(1)
    info2, *_ = ppo_sgd_update(theta, spec, ob, act, adv, Adam(theta.size), epochs=3, do_split=True)
    assert "test_kl_after" in info2


def test_snapshot_files_roundtrip(tmp_path):
    """save_agent_snapshot / load_agent_snapshot: one pickle per snapshot, named like the reference's hdf5
    keys (run_pg.py:141-142); a directory resolves to its last snapshot or to a named one (sim_agent.py:41-52)."""
    from modular_rl_b200.misc_utils import load_agent_snapshot, save_agent_snapshot
    from modular_rl_b200.filters import ZFilter
    f = ZFilter((3,), clip=5)
    for i in range(4):
        f(np.arange(3.0) * i)
    p1 = save_agent_snapshot(f, str(tmp_path), 20)
    f(np.ones(3))
    save_agent_snapshot(f, str(tmp_path), 40)
    assert p1.endswith("agent_snapshots/0020.pkl")
    assert load_agent_snapshot(p1).rs.n == 4
    assert load_agent_snapshot(str(tmp_path)).rs.n == 5
    assert load_agent_snapshot(str(tmp_path), "0020").rs.n == 4
    with pytest.raises(ValueError):
        load_agent_snapshot(str(tmp_path), "0030")


def test_cem_matches_reference_golden():
    """modular_rl_b200.cem.cem against the vectors the reference's own generator produced
    (tests/golden/make_golden.py cem): same numpy random stream, same elite selection, bit-exact."""
    import contextlib
    import io
    from modular_rl_b200.cem import cem
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cem_vectors.npz"))
    c = G["cem_center"]
    f = lambda th: -np.sum((th - c) ** 2) + 0.1 * np.sin(th).sum()
    np.random.seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        infos = list(cem(f, np.zeros(7, np.float32), 40, 6, 0.2, initial_std=1.5, extra_std=0.4, std_decay_time=3.0))
    np.testing.assert_array_equal(np.array([i["ys"] for i in infos]), G["cem_ys"])
    np.testing.assert_array_equal(np.array([i["th"] for i in infos]), G["cem_th"])
    np.testing.assert_array_equal(np.array([i["std"] for i in infos]), G["cem_std"])
    np.testing.assert_array_equal(np.array([i["ymean"] for i in infos]), G["cem_ymean"])
    assert infos[-1]["ymean"] > infos[0]["ymean"]


class _PushLeftRightAgent(object):
    """Stand-in agent for the sim_agent test: picklable, no device."""
    stochastic = True

    def obfilt(self, ob):
        return ob

    def act(self, ob):
        return int(ob[2] > 0), {}          # push towards the side the pole leans to


def test_sim_agent_replays_snapshot(tmp_path, capsys):
    """sim_agent.py: snapshot directory + stored environment id -> deterministic replays with reward totals."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sim_agent
    from modular_rl_b200.misc_utils import save_agent_snapshot, snapshot_env_id
    save_agent_snapshot(_PushLeftRightAgent(), str(tmp_path), 7, env_id="CartPole-v0")
    assert snapshot_env_id(str(tmp_path)) == "CartPole-v0"
    assert snapshot_env_id(str(tmp_path / "agent_snapshots" / "0007.pkl")) == "CartPole-v0"
    np.random.seed(0)
    totals = sim_agent.main([str(tmp_path), "--episodes", "2", "--delay", "0", "--timestep_limit", "60"])
    assert len(totals) == 2 and all(t > 5 for t in totals)
    assert "reward:" in capsys.readouterr().out
    with pytest.raises(ValueError):
        sim_agent.main([str(tmp_path), "--snapname", "0001", "--episodes", "1", "--delay", "0"])


# ----------------------------------------------------------------------------- vectorised rollouts / CEM host logic
class _HostFilter(object):
    """ZFilter through its per-sample host path only (no filter_batch attribute -> _filter_block goes row by row)."""

    def __init__(self, shape, **kw):
        from modular_rl_b200.filters import ZFilter
        self.z = ZFilter(shape, **kw)

    def __call__(self, x, update=True):
        return self.z(x, update)


class _StubPolicy(object):
    """A linear softmax policy in numpy with the act / act_batch pair of StochPolicy (core.py:261-267)."""

    def __init__(self, w):
        self.w = w

    def _probs(self, X):
        z = X @ self.w
        p = np.exp(z - z.max(axis=1, keepdims=True))
        return p / p.sum(axis=1, keepdims=True)

    def act(self, ob, stochastic=True):
        a, info = self.act_batch(ob[None], stochastic)
        return a[0], {"prob": info["prob"][0]}

    def act_batch(self, X, stochastic=True):
        p = self._probs(np.asarray(X, np.float64))
        u = np.random.rand(len(p), 1)
        a = np.argmax(np.cumsum(p, axis=1) > u, axis=1) if stochastic else p.argmax(axis=1)
        return a, {"prob": p}


class _StubAgent(object):
    stochastic = True

    def __init__(self, seed):
        self.policy = _StubPolicy(np.random.default_rng(seed).standard_normal((4, 2)))
        self.obfilter = _HostFilter((4,), clip=5)
        self.rewfilter = _HostFilter((), demean=False, clip=10)

    def act(self, ob):
        return self.policy.act(ob, self.stochastic)

    def obfilt(self, ob):
        return self.obfilter(ob)

    def rewfilt(self, rew):
        return self.rewfilter(rew)


def test_rollouts_vectorized_host_logic():
    """core.rollouts_vectorized: with one environment it is `rollout` (same filter updates, same numpy draws); with
    several it returns one path per environment with the reference's keys, and do_rollouts_vectorized keeps the
    strict 'more than n_timesteps' stopping rule of core.py:219."""
    import itertools
    from modular_rl_b200 import core, envs
    env_a, env_b = envs.make("CartPole-v0"), envs.make("CartPole-v0")
    ag_a, ag_b = _StubAgent(3), _StubAgent(3)
    np.random.seed(11)
    ref = core.rollout(env_a, ag_a, 200)
    np.random.seed(11)
    (vec,) = core.rollouts_vectorized([env_b], ag_b, 200)
    assert set(vec) == set(ref) and vec["terminated"] == ref["terminated"]
    for k in ("observation", "action", "reward", "prob"):
        np.testing.assert_array_equal(vec[k], ref[k])
    assert ag_a.obfilter.z.rs.n == ag_b.obfilter.z.rs.n and np.array_equal(ag_a.obfilter.z.rs.mean, ag_b.obfilter.z.rs.mean)
    # three environments in lockstep
    np.random.seed(5)
    many = [envs.make("CartPole-v0") for _ in range(3)]
    paths = core.rollouts_vectorized(many, _StubAgent(4), 60)
    assert len(paths) == 3
    for p in paths:
        T = len(p["reward"])
        assert 1 <= T <= 60 and p["observation"].shape == (T, 4) and p["prob"].shape == (T, 2) and p["action"].shape == (T,)
        assert p["terminated"] == (T < 60)
    seeds = itertools.count()
    got = core.do_rollouts_vectorized(envs.make("CartPole-v0"), _StubAgent(6), 50, 120, seeds, 4)
    total = sum(len(p["reward"]) for p in got)
    assert len(got) % 4 == 0 and total > 120 and total - sum(len(p["reward"]) for p in got[-4:]) <= 120


def test_cem_population_hook_and_numa_helper():
    """cem(): an objective with a `population` attribute is scored with one call per generation and gives the same
    search as the per-candidate loop; parallel.bind_to_device_numa never raises (0 = nothing bound)."""
    from modular_rl_b200.cem import cem
    from modular_rl_b200.parallel import bind_to_device_numa
    c = np.linspace(-1, 1, 5)
    f = lambda th: -float(np.sum((th - c) ** 2))
    calls = []

    def g(th):
        return f(th)
    g.population = lambda ths: (calls.append(len(ths)), np.array([f(t) for t in ths]))[1]
    np.random.seed(2)
    a = list(cem(f, np.zeros(5), 30, 5, 0.2))
    np.random.seed(2)
    b = list(cem(g, np.zeros(5), 30, 5, 0.2))
    assert calls == [30] * 5
    for ia, ib in zip(a, b):
        np.testing.assert_array_equal(ia["ys"], ib["ys"])
        np.testing.assert_array_equal(ia["th"], ib["th"])
    assert a[-1]["ymean"] > a[0]["ymean"]
    assert isinstance(bind_to_device_numa(0), int)
