"""Noisy cross-entropy method over the flat parameter vector of a DeterministicAgent - the reference's
modular_rl/cem.py API (cem, CEM_OPTIONS, run_cem_algorithm), SURVEY section 8f rank 4.

Every population member is scored by one rollout with its theta loaded into the device-resident net
(`agent.set_from_flat`), so the per-step `act` is the same device forward the policy-gradient agents use.
The search arithmetic itself is a few numpy lines on the host, as in the reference."""
from __future__ import print_function

import numpy as np

from .core import _filter_block, rollout
from .misc_utils import update_default_config

try:
    from tabulate import tabulate
except ImportError:  # pragma: no cover
    def tabulate(rows):
        return "\n".join(" ".join("%10.4g" % v for v in r) for r in rows)


def evaluate_population(env, agent, ths, timestep_limit):
    """Total reward of one rollout per candidate theta, all candidates in lockstep: per step ONE launch evaluates every
    live member's net on its own observation (`mrl_population_forward`) - the population-batched forward of SURVEY
    section 8f rank 4; the reference scores the candidates one rollout after another (cem.py:43-44,88-91).  The
    environments are copies of `env`; the agent's filters see the observations of a step as one block."""
    import copy
    from .core import Categorical
    from .device import population_forward
    pol = agent.policy
    ths = np.asarray(ths, np.float32)
    M = ths.shape[0]
    envs = [copy.deepcopy(env) for _ in range(M)]
    obs = [e.reset() for e in envs]
    total = np.zeros(M)
    live = list(range(M))
    discrete = isinstance(pol.probtype, Categorical)
    for _ in range(timestep_limit):
        if not live:
            break
        block = _filter_block(agent.obfilter, np.stack([np.asarray(obs[i], np.float64) for i in live]))
        out = population_forward(pol.dims, pol.activation, ths[live], block)
        acts = out.argmax(axis=1) if discrete else out          # maxprob of the deterministic policy (core.py:364-365,437-438)
        rews, still = [], []
        for j, i in enumerate(live):
            ob, rew, done, _ = envs[i].step(acts[j])
            obs[i] = ob
            total[i] += rew
            rews.append(rew)
            if not done:
                still.append(i)
        _filter_block(agent.rewfilter, np.asarray(rews, np.float64))
        live = still
    return total


class _GaussianSearch(object):
    """Diagonal-Gaussian search distribution of the noisy cross-entropy method (cem.py:10-50).  `var` is the elite
    VARIANCE (the reference's `th_std`, which starts as ones * initial_std); a generation is drawn with deviation
    sqrt(var + extra_std^2 * max(1 - generation / std_decay_time, 0)).  One `np.random.randn(batch, dim)` call per
    generation keeps the reference's random stream (pinned bit-exactly by tests/golden/cem_vectors.npz)."""

    def __init__(self, mean, initial_std, extra_std, std_decay_time):
        self.mean = np.asarray(mean, np.float64)
        self.var = np.full(self.mean.size, 1.0) * initial_std
        self.extra_var = np.square(extra_std)
        self.decay = float(std_decay_time)
        self.generation = 0

    def noise_left(self):
        return max(1.0 - self.generation / self.decay, 0)

    def draw(self, count):
        self.dev = np.sqrt(self.var + self.extra_var * self.noise_left())
        return self.mean[None, :] + self.dev[None, :] * np.random.randn(count, self.mean.size)

    def refit(self, candidates, scores, keep):
        best = candidates[scores.argsort()[-keep:]]
        self.mean, self.var = best.mean(axis=0), best.var(axis=0)
        self.generation += 1


def _score(f, ths, pool):
    population = getattr(f, "population", None)
    if population is not None:
        return np.asarray(population(ths))                      # all candidates in one lockstep evaluation
    return np.array([f(th) for th in ths] if pool is None else pool.map(f, ths))


def cem(f, th_mean, batch_size, n_iter, elite_frac, initial_std=1.0, extra_std=0.0, std_decay_time=1.0, pool=None):
    """Generator of one info dict per iteration - the reference's `cem` signature and info keys (cem.py:10-50):
    keep the round(batch_size * elite_frac) best of each generation and refit the search distribution to them.  An
    objective that has a `population` attribute is scored with one call for the whole generation."""
    search = _GaussianSearch(th_mean, initial_std, extra_std, std_decay_time)
    keep = int(np.round(batch_size * elite_frac))
    for _ in range(n_iter):
        print("extra var", search.noise_left())
        ths = search.draw(batch_size)
        ys = _score(f, ths, pool)
        assert ys.ndim == 1
        search.refit(ths, ys, keep)
        yield {"ys": ys, "th": search.mean, "ymean": ys.mean(), "std": search.dev}


CEM_OPTIONS = [
    ("batch_size", int, 200, "Number of episodes per batch"),
    ("n_iter", int, 200, "Number of iterations"),
    ("elite_frac", float, 0.2, "fraction of parameter settings used to fit pop"),
    ("initial_std", float, 1.0, "initial standard deviation for parameters"),
    ("extra_std", float, 0.0, "extra stdev added"),
    ("std_decay_time", float, -1.0, "number of timesteps that extra stdev decays over. negative => n_iter/2"),
    ("timestep_limit", int, 0, "maximum length of trajectories"),
    ("parallel", int, 0, "collect trajectories in parallel"),
]


def run_cem_algorithm(env, agent, usercfg=None, callback=None):
    """cem.py:63-98.  `parallel=1`: the reference forks a process pool around a CPU Theano function (every worker with
    its own copy of the agent and its filters); a forked child cannot share this process's CUDA context, so here
    parallel=1 scores the whole population in lockstep with one population-batched device forward per environment step
    (`evaluate_population`).  parallel=0 scores one candidate after another as cem.py:43 does."""
    cfg = update_default_config(CEM_OPTIONS, usercfg)
    if cfg["std_decay_time"] < 0:
        cfg["std_decay_time"] = cfg["n_iter"] / 2
    if usercfg:
        cfg.update(usercfg)
    print("cem config", {k: cfg[k] for (k, _, _, _) in CEM_OPTIONS})
    if cfg["parallel"]:
        print("parallel=1: population-batched evaluation (one device forward per step for all candidates)")
    timestep_limit = cfg["timestep_limit"]

    def objective(th):
        agent.set_from_flat(th)
        path = rollout(env, agent, timestep_limit)
        return path["reward"].sum()
    if cfg["parallel"]:
        objective.population = lambda ths: evaluate_population(env, agent, ths, timestep_limit)

    th_mean = agent.get_flat()
    for info in cem(objective, th_mean, cfg["batch_size"], cfg["n_iter"], cfg["elite_frac"],
                    cfg["initial_std"], cfg["extra_std"], cfg["std_decay_time"]):
        if callback is not None:
            callback(info)
        ps = np.linspace(0, 100, 5)
        print(tabulate([ps, np.percentile(info["ys"].ravel(), ps), np.percentile(info["std"].ravel(), ps)]))
        agent.set_from_flat(info["th"])
