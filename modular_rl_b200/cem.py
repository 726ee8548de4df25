"""Noisy cross-entropy method over the flat parameter vector of a DeterministicAgent - the reference's
modular_rl/cem.py API (cem, CEM_OPTIONS, run_cem_algorithm), SURVEY section 8f rank 4.

Every population member is scored by one rollout with its theta loaded into the device-resident net
(`agent.set_from_flat`), so the per-step `act` is the same device forward the policy-gradient agents use.
The search arithmetic itself is a few numpy lines on the host, as in the reference."""
from __future__ import print_function

import numpy as np

from .core import rollout
from .misc_utils import update_default_config

try:
    from tabulate import tabulate
except ImportError:  # pragma: no cover
    def tabulate(rows):
        return "\n".join(" ".join("%10.4g" % v for v in r) for r in rows)


def cem(f, th_mean, batch_size, n_iter, elite_frac, initial_std=1.0, extra_std=0.0, std_decay_time=1.0, pool=None):
    """Generator of one info dict per iteration (cem.py:10-50): theta ~ Normal(th_mean, sample_std), keep the
    round(batch_size*elite_frac) best, refit.  As in the reference the running `th_std` holds the elite
    VARIANCE (and starts as ones*initial_std), and the sampling deviation is
    sqrt(th_std + extra_std^2 * max(1 - iteration/std_decay_time, 0))."""
    n_elite = int(np.round(batch_size * elite_frac))
    th_mean = np.asarray(th_mean, np.float64)
    th_std = np.ones(th_mean.size) * initial_std
    for iteration in range(n_iter):
        extra_var_multiplier = max(1.0 - iteration / float(std_decay_time), 0)
        print("extra var", extra_var_multiplier)
        sample_std = np.sqrt(th_std + np.square(extra_std) * extra_var_multiplier)
        ths = th_mean[None, :] + sample_std[None, :] * np.random.randn(batch_size, th_mean.size)
        ys = np.array(list(map(f, ths)) if pool is None else pool.map(f, ths))
        assert ys.ndim == 1
        elite_inds = ys.argsort()[-n_elite:]
        elite_ths = ths[elite_inds]
        th_mean = elite_ths.mean(axis=0)
        th_std = elite_ths.var(axis=0)
        yield {"ys": ys, "th": th_mean, "ymean": ys.mean(), "std": sample_std}


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
    """cem.py:63-98.  `parallel` is accepted and ignored: the reference forks a process pool around a CPU
    Theano function; a forked child cannot share this process's CUDA context, and one GPU serves the
    population sequentially faster than the hosts' cores run the environment."""
    cfg = update_default_config(CEM_OPTIONS, usercfg)
    if cfg["std_decay_time"] < 0:
        cfg["std_decay_time"] = cfg["n_iter"] / 2
    if usercfg:
        cfg.update(usercfg)
    print("cem config", {k: cfg[k] for (k, _, _, _) in CEM_OPTIONS})
    if cfg["parallel"]:
        print("parallel=1: evaluating the population sequentially on the device")
    timestep_limit = cfg["timestep_limit"]

    def objective(th):
        agent.set_from_flat(th)
        path = rollout(env, agent, timestep_limit)
        return path["reward"].sum()

    th_mean = agent.get_flat()
    for info in cem(objective, th_mean, cfg["batch_size"], cfg["n_iter"], cfg["elite_frac"],
                    cfg["initial_std"], cfg["extra_std"], cfg["std_decay_time"]):
        if callback is not None:
            callback(info)
        ps = np.linspace(0, 100, 5)
        print(tabulate([ps, np.percentile(info["ys"].ravel(), ps), np.percentile(info["std"].ravel(), ps)]))
        agent.set_from_flat(info["th"])
