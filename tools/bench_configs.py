#!/usr/bin/env python
"""Timings of BASELINE.json configs 2, 4 and 5 on one B200 (the headline config 3 is bench.py).

  config 2: Hopper-shaped 50k-timestep TRPO update
  config 4: PpoLbfgsUpdater penalised-KL update, Walker2d-shaped 200k timesteps (scipy L-BFGS-B on the
            host driving mrl_net_ppo_lossgrad)
  config 5: GAE + standardise + NnVf fit on a 4M-timestep ragged Categorical batch (obs 128, 18 actions)

Prints one JSON object; used to fill profiles/rNN_configs.json.  Inputs are resident on the device
unless stated; CUDA-event / perf_counter timing after warm-up.
"""
import json
import os
import sys
import time

import numpy as np
import scipy.optimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from modular_rl_b200 import _lib as L, synth  # noqa: E402
from modular_rl_b200.device import DeviceBatch, DeviceNet  # noqa: E402


def sync():
    torch.cuda.synchronize()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync()
    return (time.perf_counter() - t0) / reps


def setup(wl, N):
    rng = np.random.default_rng(wl.seed)
    theta = synth.init_params(wl.dims, wl.head, rng)
    ob = synth.make_obs(N, wl.dims[0], rng)
    offsets, terminated = synth.make_paths(N, wl.t_max, rng)
    net = DeviceNet(wl.dims, wl.head)
    batch = DeviceBatch(wl.dims[0], True)
    batch.set_obs(ob).set_paths(offsets, terminated, float(wl.t_max))
    net.set_params(theta)
    out = net.forward(batch)
    if wl.head == synth.GAUSS:
        d = wl.dims[-1]
        oldprob = np.concatenate([out, np.broadcast_to(np.exp(theta[-d:])[None], out.shape)], 1).astype(np.float32)
    else:
        oldprob = out
    act = synth.sample_actions(wl.head, oldprob, rng)
    adv = rng.standard_normal(N).astype(np.float32)
    adv = (adv - adv.mean()) / adv.std()
    batch.set_policy_inputs(wl.head, wl.dims[-1], act, adv, oldprob)
    theta_cur = synth.perturb(theta, 0.01, wl.seed + 7)
    reward = rng.standard_normal(N)
    return dict(net=net, batch=batch, theta=theta_cur, reward=reward, offsets=offsets, n_paths=len(terminated))


def config2():
    wl = synth.WORKLOADS["hopper"]
    s = setup(wl, wl.N)
    net, batch = s["net"], s["batch"]

    def step():
        net.set_params(s["theta"])
        net.trpo_step(batch, 0.1, 0.01)
    t = timeit(step, reps=20, warm=5)
    return {"workload": "Hopper 50k TRPO update (no GAE)", "ms_per_update": t * 1e3, "timesteps_per_s": wl.N / t}


def config4():
    wl = synth.WORKLOADS["walker2d"]
    s = setup(wl, wl.N)
    net, batch = s["net"], s["batch"]
    evals = [0]

    def lossandgrad(th):
        evals[0] += 1
        net.set_params(th)
        l, g, _ = net.ppo_lossgrad(batch, 1.0, 0.02)
        return l, g

    def update():
        net.set_params(s["theta"])
        scipy.optimize.fmin_l_bfgs_b(lossandgrad, s["theta"].astype(np.float64), maxiter=25)
    update()
    evals[0] = 0
    t = timeit(update, reps=3, warm=1)
    n_eval = evals[0] / 4
    t_eval = timeit(lambda: net.ppo_lossgrad(batch, 1.0, 0.02), reps=20, warm=3)
    return {"workload": "Walker2d 200k PpoLbfgs update, maxiter 25", "ms_per_update": t * 1e3,
            "lossgrad_evals_per_update": n_eval, "ms_per_lossgrad_eval": t_eval * 1e3,
            "timesteps_per_s_update": wl.N / t, "timestep_evals_per_s": wl.N / t_eval}


def config5():
    wl = synth.WORKLOADS["cat128"]
    N = wl.N
    s = setup(wl, N)
    batch = s["batch"]
    vdims = (wl.dims[0] + 1, 64, 64, 1)
    vf = DeviceNet(vdims, synth.VALUE)
    vtheta = synth.init_params(vdims, synth.VALUE, np.random.default_rng(9), last_scale=1.0)
    vf.set_params(vtheta)
    reward_d = torch.from_numpy(s["reward"]).cuda()

    def gae():
        vf.predict_into_baseline(batch)
        batch.gae(reward_d, None, 0.995, 0.97, True, None, want_outputs=False)
    t_gae = timeit(gae, reps=10, warm=3)

    lib = L.lib()
    import ctypes as C
    lib.mrl_profile_enable(1)
    for _ in range(10):
        batch.gae(reward_d, None, 0.995, 0.97, True, None, want_outputs=False)
    nk = lib.mrl_profile_kinds()
    ms, cnt = (C.c_double * nk)(), (C.c_longlong * nk)()
    lib.mrl_profile_read(ms, cnt)
    lib.mrl_profile_enable(0)
    names = [lib.mrl_profile_kind_name(k).decode() for k in range(nk)]
    k_gae = names.index("gae")
    gae_kernel_ms = ms[k_gae] / max(cnt[k_gae], 1)

    batch.mix_vf_target(0.1)
    evals = [0]

    def lossandgrad(th):
        evals[0] += 1
        vf.set_params(th)
        ls, g = vf.vf_lossgrad(batch, 1e-3)
        return ls[0], g

    def fit():
        vf.set_params(vtheta)
        scipy.optimize.fmin_l_bfgs_b(lossandgrad, vtheta.astype(np.float64), maxiter=25)
    fit()
    evals[0] = 0
    t_fit = timeit(fit, reps=2, warm=0)
    n_eval = evals[0] / 2
    t_eval = timeit(lambda: vf.vf_lossgrad(batch, 1e-3), reps=10, warm=2)
    return {"workload": "cat128 4M ragged: VF predict + GAE + standardise; NnVf fit maxiter 25",
            "n_paths": s["n_paths"], "ms_predict_gae_standardise": t_gae * 1e3,
            "gae_kernel_ms": gae_kernel_ms, "gae_kernel_GBps_at_24B_per_step": 24.0 * N / (gae_kernel_ms * 1e-3) / 1e9,
            "timesteps_per_s_gae_path": N / t_gae, "ms_vf_fit": t_fit * 1e3, "vf_lossgrad_evals": n_eval,
            "ms_per_vf_lossgrad_eval": t_eval * 1e3, "timestep_evals_per_s": N / t_eval}


if __name__ == "__main__":
    out = {"config2": config2(), "config4": config4(), "config5": config5()}
    print(json.dumps(out, indent=1))
