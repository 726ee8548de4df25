#!/usr/bin/env python
"""Load a snapshotted agent and replay its behaviour - the reference's sim_agent.py (positional snapshot file,
--timestep_limit, --snapname), reading either the hdf5 results file of `--use_hdf 1` or the .pkl snapshots
run_pg.py / run_cem.py write without h5py (a snapshot file, or the <outfile>.dir directory).

  python sim_agent.py /tmp/a.h5.dir --episodes 3

The agent acts deterministically (sim_agent.py:53) through its device-resident policy.  --episodes N stops after
N rollouts instead of waiting for the enter key between them; --env overrides the stored environment id.
"""
import argparse
import time
from collections import defaultdict

import numpy as np


def animate_rollout(env, agent, n_timesteps, delay=.01):
    """sim_agent.py:11-31: one rendered rollout; returns the per-step infos, observations, rewards, actions."""
    infos = defaultdict(list)
    ob = env.reset()
    if hasattr(agent, "reset"):
        agent.reset()
    env.render()
    for i in range(n_timesteps):
        ob = agent.obfilt(ob)
        a, _info = agent.act(ob)
        (ob, rew, done, info) = env.step(a)
        env.render()
        if done:
            print("terminated after %s timesteps" % i)
            break
        for (k, v) in info.items():
            infos[k].append(v)
        infos['ob'].append(ob)
        infos['reward'].append(rew)
        infos['action'].append(a)
        if delay:
            time.sleep(delay)
    return infos


def main(argv=None):
    from modular_rl_b200.envs import make
    from modular_rl_b200.misc_utils import load_agent_snapshot, snapshot_env_id
    parser = argparse.ArgumentParser()
    parser.add_argument("hdf", help="hdf5 results file, snapshot .pkl, or the <outfile>.dir directory")
    parser.add_argument("--timestep_limit", type=int)
    parser.add_argument("--snapname")
    parser.add_argument("--env", help="environment id (default: the one stored with the snapshots)")
    parser.add_argument("--episodes", type=int, default=0, help="stop after this many rollouts (0: ask after each)")
    parser.add_argument("--delay", type=float, default=None, help="seconds between frames (default 1/fps)")
    args = parser.parse_args(argv)

    env_id = args.env or snapshot_env_id(args.hdf)
    if env_id is None:
        raise ValueError("no environment id stored with %s; pass --env" % args.hdf)
    env = make(env_id)
    agent = load_agent_snapshot(args.hdf, args.snapname)
    agent.stochastic = False

    timestep_limit = args.timestep_limit or env.spec.timestep_limit
    fps = getattr(env, "metadata", {}).get('video.frames_per_second', 30)
    delay = 1.0 / fps if args.delay is None else args.delay
    totals = []
    while True:
        infos = animate_rollout(env, agent, n_timesteps=timestep_limit, delay=delay)
        for (k, v) in infos.items():
            if k.startswith("reward"):
                print("%s: %f" % (k, np.sum(v)))
        totals.append(float(np.sum(infos["reward"])))
        if args.episodes:
            if len(totals) >= args.episodes:
                break
        else:
            input("press enter to continue")
    return totals


if __name__ == "__main__":
    main()
