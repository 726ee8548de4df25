#!/usr/bin/env python
"""This script runs the cross-entropy method - same flags as the reference's run_cem.py (GENERAL_OPTIONS +
--env --agent --plot + the agent's options + CEM_OPTIONS, run_cem.py:13-30), with every population member
evaluated through the device-resident deterministic policy.

  python run_cem.py --env CartPole-v0 --agent modular_rl.agentzoo.DeterministicAgent --n_iter 10 --batch_size 40
"""
import argparse
import os
import pickle
import shutil
import sys

import numpy as np

from modular_rl import *  # noqa: F401,F403
from modular_rl_b200.envs import make

try:
    from tabulate import tabulate
except ImportError:  # pragma: no cover
    def tabulate(rows):
        return "\n".join("%-24s %s" % (k, v) for k, v in rows)


def main():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    update_argument_parser(parser, GENERAL_OPTIONS)
    parser.add_argument("--env", required=True)
    parser.add_argument("--agent", required=True)
    parser.add_argument("--plot", action="store_true")
    args, _ = parser.parse_known_args([arg for arg in sys.argv[1:] if arg not in ('-h', '--help')])
    env = make(args.env)
    env_spec = env.spec
    # read the snapshot before the results directory of a previous run (which may hold it) is cleared
    snapshot_agent = load_agent_snapshot(args.load_snapshot) if args.load_snapshot else None
    mondir = args.outfile + ".dir"
    if os.path.exists(mondir):
        shutil.rmtree(mondir)
    os.makedirs(mondir)
    agent_ctor = get_agent_cls(args.agent)
    update_argument_parser(parser, agent_ctor.options)
    update_argument_parser(parser, CEM_OPTIONS)
    args = parser.parse_args()
    if args.timestep_limit == 0:
        args.timestep_limit = env_spec.max_episode_steps
    cfg = args.__dict__
    np.random.seed(args.seed)
    agent = snapshot_agent if args.load_snapshot else agent_ctor(env.observation_space, env.action_space, cfg)
    if args.use_hdf:
        hdf, diagnostics = prepare_h5_file(args)

    counter = [0]

    def callback(stats):
        if args.use_hdf:
            for (stat, val) in stats.items():
                diagnostics[stat].append(val)
        if args.plot:
            animate_rollout(env, agent, min(500, args.timestep_limit))
        print("*********** Iteration %i ****************" % counter[0])
        print(tabulate([(k, v) for k, v in stats.items() if np.asarray(v).size == 1]))
        counter[0] += 1
        if args.snapshot_every and ((counter[0] % args.snapshot_every == 0) or (counter[0] == args.n_iter)):
            agent.set_from_flat(stats["th"])
            if args.use_hdf:
                hdf['/agent_snapshots/%0.4i' % counter[0]] = np.array(pickle.dumps(agent, -1))
            else:
                save_agent_snapshot(agent, mondir, counter[0], env_id=env_spec.id)

    run_cem_algorithm(env, agent, callback=callback, usercfg=cfg)

    if args.use_hdf:
        hdf['env_id'] = env_spec.id
    env.close()


if __name__ == "__main__":
    main()
