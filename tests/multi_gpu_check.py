#!/usr/bin/env python
"""Data-parallel parity check, run under torchrun with >= 2 ranks (one per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_check.py

Every rank builds the same seeded global batch, keeps its shard of WHOLE trajectories, runs
VF predict -> GAE -> standardise (moments merged over NCCL) -> TRPO step (gradient / Fvp / loss
sums all-reduced, CG replicated) and compares with the oracle on the full batch; ranks must end
with bit-identical parameters."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from modular_rl_b200 import synth
    from modular_rl_b200.device import DeviceBatch, DeviceNet
    from modular_rl_b200.parallel import comm_from_torch_distributed, shard_bounds
    from oracle import advantage as oadv, natgrad, policy_math as pm, valuefn

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = comm_from_torch_distributed(local, p2p=os.environ.get("MRL_NCCL_ONLY") != "1")
    if rank == 0:
        print("transport:", "nvlink peer-memory push" if comm.p2p else "nccl", flush=True)

    failures = []
    for name, dims, head in (("gauss", (23, 32, 16, 4), synth.GAUSS), ("cat", (10, 24, 24, 5), synth.CAT)):
        wl = synth.Workload(name, dims, head, 9000, 150, 21)
        spec = pm.NetSpec(dims, pm.GAUSS if head == synth.GAUSS else pm.CAT)

        def fwd(th, ob):
            _, z = pm.forward(th, spec, ob)
            return z if head == synth.GAUSS else pm.softmax(z)
        data = synth.policy_batch(wl, fwd)
        theta = synth.perturb(data["theta"], 0.02, 4)
        vdims = (dims[0] + 1, 16, 1)
        vspec = pm.NetSpec(vdims, pm.VALUE)
        vtheta = synth.init_params(vdims, synth.VALUE, np.random.default_rng(5), last_scale=1.0)
        off, term = data["offsets"], data["terminated"]
        N = int(off[-1])
        a, b = shard_bounds(np.diff(off), world)[rank]
        lo, hi = int(off[a]), int(off[b])
        sl = slice(lo, hi)

        net = DeviceNet(dims, head, device=local)
        vf = DeviceNet(vdims, synth.VALUE, device=local)
        net.set_comm(comm)
        batch = DeviceBatch(dims[0], True, device=local)
        batch.set_obs(data["ob"][sl]).set_paths(off[a:b + 1] - lo, term[a:b], 150.0)
        batch.set_global_n(N)
        net.set_params(theta)
        vf.set_params(vtheta)
        vf.predict_into_baseline(batch)
        ret, adv = batch.gae(data["reward"][sl], None, 0.99, 0.95, standardize=True, comm=comm)
        batch.set_policy_inputs(head, dims[-1], data["act"][sl], None, data["oldprob"][sl])
        stats, info = net.trpo_step(batch, cg_damping=0.1, max_kl=0.01)
        th_new = net.get_params()

        # ---- oracle on the full batch
        t_idx, _ = oadv.time_index(off)
        x = np.concatenate([data["ob"], (t_idx / 150.0)[:, None]], axis=1)
        base = valuefn.vf_forward(vtheta, vspec, x)[:, 0]
        oret, oad = oadv.gae_flat(data["reward"], base, off, term, 0.99, 0.95)
        osad = oadv.standardize(oad)
        ostats, oinfo = natgrad.trpo_update(theta, spec, data["ob"], data["act"], osad.astype(np.float32),
                                            data["oldprob"], 0.1, 0.01)
        want = np.array([ostats[k] for k in ("surr_before", "surr_after", "kl_before", "kl_after",
                                             "ent_before", "ent_after")])
        rel = lambda u, v: float(np.linalg.norm(np.asarray(u, np.float64) - v) / max(np.linalg.norm(v), 1e-300))
        checks = {
            "returns": np.allclose(ret, oret[sl], rtol=1e-5, atol=1e-5),
            "advantages": rel(adv, osad[sl]) < 1e-5,
            "stats": np.allclose(stats, want, rtol=1e-4, atol=1e-6),
            "theta": rel(th_new, oinfo["theta_new"]) < 5e-5,
            "success": info["success"] == int(oinfo["success"]),
        }
        gathered = [None] * world
        dist.all_gather_object(gathered, th_new.tobytes())
        checks["replicas_bit_identical"] = all(g == gathered[0] for g in gathered)
        for k, ok in checks.items():
            if not ok:
                failures.append(f"{name}/{k} (rank {rank}) stats={stats} want={want}")
        if rank == 0:
            print(name, "shard sizes ok;", {k: bool(v) for k, v in checks.items()}, flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    if failures:
        print("FAIL", failures, flush=True)
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
