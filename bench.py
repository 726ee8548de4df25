#!/usr/bin/env python
"""bench.py - TRPO update timesteps/s (Fvp + CG + line search + GAE) on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one full policy update on one synthetic batch of the named workload:
value-net forward (NnVf.predict) -> GAE + standardise (compute_advantage, core.py:63-105) ->
TrpoUpdater.__call__ (trpo.py:72-140: gradient, 10 CG Fisher-vector products, shs product,
backtracking line search).  Parameters are reset to the same theta before every step so that
all steps do identical work.

`value`: inputs resident in HBM when the timed region starts.  `e2e`: the same update through
the C ABI with HOST (pinned) buffers - observations, actions, old probabilities, rewards and
parameters are copied host->device and returns/advantages/stats device->host inside the timed
region.  Multi-GPU: the 1M-timestep batch is sharded over the ranks (strong scaling); partial
sums are combined with NCCL all-reduces; CG runs replicated.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(cg_damping=0.1, max_kl=0.01, gamma=0.995, lam=0.97)   # battery-trpo.yaml:7-11
METRIC = "trpo_update_timesteps_per_sec"
UNIT = "timesteps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="humanoid")
    ap.add_argument("--timesteps", type=int, default=0, help="total timesteps (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="timesteps of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="sum over ranks with NCCL instead of the NVLink peer-memory push")
    return ap.parse_args()


def workload_desc(wl, n_total):
    hid = "-".join(str(d) for d in wl.dims[1:-1])
    head = "DiagGauss" if wl.head == 0 else "Categorical"
    return (f"{wl.name}: obs {wl.dims[0]}, act {wl.dims[-1]}, {hid} tanh MLP, {head}, {n_total} timesteps; "
            "step = VF predict + GAE + standardise + TRPO update (1 grad, 10 CG Fvp + 1, line search)")


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_update_rate(wl, n_sample, steps, warmup):
    """The oracle port of the reference's CPU path (float32 = the fork's floatX), timed on this box's
    host cores with all BLAS threads: GAE (scipy lfilter per path, as the reference) + one TRPO update."""
    from modular_rl_b200 import synth
    from oracle import advantage as oadv, natgrad, policy_math as pm
    spec = pm.NetSpec(wl.dims, pm.GAUSS if wl.head == 0 else pm.CAT)

    def fwd(th, ob):
        _, z = pm.forward(th, spec, ob, np.float32)
        return z if wl.head == 0 else pm.softmax(z)
    data = synth.policy_batch(wl, fwd, N=n_sample)
    base = np.tanh(data["ob"][:, 0]).astype(np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ret, adv = oadv.gae_flat(data["reward"], base, data["offsets"], data["terminated"], CFG["gamma"], CFG["lam"])
        adv = oadv.standardize(adv).astype(np.float32)
        natgrad.trpo_update(data["theta"], spec, data["ob"], data["act"], adv, data["oldprob"],
                            CFG["cg_damping"], CFG["max_kl"], dtype=np.float32)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n_sample * len(times) / sum(times), float(np.mean(times))


def run_reference(args, wl, n_total, rank, world):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_sample = min(args.cpu_sample, n_total)
    rate, sec = cpu_update_rate(wl, n_sample, args.steps, args.warmup)
    sample = f"{n_sample} of {n_total} timesteps per step, oracle port (numpy float32 + scipy lfilter), {cores} BLAS threads"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(wl, n_total), "parallelism": "cpu", **CFG},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- B200 arm
def chain_mma_flops_per_timestep(dims):
    """TF32 tensor-core flops the backward chain kernel ISSUES per timestep for an Fvp (3xTF32, widths padded to 8,
    weight-gradient m-tiles padded to 16): chain phase 3 GEMM passes over the (k-step, n-tile) pairs (W and V in the
    R-forward, W^T in the reverse sweep), grad phase one m16n8k8 block per 8 timesteps."""
    nt = [-(-d // 8) for d in dims]
    pairs = sum(nt[l - 1] * nt[l] for l in range(2, len(dims)))
    blocks = sum(-(-dims[l - 1] // 16) * nt[l] for l in range(2, len(dims)))
    return 3 * (3 * pairs) * 2048 / 16 + blocks * 3 * 2048 / 8


def algorithmic_flops_per_timestep(dims):
    d0d1 = dims[0] * dims[1]
    S = sum(dims[l - 1] * dims[l] for l in range(2, len(dims)))
    return {"l1_forward": 2 * d0d1, "l1_grad": 2 * d0d1, "mid_forward": 2 * S, "mid_backward_grad": 4 * S,
            "mid_backward_fvp": 8 * S}


def main():
    args = parse()
    from modular_rl_b200 import synth
    wl = synth.WORKLOADS[args.workload]
    n_total = args.timesteps or wl.N
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, wl, n_total, rank, world)
        return

    import torch
    import torch.distributed as dist
    from modular_rl_b200 import _lib as L
    from modular_rl_b200.device import Comm, DeviceBatch, DeviceNet
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()

    # ---- this rank's shard of the synthetic batch (strong scaling: n_total is fixed)
    n_local = n_total // world + (1 if rank < n_total % world else 0)
    rng = np.random.default_rng(wl.seed * 1000 + rank)
    theta = synth.init_params(wl.dims, wl.head, np.random.default_rng(wl.seed))
    vdims = (wl.dims[0] + 1,) + tuple(wl.dims[1:-1]) + (1,)
    vtheta = synth.init_params(vdims, synth.VALUE, np.random.default_rng(wl.seed + 1), last_scale=1.0)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    ob_h = pin(synth.make_obs(n_local, wl.dims[0], rng))
    offsets, terminated = synth.make_paths(n_local, wl.t_max, rng)
    reward_h = pin(rng.standard_normal(n_local))

    net = DeviceNet(wl.dims, wl.head, device=local_rank)
    vf = DeviceNet(vdims, synth.VALUE, device=local_rank)
    batch = DeviceBatch(wl.dims[0], True, device=local_rank)
    comm = None
    if world > 1:
        from modular_rl_b200.parallel import comm_from_torch_distributed
        comm = comm_from_torch_distributed(local_rank, p2p=not args.nccl_only)
        net.set_comm(comm)
    batch.set_obs(ob_h.numpy()).set_paths(offsets, terminated, float(wl.t_max))
    batch.set_global_n(n_total)
    net.set_params(theta)
    vf.set_params(vtheta)
    out = net.forward(batch)                                   # the policy's own output = path["prob"]
    if wl.head == synth.GAUSS:
        d = wl.dims[-1]
        oldprob = np.concatenate([out, np.broadcast_to(np.exp(theta[-d:])[None], out.shape)], 1).astype(np.float32)
    else:
        oldprob = out
    act = synth.sample_actions(wl.head, oldprob, rng)
    act_h, oldprob_h = pin(act), pin(oldprob)
    theta_h = pin(theta)
    # move the policy slightly off theta_old, as after a few updates, so that ratios/KL are not trivial
    theta_cur = pin(synth.perturb(theta, 0.01, wl.seed + 7))
    reward_d = reward_h.to(dev)
    theta_d = theta_cur.to(dev)
    batch.set_policy_inputs(wl.head, wl.dims[-1], act_h.numpy(), np.zeros(n_local, np.float32), oldprob_h.numpy())
    del out

    def step_resident():
        vf.predict_into_baseline(batch)
        batch.gae(reward_d, None, CFG["gamma"], CFG["lam"], True, comm, want_outputs=False)
        batch.refresh_advantages()
        net.set_params(theta_d)
        return net.trpo_step(batch, CFG["cg_damping"], CFG["max_kl"])

    batch2 = DeviceBatch(wl.dims[0], True, device=local_rank)
    ret_h = torch.empty(n_local, dtype=torch.float64).pin_memory()
    adv_h = torch.empty(n_local, dtype=torch.float64).pin_memory()

    def step_e2e():
        batch2.set_obs(ob_h.numpy()).set_paths(offsets, terminated, float(wl.t_max))
        batch2.set_global_n(n_total)
        vf.predict_into_baseline(batch2)
        batch2.gae(reward_h.numpy(), None, CFG["gamma"], CFG["lam"], True, comm, out=(ret_h.numpy(), adv_h.numpy()))
        batch2.set_policy_inputs(wl.head, wl.dims[-1], act_h.numpy(), None, oldprob_h.numpy())
        net.set_params(theta_cur.numpy())
        return net.trpo_step(batch2, CFG["cg_damping"], CFG["max_kl"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        if profile:
            lib.mrl_profile_enable(1)
        l0 = lib.mrl_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), int(lib.mrl_launch_count() - l0), res

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The timed region proper carries no per-kernel events: an event pair around each of the ~108 launches
    # of a step opens ~4 us of launch gap each (0.43 ms per step, measured; 1 % at 1M timesteps on one
    # GPU but 7 % of the 8-GPU step).  The per-kernel table and the roofline come from a second pass of the
    # same K steps, in this process, with the library's CUDA events on the launching stream.
    ms, launches, (stats, info) = timed(step_resident, args.steps, max(args.warmup, 3))
    pms_total, _, _ = timed(step_resident, args.steps, 1, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    nk = lib.mrl_profile_kinds()
    pms, pcnt = (C.c_double * nk)(), (C.c_longlong * nk)()
    L.check(lib.mrl_profile_read(pms, pcnt))
    lib.mrl_profile_enable(0)
    value = n_total * args.steps / (ms * 1e-3)
    e2e = None
    if not args.no_e2e:
        ems, _, _ = timed(step_e2e, args.steps, 3)
        h2d = (ob_h.numel() * 4 + act_h.numel() * act_h.element_size() + oldprob_h.numel() * 4 +
               reward_h.numel() * 8 + theta_cur.numel() * 4 + offsets.nbytes + terminated.nbytes) * world
        d2h = (ret_h.numel() * 8 + adv_h.numel() * 8 + 6 * 8 + 4 * 32) * world
        e2e = {"value": n_total * args.steps / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / args.steps}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp32, mma32 = C.c_double(), C.c_double()
        L.check(lib.mrl_measure_fp32_tflops(local_rank, C.byref(fp32)))
        L.check(lib.mrl_measure_mma_tf32_tflops(local_rank, C.byref(mma32)))
        flops = algorithmic_flops_per_timestep(wl.dims)
        kernels = {}
        for k in range(nk):
            name = lib.mrl_profile_kind_name(k).decode()
            if pcnt[k] == 0:
                continue
            avg_ms = pms[k] / pcnt[k]
            ent = {"launches_per_step": pcnt[k] / args.steps, "ms_per_step": pms[k] / args.steps, "avg_ms": avg_ms}
            if name in flops:
                tf = flops[name] * n_local / (avg_ms * 1e-3) / 1e12
                ent.update(algo_tflops=tf, frac_fp32_peak=tf / fp32.value)
            if name == "gae":
                ent.update(algo_gbs=24.0 * n_local / (avg_ms * 1e-3) / 1e9)
            kernels[name] = ent
        top = max((k for k in kernels if k in flops), key=lambda k: kernels[k]["ms_per_step"])
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tr["workload"] == wl.name and tr["timesteps"] == n_local and top in tr["kernels"]:
                traffic = tr["kernels"][top]["read"] + tr["kernels"][top]["write"]
        except Exception:
            pass
        pipes = {"mid_backward_fvp": "per-warp register chain, mma.sync TF32 x3 split precision (FP32-class accuracy) + FP32 epilogues",
                 "mid_backward_grad": "per-warp register chain, mma.sync TF32 x3 split precision + FP32 heads",
                 "mid_forward": "per-warp register chain, mma.sync TF32 x3 split precision + FP32 heads",
                 "l1_forward": "tcgen05.mma kind::tf32 x3 split precision, TMEM accumulators, raw-fp32 operand split in shared memory",
                 "l1_grad": "tcgen05.mma kind::tf32 x3 split precision, TMEM accumulators, raw-fp32 operand split in shared memory"}
        roof = {"kernel": top, "bound": "tensor", "achieved": kernels[top]["algo_tflops"], "peak": peak_tf,
                "unit": "TFLOP/s", "frac": kernels[top]["algo_tflops"] / peak_tf, "traffic": traffic,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else
                                "fallback 1.4 PFLOP/s sustained (of fallback)"),
                "pipe": pipes.get(top, ""), "flops_counted": "algorithmic FP32 flops; each costs 3 TF32 MMAs",
                "tf32_mma_tflops_issued": (chain_mma_flops_per_timestep(wl.dims) * n_local / (kernels[top]["avg_ms"] * 1e-3) / 1e12
                                           if top == "mid_backward_fvp" else 3.0 * kernels[top]["algo_tflops"]),
                "mma_sync_tf32_peak_tflops": mma32.value,
                "fp32_fma_peak_tflops": fp32.value,
                "frac_of_fp32_fma_peak": kernels[top]["algo_tflops"] / fp32.value,
                "share_of_step": kernels[top]["ms_per_step"] / (pms_total / args.steps),
                "kernel_events": "second pass of the same %d steps with per-kernel CUDA events (%.3f ms per step; "
                                 "the timed region itself has none)" % (args.steps, pms_total / args.steps)}
        roof["frac_of_mma_sync_tf32_peak"] = roof["tf32_mma_tflops_issued"] / mma32.value if mma32.value else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            n_sample = min(args.cpu_sample, n_total)
            rate, sec = cpu_update_rate(wl, n_sample, 1, 0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"one update on {n_sample} of {n_total} timesteps ({sec:.1f} s), oracle port "
                             f"(numpy float32 + scipy lfilter), {cores} BLAS threads"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_desc(wl, n_total), "timesteps": n_total,
                           "timesteps_per_gpu": n_local, "parallelism": f"dp{world}",
                           "allreduce": ("none" if world == 1 else ("nvlink peer-memory push (fused into the slab reduce)"
                                                                    if comm.p2p else "nccl")),
                           "l2": "inputs larger than L2 (observations %.2f GB per GPU)" % (n_local * wl.dims[0] * 4 / 1e9),
                           **CFG},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "kernels": kernels,
                "update": {"stats": [float(s) for s in stats], "info": info}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if comm:
            comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
