"""Data-parallel driver logic: one process per GPU, the timestep batch sharded by WHOLE
trajectories (so GAE and the within-path time index need no halo), partial sums combined by an
all-reduce of the parameter-sized vector, CG / line search / L-BFGS replicated on identical bits.

The reference has no parallel path (core.py:123-124 raises NotImplementedError); this module is the
host-side half of DESIGN.md "Multi-GPU".  Everything here is plain Python/numpy so that it can be
tested on CPU with the gloo backend; the device-side sums run in libmrl_b200.so over NCCL.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def bind_to_device_numa(device_index: int) -> int:
    """Restrict this process to the CPU cores closest to GPU `device_index` (NVML's ideal affinity mask, intersected
    with the cores the process may use) so that the pinned host buffers it allocates afterwards are first-touched on
    the GPU's own NUMA node.  With 8 ranks on a two-socket host the uploads otherwise share one socket's memory and
    inter-socket links (measured 22 GB/s per GPU at N = 8 against 54 GB/s at N = 2).  Returns the number of cores
    bound to, 0 when NVML or the affinity call is unavailable (nothing changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cores = ideal & set(os.sched_getaffinity(0))
        if not cores:
            return 0
        os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


def shard_bounds(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of whole trajectories, balanced by timestep count.

    Path p goes to the rank whose ideal span [r*N/world, (r+1)*N/world) contains the path's FIRST
    timestep.  Returns [(first_path, end_path)] per rank; blocks are disjoint, ordered and cover every
    path; a rank may be empty when there are fewer paths than ranks."""
    lengths = np.asarray(lengths, np.int64)
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) if len(lengths) else np.zeros(0, np.int64)
    N = int(lengths.sum())
    out = []
    for r in range(world):
        lo_t = (N * r) // world
        hi_t = (N * (r + 1)) // world
        a = int(np.searchsorted(starts, lo_t, side="left"))
        b = int(np.searchsorted(starts, hi_t, side="left")) if r < world - 1 else len(lengths)
        out.append((a, b))
    return out


def shard_paths(paths: list, rank: int, world: int) -> list:
    a, b = shard_bounds([len(p["reward"]) for p in paths], world)[rank]
    return paths[a:b]


def merge_moments(triples: Sequence[Tuple[float, float, float]]) -> Tuple[float, float, float]:
    """Chan merge of (n, mean, M2) in the given (rank) order - the same arithmetic, in the same order,
    as merge_moments_kernel, so every rank derives bit-identical mean/std for the advantage
    standardisation (core.py:100-105)."""
    n, mean, m2 = 0.0, 0.0, 0.0
    for bn, bm, b2 in triples:
        if bn == 0:
            continue
        if n == 0:
            n, mean, m2 = float(bn), float(bm), float(b2)
            continue
        tot = n + bn
        d = bm - mean
        mean = mean + d * (bn / tot)
        m2 = m2 + b2 + d * d * (n * bn / tot)
        n = tot
    return n, mean, m2


def zfilter_prefix(states: Sequence[Tuple[float, np.ndarray, np.ndarray]], rank: int):
    """Exclusive prefix of Welford states over ranks: the state a rank's ZFilter scan must start from
    when consecutive sample blocks live on consecutive ranks (running_stat.py semantics)."""
    n, M, S = 0.0, None, None
    for bn, bM, bS in states[:rank]:
        bM, bS = np.asarray(bM, np.float64), np.asarray(bS, np.float64)
        if bn == 0:
            continue
        if n == 0:
            n, M, S = float(bn), bM.copy(), bS.copy()
            continue
        tot = n + bn
        d = bM - M
        M = M + d * (bn / tot)
        S = S + bS + d * d * (n * bn / tot)
        n = tot
    return n, M, S


P2P_MAX_DOUBLES = 1 << 17     # receive-slot size: covers parameter vectors up to 131 072 entries


def comm_from_torch_distributed(device: int, p2p: bool = True):
    """Build the library's communicator inside an initialised torch.distributed job: rank 0 creates the
    NCCL unique id, torch broadcasts it (plumbing only).  With p2p (default) the ranks then exchange the
    CUDA IPC handles of their NVLink receive buffers, so that parameter-sized sums bypass NCCL; if any rank
    cannot map a peer (no peer access between the GPUs) every rank stays on NCCL."""
    import torch.distributed as dist
    from .device import Comm
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = Comm(box[0], rank, world, device)
    if p2p and 1 < world <= 8:
        enable_p2p(comm)
    return comm


def enable_p2p(comm, max_doubles: int = P2P_MAX_DOUBLES) -> bool:
    """All ranks must call this together.  Returns True when the peer-memory transport is active."""
    import torch.distributed as dist
    try:
        handle, err = comm.p2p_export(max_doubles), None
    except RuntimeError as e:                       # every rank must still take part in the all-gather
        handle, err = b"", str(e)
    handles = [None] * comm.world
    dist.all_gather_object(handles, (handle, err))
    if any(e is not None for _, e in handles):
        return False
    ok = True
    try:
        comm.p2p_connect([h for h, _ in handles])
    except RuntimeError:
        ok = False
    flags = [None] * comm.world
    dist.all_gather_object(flags, ok)
    if not all(flags):
        return False
    comm.p2p_enable(True)
    return True
