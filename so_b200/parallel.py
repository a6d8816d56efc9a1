"""Multi-GPU plumbing of the SO hot path: one process per GPU.

Halos are independent units (kdRvir reads no cross-halo state, kd2.c:723-831), so they are
sharded across ranks by longest-processing-time-first on an estimated particle count, while the
particle array is replicated on every GPU by ONE broadcast over NCCL/NVLink (the only exchange
step of the path).  No collective sits inside any kernel loop.  Results are gathered to rank 0
and put back in catalog order, so they do not depend on the number of ranks.

torch.distributed is used for the plumbing only (gloo on CPU in the tests, nccl on GPUs).
"""
from __future__ import annotations

import numpy as np


def halo_cost(rgtp, n_particles, volume, overdensity=200.0):
    """Estimated r^2 evaluations of one halo: particles inside the final ball (1.2 R, R ~ 1.25
    rgtp) at mean enclosed density `overdensity` x the mean number density, plus a floor for
    the fixed per-halo work."""
    r = 1.2 * 1.25 * np.asarray(rgtp, np.float64)
    nbar = n_particles / float(volume)
    return overdensity * nbar * (4.0 * np.pi / 3.0) * r ** 3 * 0.6 + 64.0


def lpt_assign(cost, n_ranks):
    """Longest-processing-time-first: returns rank[h] for every halo and the load per rank.
    Deterministic (ties by index).  Big halos first, each to the least-loaded rank; the long tail
    of small halos is dealt in blocks to keep this O(H log H) without a heap per item."""
    cost = np.asarray(cost, np.float64)
    h = len(cost)
    rank = np.zeros(h, np.int32)
    load = np.zeros(n_ranks, np.float64)
    if n_ranks <= 1 or h == 0:
        load[0] = cost.sum()
        return rank, load
    order = np.lexsort((np.arange(h), -cost))
    import heapq
    heap = [(0.0, r) for r in range(n_ranks)]
    for i in order:
        l, r = heapq.heappop(heap)
        rank[i] = r
        l += cost[i]
        load[r] = l
        heapq.heappush(heap, (l, r))
    return rank, load


def shard_indices(rank_of, r):
    """Catalog indices owned by rank r, ascending."""
    return np.nonzero(np.asarray(rank_of) == r)[0].astype(np.int64)


def broadcast_particles(xyzm, src=0, group=None):
    """Replicate the packed float4 particle tensor from `src` to every rank (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(xyzm, src=src, group=group)
    return xyzm


def gather_results(local_idx, local_arrays, h_total, group=None, device=None):
    """All ranks contribute per-halo arrays for their shard; every rank gets the full arrays in
    catalog order.  local_arrays: dict name -> 1-D numpy array aligned with local_idx."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = {}
    if world == 1:
        for k, a in local_arrays.items():
            full = np.zeros(h_total, a.dtype)
            full[local_idx] = a
            out[k] = full
        return out
    dev = device if device is not None else "cpu"
    # pad every shard to the same length so a plain all_gather works on every backend
    n_local = torch.tensor([len(local_idx)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    idx_pad = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    idx_pad[:len(local_idx)] = torch.as_tensor(local_idx, dtype=torch.int64, device=dev)
    idx_all = [torch.empty_like(idx_pad) for _ in range(world)]
    dist.all_gather(idx_all, idx_pad, group=group)
    for k, a in local_arrays.items():
        t = torch.zeros((nmax,), dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, device=dev)
        t[:len(a)] = torch.as_tensor(a, device=dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        full = np.zeros(h_total, a.dtype)
        for r in range(world):
            ii = idx_all[r][:counts[r]].cpu().numpy()
            full[ii] = parts[r][:counts[r]].cpu().numpy()
        out[k] = full
    return out


def distributed_so(compute, centers, rgtp, n_particles, volume, group=None, device=None):
    """Shard `centers/rgtp` over the ranks of the default process group, run
    compute(centers_shard, rgtp_shard) -> dict of per-halo arrays on each rank, and return the
    merged catalog-order arrays on every rank.  `compute` is the single-GPU call (SoGpu.so)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    me = dist.get_rank(group) if dist.is_initialized() else 0
    rank_of, load = lpt_assign(halo_cost(rgtp, n_particles, volume), world)
    mine = shard_indices(rank_of, me)
    local = compute(np.ascontiguousarray(centers[mine]), np.ascontiguousarray(rgtp[mine])) if len(mine) else {}
    if not len(mine):
        local = {"rvir": np.zeros(0, np.float32), "mvir": np.zeros(0, np.float32), "ndelta": np.zeros(0, np.int32)}
    merged = gather_results(mine, {k: np.asarray(v) for k, v in local.items() if np.ndim(v) == 1},
                            len(rgtp), group=group, device=device)
    merged["rank_of"] = rank_of
    merged["load"] = load
    return merged
