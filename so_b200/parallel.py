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


# ---------------------------------------------------------------------------------------------
# domain runs: every rank holds a slice of the snapshot and a compact share of the halos
# ---------------------------------------------------------------------------------------------

def spatial_assign(centers, cost, n_ranks, period=1.0, cells=64):
    """Spatially compact, cost-balanced halo shares: halos are ordered along a tiled (8x8x8 blocks of
    coarse cells, blocks row-major) curve through the box and cut into n_ranks consecutive pieces of
    equal estimated cost.  Compact shares keep the focus masks of the ranks nearly disjoint, so that
    few particles are sent to more than one rank.  Returns rank[h] and the load per rank."""
    centers = np.asarray(centers, np.float64).reshape(-1, 3)
    cost = np.asarray(cost, np.float64)
    h = len(cost)
    rank = np.zeros(h, np.int32)
    load = np.zeros(n_ranks, np.float64)
    if h == 0 or n_ranks <= 1:
        load[0] = cost.sum()
        return rank, load
    c = np.floor(((centers / period) % 1.0) * cells).astype(np.int64) % cells
    hi, lo = c >> 3, c & 7
    nb = max(cells >> 3, 1)
    key = (((hi[:, 2] * nb + hi[:, 1]) * nb + hi[:, 0]) << 9) | (lo[:, 2] << 6) | (lo[:, 1] << 3) | lo[:, 0]
    order = np.lexsort((np.arange(h), key))
    csum = np.cumsum(cost[order])
    total = csum[-1]
    # halo k goes to the piece that holds the midpoint of its cost interval
    mid = csum - 0.5 * cost[order]
    r = np.minimum((mid / total * n_ranks).astype(np.int64), n_ranks - 1)
    rank[order] = r.astype(np.int32)
    np.add.at(load, rank, cost)
    return rank, load


def slice_bounds(n, n_ranks):
    """[start, end) of every rank's slice of the particle array."""
    b = [(n * r) // n_ranks for r in range(n_ranks + 1)]
    return [(b[r], b[r + 1]) for r in range(n_ranks)]


def exchange_plan(count_matrix):
    """count_matrix[src][dst] = particles src sends to dst.  Returns (recv_total[dst],
    recv_offset[src][dst] = where src's records start in dst's receive buffer)."""
    cm = np.asarray(count_matrix, np.int64)
    off = np.zeros_like(cm)
    off[1:, :] = np.cumsum(cm, axis=0)[:-1, :]
    return cm.sum(axis=0), off


class VirtualDomainRun:
    """The domain run with all `n_ranks` ranks living in THIS process on one device (tests, and the
    reference implementation of the protocol): slices, masks, routing, per-rank builds and solves,
    results merged in catalog order.  `DomainRun` below is the same protocol over torch.distributed."""

    def __init__(self, n_ranks, n_balls=4):
        self.R = int(n_ranks)
        self.n_balls = int(n_balls)

    def run(self, pos, mass, centers, rgtp, thr, n_members=8, period=(1.0, 1.0, 1.0)):
        import torch
        from so_b200 import api
        R = self.R
        n = len(pos)
        dev = torch.device("cuda")
        xyzm = torch.empty((n, 4), dtype=torch.float32, device=dev)
        xyzm[:, :3] = torch.from_numpy(np.ascontiguousarray(pos, np.float32)).to(dev)
        xyzm[:, 3] = float(mass)
        torch.cuda.synchronize()
        rank_of, _ = spatial_assign(centers, halo_cost(rgtp, n, float(np.prod(period))), R, period[0])
        gs = [api.SoGpu() for _ in range(R)]
        words = gs[0].domain_mask_words(n)
        out = {"rvir": np.zeros(len(rgtp), np.float32), "mvir": np.zeros(len(rgtp), np.float32),
               "ndelta": np.zeros(len(rgtp), np.int32), "members": [None] * len(rgtp), "sent": 0, "rounds": 0}
        todo = [shard_indices(rank_of, r) for r in range(R)]
        n_balls = self.n_balls
        while any(len(t) for t in todo):
            out["rounds"] += 1
            masks = torch.zeros((R, words), dtype=torch.int32, device=dev)
            torch.cuda.synchronize()            # (torch's stream is not the handles' stream)
            for r in range(R):
                gs[r].domain_mask(n, centers[todo[r]], rgtp[todo[r]], n_balls, masks[r].data_ptr(), period)
            torch.cuda.synchronize()
            bounds = slice_bounds(n, R)
            cm = np.zeros((R, R), np.int64)
            for s in range(R):
                a, b = bounds[s]
                cm[s] = gs[s].domain_route_count(n, xyzm[a:].data_ptr(), b - a, masks.data_ptr(), R)
            recv_total, recv_off = exchange_plan(cm)
            recv = [torch.empty((max(int(t), 1), 4), dtype=torch.float32, device=dev) for t in recv_total]
            for s in range(R):
                a, b = bounds[s]
                gs[s].domain_route_scatter(n, xyzm[a:].data_ptr(), b - a, a, masks.data_ptr(),
                                           [t.data_ptr() for t in recv], recv_off[s])
            torch.cuda.synchronize()
            out["sent"] += int(recv_total.sum())
            nxt = [np.zeros(0, np.int64)] * R
            for r in range(R):
                mine = todo[r]
                if not len(mine):
                    continue
                g = gs[r]
                if recv_total[r] == 0:                       # nothing within reach: -1 for every halo (kd2.c:772-778)
                    out["rvir"][mine] = -1.0
                    out["mvir"][mine] = -1.0
                    continue
                g.set_particles_device_indexed(recv[r].data_ptr(), int(recv_total[r]), n, mass, period)
                g.build_grid_for(centers[mine], rgtp[mine], n_balls)
                d_c = torch.from_numpy(np.ascontiguousarray(centers[mine])).to(dev)
                d_r = torch.from_numpy(np.ascontiguousarray(rgtp[mine])).to(dev)
                d_n = torch.empty(len(mine), dtype=torch.int32, device=dev)
                d_m = torch.empty(len(mine), dtype=torch.float32, device=dev)
                torch.cuda.synchronize()        # the uploads above ran on torch's stream
                g.so_device(d_c.data_ptr(), d_r.data_ptr(), len(mine), thr, n_members, d_n.data_ptr(), d_m.data_ptr())
                torch.cuda.synchronize()
                code = d_n.cpu().numpy()
                done = code != -103
                fin = g.finish_host(np.where(done, code, -1).astype(np.int32), d_m.cpu().numpy(), thr)
                off, mem = g.members(copy=True)
                for k in np.nonzero(done)[0]:
                    i = mine[k]
                    out["rvir"][i], out["mvir"][i], out["ndelta"][i] = fin["rvir"][k], fin["mvir"][k], fin["ndelta"][k]
                    out["members"][i] = mem[off[k]:off[k + 1]].copy()
                nxt[r] = mine[~done]                          # balls that left the mask: again, further out
            todo = nxt
            n_balls += 4
        for g in gs:
            g.close()
        return out


class DomainRun:
    """One rank of a domain run over torch.distributed (one process per GPU, NCCL).

    Every rank holds the slice [a, b) of the particle array (device, float4 {x,y,z,m}) and the full halo
    catalog.  step(): masks of all ranks (all_gather, a few MB) -> per-destination counts of my slice
    -> count matrix (all_gather, R ints) -> scatter of {x,y,z,global index} records
        transport "p2p":  straight into the receivers' buffers, mapped once with cudaIpc: the routing
                          kernel's stores travel over NVLink, no staging copy, no collective for the data;
        transport "nccl": into a local staging buffer, then all_to_all_single
    -> barrier -> grid over what arrived (same cell size on every rank) -> SO solve of my halos.
    Halos whose balls leave the mask (code -103) are re-run with more schedule balls in the mask."""

    def __init__(self, gpu, n_total, mass, period=(1.0, 1.0, 1.0), transport="p2p", n_balls=4, group=None):
        import torch
        import torch.distributed as dist
        self.g, self.n_total, self.mass, self.period = gpu, int(n_total), float(mass), tuple(period)
        self.transport, self.n_balls, self.group = transport, int(n_balls), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.words = gpu.domain_mask_words(n_total)
        self.masks = torch.zeros((self.world, self.words), dtype=torch.int32, device=self.dev)
        self.my_mask = torch.zeros(self.words, dtype=torch.int32, device=self.dev)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.recv_cap = 0
        self.recv_ptr = 0                 # my receive buffer (peer-shareable allocation)
        self.peer_ptrs = []               # everyone's receive buffer as seen from this process
        self.send = None
        self.stats = {}

    # -- receive buffers: allocated through the library so that other processes can map them ---------
    def _ensure_recv(self, need):
        import torch
        import torch.distributed as dist
        # `need` is the largest receive count over ALL ranks (every rank knows the whole count matrix),
        # so every rank takes the same decision here without talking to the others
        if need <= self.recv_cap:
            return
        self.close_buffers()
        cap = int(need * 1.1) + 1024
        self.recv_ptr, handle = self.g.peer_alloc(cap * 16)
        self.recv_cap = cap
        if self.transport == "p2p" and self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=self.group)
            self.peer_ptrs = [self.recv_ptr if r == self.rank else self.g.peer_open(handles[r]) for r in range(self.world)]
        else:
            self.peer_ptrs = [self.recv_ptr]

    def close_buffers(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank and self.transport == "p2p" and self.world > 1:
                self.g.peer_close(p)
        self.peer_ptrs = []
        if self.recv_ptr:
            self.g.peer_free(self.recv_ptr)
        self.recv_ptr, self.recv_cap = 0, 0

    def exchange(self, d_slice, n_slice, index_base, centers_mine, rgtp_mine, n_balls):
        """Masks -> counts -> records in place.  Returns the number of records this rank received."""
        import torch
        import torch.distributed as dist
        g, R, me = self.g, self.world, self.rank
        g.domain_mask(self.n_total, centers_mine, rgtp_mine, n_balls, self.my_mask.data_ptr(), self.period)
        if R > 1:
            dist.all_gather_into_tensor(self.masks.view(-1), self.my_mask, group=self.group)
        else:
            self.masks[0].copy_(self.my_mask)
        counts = g.domain_route_count(self.n_total, d_slice, n_slice, self.masks.data_ptr(), R)
        cm = torch.zeros((R, R), dtype=torch.int64, device=self.dev)
        mine = torch.from_numpy(counts).to(self.dev)
        if R > 1:
            dist.all_gather_into_tensor(cm.view(-1), mine, group=self.group)
        else:
            cm[0] = mine
        cm = cm.cpu().numpy()
        recv_total, recv_off = exchange_plan(cm)
        self._ensure_recv(int(recv_total.max()))
        if self.transport == "p2p" or R == 1:
            g.domain_route_scatter(self.n_total, d_slice, n_slice, index_base, self.masks.data_ptr(),
                                   self.peer_ptrs if R > 1 else [self.recv_ptr], recv_off[me])
            if R > 1:
                # stream-ordered barrier: my grid build (same stream) starts after every rank's routing
                # kernel has finished, i.e. after all stores into my buffer have been performed
                dist.all_reduce(self._token, group=self.group)
        else:
            tot = int(cm[me].sum())
            if self.send is None or self.send.shape[0] < tot:
                self.send = torch.empty((int(tot * 1.1) + 1024, 4), dtype=torch.float32, device=self.dev)
            seg = np.concatenate([[0], np.cumsum(cm[me])[:-1]])
            g.domain_route_scatter(self.n_total, d_slice, n_slice, index_base, self.masks.data_ptr(),
                                   [self.send.data_ptr()] * R, seg)
            torch.cuda.current_stream().synchronize()
            recv = _as_tensor(self.recv_ptr, int(recv_total[me]), self.dev)
            dist.all_to_all_single(recv, self.send[:tot], output_split_sizes=[int(x) for x in cm[:, me]],
                                   input_split_sizes=[int(x) for x in cm[me]], group=self.group)
            torch.cuda.current_stream().synchronize()
        self.stats = {"sent": int(cm[me].sum()), "received": int(recv_total[me]), "count_matrix": cm}
        return int(recv_total[me])

    def solve(self, n_recv, d_centers, d_rgtp, nh, thr, n_members, n_balls, d_out_n, d_out_m):
        """Grid over the received records + SO solve of my halos (asynchronous)."""
        g = self.g
        if nh == 0:
            return
        if n_recv == 0:
            import torch
            _as_tensor_i32(d_out_n, nh, self.dev).fill_(-1)
            return
        g.set_particles_device_indexed(self.recv_ptr, n_recv, self.n_total, self.mass, self.period)
        g.build_grid_for_device(d_centers, d_rgtp, nh, n_balls)
        g.so_device(d_centers, d_rgtp, nh, thr, n_members, d_out_n, d_out_m)


def _as_tensor(ptr, n_records, dev):
    """A torch view of `n_records` 16-byte records at a raw device pointer (no ownership)."""
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (max(n_records, 1), 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(a, device=dev)[:n_records]


def _as_tensor_i32(ptr, n, dev):
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(a, device=dev)[:n]


# ---------------------------------------------------------------------------------------------
# the domain STEP: stream-ordered from the slice to the results (so_b200/csrc/domain_step.cuh)
# ---------------------------------------------------------------------------------------------

NOT_MINE = np.int32(-2139062144)      # 0x80808080: halo solved by another rank

FLAG_TEXT = {1: "staging area / hit list too small", 2: "receive buffer too small", 4: "a peer's receive buffer too small",
             8: "a peer never reached the barrier"}


def flags_text(flags):
    return ", ".join(t for b, t in FLAG_TEXT.items() if flags & b) or "ok"


def default_caps(n_total, n_ranks, frac=0.30):
    """(recv_cap, stage_cap) in records: the share of the snapshot some halo can reach is well below `frac`
    for the BASELINE catalogs (0.18 at 1024^3 / 10^5 halos); overflow is detected and reported, never silent."""
    recv = int(n_total * frac / n_ranks * 1.5) + (1 << 16)
    stage = int(n_total * frac / (n_ranks * n_ranks) * 2.0) + (1 << 16) if n_ranks > 1 else 0
    return min(recv, int(n_total) + 1024), stage


class DomainStep:
    """One rank of a domain run (one process per GPU).  Per step, identical calls on every rank:

        begin(catalog) -> route(slice) -> push -> solve        all enqueued, no host round trip
        result()                                                synchronises: flags, counts, owners

    The ranks only meet at the flag barrier inside push (peer memory); torch.distributed is used once,
    at construction, to exchange the cudaIpc handles of the receive buffers."""

    def __init__(self, gpu, n_total, mass, recv_cap=None, stage_cap=None, period=(1.0, 1.0, 1.0),
                 center=(0.0, 0.0, 0.0), group=None):
        import torch.distributed as dist
        self.g = gpu
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        rc, sc = default_caps(n_total, self.world)
        self.recv_cap, self.stage_cap = int(recv_cap or rc), int(stage_cap if stage_cap is not None else sc)
        handles = gpu.domain_open(self.rank, self.world, n_total, mass, self.recv_cap, self.stage_cap, period, center)
        self._opened = []
        if self.world > 1:
            allh = [None] * self.world
            dist.all_gather_object(allh, handles, group=group)
            ptrs = [[0] * self.world for _ in range(3)]
            for r in range(self.world):
                if r == self.rank:
                    continue
                for k in range(3):
                    if k == 1 and allh[r][1] == allh[r][0]:
                        ptrs[1][r] = ptrs[0][r]
                        continue
                    ptrs[k][r] = gpu.peer_open(allh[r][k])
                    self._opened.append(ptrs[k][r])
            gpu.domain_connect(ptrs[0], ptrs[1], ptrs[2])
            dist.barrier(group=group)          # every control block is zeroed and mapped before the first step
        self.group = group

    def step(self, d_centers, d_rgtp, nh, n_balls, d_slice, n_slice, index_base, thr, n_members=8,
             d_out_n=0, d_out_m=0, host_xyz=0):
        g = self.g
        g.domain_begin(d_centers, d_rgtp, nh, n_balls)
        if host_xyz:
            g.domain_route_host(host_xyz, n_slice, index_base, d_slice)
        else:
            g.domain_route(d_slice, n_slice, index_base)
        g.domain_push(True)
        g.domain_solve(thr, n_members, d_out_n, d_out_m)

    def result(self, nh=0):
        return self.g.domain_result(nh)

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        for p in self._opened:
            self.g.peer_close(p)
        self._opened = []
        self.g.domain_close()


class VirtualDomainStep:
    """All ranks of a domain run inside ONE process (tests; also a one-process multi-GPU run when `devices`
    lists several ordinals): one handle per rank, plain device pointers instead of cudaIpc handles.  On a single
    device the peer barrier is replaced by a device synchronisation between the push and the solve phases (the
    ranks' kernels are serialised there anyway)."""

    def __init__(self, n_ranks, n_total, mass, devices=None, recv_cap=None, stage_cap=None, period=(1.0, 1.0, 1.0),
                 frac=0.30):
        from so_b200 import api
        self.R = int(n_ranks)
        self.devices = list(devices) if devices else [0] * self.R
        self.multi = len(set(self.devices)) > 1
        rc, sc = default_caps(n_total, self.R, frac)
        self.recv_cap, self.stage_cap = int(recv_cap or rc), int(stage_cap if stage_cap is not None else sc)
        self.gs = [api.SoGpu(device=self.devices[r]) for r in range(self.R)]
        self.n_total, self.mass, self.period = int(n_total), mass, tuple(period)
        for r, g in enumerate(self.gs):
            g.domain_open(r, self.R, n_total, mass, self.recv_cap, self.stage_cap, period)
        ptr = [g.domain_pointers() for g in self.gs]
        for r, g in enumerate(self.gs):
            for q in range(self.R):
                if self.devices[q] != self.devices[r]:
                    g.enable_peer_access(self.devices[q])
            if self.R > 1:
                g.domain_connect([p[0] for p in ptr], [p[1] for p in ptr], [p[2] for p in ptr])

    def _sync(self):
        import torch
        for d in set(self.devices):
            torch.cuda.synchronize(d)

    def run(self, xyzm, centers, rgtp, thr, n_members=8, n_balls=4, want_members=True):
        """xyzm: list of per-rank device tensors (n_r, 4) float32, rank r's slice starting at global index
        sum(n_0..n_{r-1}); centers/rgtp: the whole catalog (numpy).  Returns merged catalog-order results."""
        import torch
        R, nh = self.R, len(rgtp)
        base = np.concatenate([[0], np.cumsum([len(x) for x in xyzm])]).astype(np.int64)
        cat = []
        for r in range(R):
            dev = torch.device("cuda", self.devices[r])
            cat.append((torch.from_numpy(np.ascontiguousarray(centers, np.float32)).to(dev),
                        torch.from_numpy(np.ascontiguousarray(rgtp, np.float32)).to(dev),
                        torch.empty(nh, dtype=torch.int32, device=dev), torch.empty(nh, dtype=torch.float32, device=dev)))
        out = {"rvir": np.zeros(nh, np.float32), "mvir": np.zeros(nh, np.float32), "ndelta": np.zeros(nh, np.int32),
               "members": [None] * nh, "rounds": 0, "n_recv": [], "owner": None}
        balls = int(n_balls)
        done = np.zeros(nh, bool)
        while not done.all():
            out["rounds"] += 1
            self._sync()
            for r, g in enumerate(self.gs):
                g.domain_begin(cat[r][0].data_ptr(), cat[r][1].data_ptr(), nh, balls)
                g.domain_route(xyzm[r].data_ptr(), len(xyzm[r]), int(base[r]))
            if not self.multi:
                self._sync()                      # one device: reservations of every rank before any push reads them
            for g in self.gs:
                g.domain_push(barrier=self.multi)
            if not self.multi:
                self._sync()
            for r, g in enumerate(self.gs):
                g.keep_member_d2(True)
                g.domain_solve(thr, n_members, cat[r][2].data_ptr(), cat[r][3].data_ptr())
            self._sync()
            n_recv = []
            for r, g in enumerate(self.gs):
                res = g.domain_result(nh)
                if res["flags"]:
                    raise RuntimeError("domain step, rank %d: %s" % (r, flags_text(res["flags"])))
                n_recv.append(res["n_recv"])
                owner = res["owner"]
                if out["owner"] is None:
                    out["owner"] = owner.copy()
                assert np.array_equal(owner, out["owner"]), "ranks disagree on the halo ownership"
                code, m = cat[r][2].cpu().numpy(), cat[r][3].cpu().numpy()
                mine = np.nonzero((owner == r) & ~done)[0]
                assert np.all(code[owner != r] == NOT_MINE)
                ok = mine[code[mine] != -103]
                if len(ok):
                    fin = g.finish_host(code[ok], m[ok], thr)
                    out["rvir"][ok], out["mvir"][ok], out["ndelta"][ok] = fin["rvir"], fin["mvir"], fin["ndelta"]
                    if want_members:
                        off, mem = g.members(sorted=True, copy=True)
                        for i in ok:
                            out["members"][i] = mem[off[i]:off[i + 1]].copy()
                    done[ok] = True
            out["n_recv"].append(n_recv)
            balls += 4                            # balls that left the mask: again, further out (whole step)
            if out["rounds"] > 16:
                raise RuntimeError("domain step does not converge")
        return out

    def close(self):
        self._sync()
        for g in self.gs:
            g.domain_close()
            g.close()


def owner_numpy(centers, rgtp, n_total, n_ranks, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0), return_bins=False):
    """The halo ownership of a domain step, restated in numpy operation for operation (k_assign_hist / _scan /
    _owner in so_b200/csrc/domain_step.cuh: 32^3 bins along a tiled curve, integer costs, cut at equal cost).
    Every rank of a domain step computes exactly this on its device; host-side planning and the CPU tests use it."""
    centers = np.asarray(centers, np.float32).reshape(-1, 3)
    rgtp = np.asarray(rgtp, np.float32)
    h = len(rgtp)
    if n_ranks <= 1:
        return (np.zeros(h, np.uint8), np.zeros(32768, np.uint8)) if return_bins else np.zeros(h, np.uint8)
    L = np.asarray(period, np.float32).astype(np.float64)
    g0 = (np.asarray(center, np.float32).astype(np.float64) - 0.5 * L).astype(np.float32).astype(np.float64)
    t = (centers.astype(np.float64) - g0) / L
    t -= np.floor(t)
    c = np.clip((t * 32.0).astype(np.int64), 0, 31)
    hi, lo = c >> 3, c & 7
    key = ((((hi[:, 2] * 4 + hi[:, 1]) * 4 + hi[:, 0]) << 9) | (lo[:, 2] << 6) | (lo[:, 1] << 3) | lo[:, 0]).astype(np.int64)
    nbar = float(n_total) / float(np.prod(L))
    r = 1.2 * 1.25 * rgtp.astype(np.float64)
    cost = (200.0 * nbar * 4.18879020478639 * r * r * r * 0.6 + 64.0).astype(np.uint64)
    bins = np.zeros(32768 + 1, np.uint64)
    np.add.at(bins, key, cost)
    excl = np.concatenate([[0], np.cumsum(bins[:-1], dtype=np.uint64)]).astype(np.uint64)   # excl[b], excl[32768] = total
    total = int(excl[32768])
    lo_, hi_ = excl[key].astype(object), excl[key + 1].astype(object)
    mid = lo_ + (hi_ - lo_) // 2
    own = np.array([min(n_ranks - 1, (int(m) * n_ranks) // total) if total else 0 for m in mid], np.uint8)
    if return_bins:                       # k_assign_bin_owner: the same rule for every bin, occupied or not
        lo_, hi_ = excl[:-1].astype(object), excl[1:].astype(object)
        mid = lo_ + (hi_ - lo_) // 2
        bin_own = np.array([min(n_ranks - 1, (int(m) * n_ranks) // total) if total else 0 for m in mid], np.uint8)
        return own, bin_own
    return own


def bin_index_numpy(kx, ky, kz):
    """Index of the ownership bin with coordinates (kx, ky, kz) in 0..31 along the tiled curve (assign_bin_index)."""
    kx, ky, kz = (np.asarray(v, np.int64) for v in (kx, ky, kz))
    return ((((kz >> 3) * 4 + (ky >> 3)) * 4 + (kx >> 3)) << 9) | ((kz & 7) << 6) | ((ky & 7) << 3) | (kx & 7)


def destinations_numpy(cubes, owner, bin_owner, mb):
    """Destination ranks of every coarse cell (2^mb per axis) as the domain step derives them (k_halo_cubes,
    k_mark_table, k_route_split in so_b200/csrc/domain_step.cuh), restated in numpy:

      * a halo is "crossing" if its cube (x0, y0, z0, nx, ny, nz in coarse cells, periodic) touches a 32^3 ownership
        bin of another owner, "plain" otherwise;
      * plain halos set a bit per cell in `plain`, crossing halos set `listed` and OR their owner into `table`;
      * destinations(cell) = (plain ? {owner of the cell's bin} : {}) | (listed ? table : {}).

    Returns (dest, crossing): dest = uint32 bit mask of ranks per cell, shape (2^mb,)*3 indexed [z, y, x]."""
    nm = 1 << mb
    sh = mb - 5
    plain = np.zeros((nm, nm, nm), bool)
    listed = np.zeros((nm, nm, nm), bool)
    table = np.zeros((nm, nm, nm), np.uint32)
    crossing = np.zeros(len(cubes), bool)
    for h, (x0, y0, z0, nx, ny, nz) in enumerate(cubes):
        xs, ys, zs = (np.arange(a, a + n) for a, n in ((x0, nx), (y0, ny), (z0, nz)))
        if sh < 0:
            crossing[h] = True
        else:
            bx, by, bz = (np.unique((v >> sh) & 31) for v in (xs, ys, zs))
            b = bin_index_numpy(*np.meshgrid(bx, by, bz, indexing="ij"))
            crossing[h] = bool(np.any(bin_owner[b] != owner[h]))
        ix = np.ix_(zs % nm, ys % nm, xs % nm)
        if crossing[h]:
            listed[ix] = True
            table[ix] |= np.uint32(1 << int(owner[h]))
        else:
            plain[ix] = True
    dest = np.where(listed, table, np.uint32(0))
    if sh >= 0:
        c = np.arange(nm) >> sh
        kz, ky, kx = np.meshgrid(c, c, c, indexing="ij")
        of_bin = (np.uint32(1) << bin_owner[bin_index_numpy(kx, ky, kz)].astype(np.uint32)).astype(np.uint32)
        dest = dest | np.where(plain, of_bin, np.uint32(0))
    return dest, crossing


def coarse_coord_numpy(x, g0, invh, ms, mb):
    """The routing kernels' cell coordinate (coarse_coord in domain_step.cuh): t = fl(fl(x - g0) * invh) in fp32, then
    ONE fp32 add of 1.5 * 2^(23 + ms) rounded towards -infinity, whose low mantissa bits are floor(t / 2^ms); & (2^mb - 1).
    Emulated exactly: the sum of two fp32 numbers is exact in fp64, and rounding it down to fp32 is a comparison."""
    x = np.asarray(x, np.float32)
    t = ((x - np.float32(g0)).astype(np.float32) * np.float32(invh)).astype(np.float32)
    magic = np.float32(1.5 * 2.0 ** (23 + ms))
    exact = t.astype(np.float64) + np.float64(magic)
    f = exact.astype(np.float32)
    f = np.where(f.astype(np.float64) > exact, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)
    return f.view(np.uint32) & np.uint32((1 << mb) - 1)


def merge_owned(code, m, group=None):
    """Per-halo results of a domain step (torch tensors over the WHOLE catalog: N_Delta / code and M_Delta for the
    halos this rank owns, the NOT_MINE pattern elsewhere) -> the complete arrays on every rank.  Every halo is owned
    by exactly one rank, so a MAX reduction over (code, mass-or-minus-infinity) merges them; works on any backend."""
    import torch
    import torch.distributed as dist
    code = code.clone()
    mm = torch.where(code == int(NOT_MINE), torch.full_like(m, -float("inf")), m)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(code, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
    return code, mm
