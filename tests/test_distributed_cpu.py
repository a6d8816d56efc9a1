"""Multi-rank plumbing on CPU: gloo backend, world_size 2 (and 3).  The per-rank compute is the
oracle here (tests may use it); on GPUs the same code path calls SoGpu.so per rank (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from so_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        # rank 0 owns the snapshot; the packed float4 particle array is replicated by ONE broadcast
        s = synth.make_snapshot(24 ** 3, 40, seed=seed, nmax=1500)
        xyzm = torch.zeros((s.n, 4), dtype=torch.float32)
        if rank == 0:
            xyzm[:, :3] = torch.from_numpy(s.pos)
            xyzm[:, 3] = float(s.mass)
        parallel.broadcast_particles(xyzm, src=0)
        pos = xyzm[:, :3].numpy().copy()
        mass = np.float32(xyzm[0, 3].item())
        o = po.Oracle(pos, mass)

        def compute(c, r):
            res = o.so(c, r, np.float32(200.0), 8, want_members=False)
            return {"rvir": res["rvir"], "mvir": res["mvir"], "ndelta": res["ndelta"]}

        merged = parallel.distributed_so(compute, s.centers, s.rgtp, s.n, 1.0)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: v for k, v in merged.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_so_equals_single_rank(tmp_path, world):
    from oracle import pyoracle as po
    seed = 500 + world
    port = _free_port()
    mp.spawn(_worker, args=(world, port, seed, str(tmp_path)), nprocs=world, join=True)
    s = synth.make_snapshot(24 ** 3, 40, seed=seed, nmax=1500)
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8, want_members=False)
    for r in range(world):
        z = np.load(str(tmp_path / ("rank%d.npz" % r)))
        assert z["rvir"].tobytes() == ref["rvir"].tobytes()
        assert z["mvir"].tobytes() == ref["mvir"].tobytes()
        assert np.array_equal(z["ndelta"], ref["ndelta"])
        assert set(np.unique(z["rank_of"])) <= set(range(world))


def test_lpt_assignment_is_balanced_and_deterministic():
    rng = np.random.default_rng(3)
    cost = np.concatenate([rng.pareto(1.2, 5000) * 100 + 20, [1e6, 8e5, 5e5]])
    for world in (1, 2, 4, 8):
        rank, load = parallel.lpt_assign(cost, world)
        rank2, _ = parallel.lpt_assign(cost, world)
        assert np.array_equal(rank, rank2)
        assert rank.min() >= 0 and rank.max() < world
        for r in range(world):
            assert np.isclose(load[r], cost[rank == r].sum())
        # LPT bound: max load <= mean + largest item
        assert load.max() <= cost.sum() / world + cost.max() + 1e-6
        if world > 1:
            assert load.max() / (cost.sum() / world) < 1.0 + world * cost.max() / cost.sum() + 1e-9
    # every halo exactly once
    rank, _ = parallel.lpt_assign(cost, 8)
    parts = [parallel.shard_indices(rank, r) for r in range(8)]
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(len(cost)))


def test_halo_cost_scales_with_radius_cubed():
    c = parallel.halo_cost(np.array([0.01, 0.02]), 1e6, 1.0)
    assert np.isclose((c[1] - 64) / (c[0] - 64), 8.0)


def test_spatial_assignment_is_balanced_and_compact():
    """Domain runs: halos are cut into spatially compact, cost-balanced shares (so_b200.parallel.spatial_assign)."""
    rng = np.random.default_rng(3)
    c = (rng.random((20000, 3)) - 0.5).astype(np.float32)
    cost = 10.0 ** rng.uniform(1, 4, 20000)
    for r in (1, 2, 3, 8):
        rank, load = parallel.spatial_assign(c, cost, r)
        assert rank.min() == 0 and rank.max() == r - 1
        np.testing.assert_allclose(load.sum(), cost.sum())
        assert load.max() / load.mean() < 1.05
        again, _ = parallel.spatial_assign(c, cost, r)
        assert np.array_equal(rank, again)                  # deterministic: every rank computes the same split
    # compactness: a share occupies far fewer coarse blocks than a random share of the same size would
    rank, _ = parallel.spatial_assign(c, cost, 8)
    blocks = (np.floor((c + 0.5) * 8).astype(int) % 8) @ np.array([1, 8, 64])
    mine = len(np.unique(blocks[rank == 3]))
    rand = len(np.unique(blocks[rng.permutation(len(c))[: (rank == 3).sum()]]))
    assert mine < 0.4 * rand


def test_exchange_plan_offsets():
    cm = np.array([[5, 0, 2], [1, 7, 0], [0, 3, 4]])
    total, off = parallel.exchange_plan(cm)
    assert total.tolist() == [6, 10, 6]
    assert off.tolist() == [[0, 0, 0], [5, 0, 2], [6, 7, 2]]
    # every destination's buffer is tiled exactly by the sources' ranges
    for d in range(3):
        spans = sorted((off[s, d], off[s, d] + cm[s, d]) for s in range(3))
        pos = 0
        for a, b in spans:
            assert a == pos
            pos = b
        assert pos == total[d]
    assert parallel.slice_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]


def _needed_by(pos, centers, reach):
    """numpy stand-in for k_mark_mask + k_route: particle i is needed by a rank if it lies inside the periodic
    cube of half-width reach[h] around one of the rank's halo centres."""
    need = np.zeros(len(pos), bool)
    for c, b in zip(centers, reach):
        d = np.abs(pos - c)
        d = np.minimum(d, 1.0 - d)
        need |= (d <= b).all(axis=1)
    return need


def _domain_worker(rank, world, port, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        s = synth.make_snapshot(24 ** 3, 40, seed=seed, nmax=1500)       # every rank can regenerate the catalog ...
        a, b = parallel.slice_bounds(s.n, world)[rank]
        my_pos = s.pos[a:b]                                               # ... but only holds ITS slice of the particles
        rank_of, _ = parallel.spatial_assign(s.centers, parallel.halo_cost(s.rgtp, s.n, 1.0), world)
        reach = np.maximum(s.rgtp.astype(np.float64) * 1.2 ** 8, 0.02)
        # which of my particles does every rank need; count matrix by all_gather; offsets by exchange_plan
        need = [_needed_by(my_pos, s.centers[rank_of == r], reach[rank_of == r]) for r in range(world)]
        counts = torch.tensor([int(n.sum()) for n in need], dtype=torch.int64)
        rows = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, counts)
        cm = torch.stack(rows).numpy()
        recv_total, recv_off = parallel.exchange_plan(cm)
        # records {x, y, z, global index} grouped by destination, then the all-to-all
        send = np.concatenate([np.concatenate([my_pos[n], (a + np.nonzero(n)[0]).astype(np.float64)[:, None]], axis=1)
                               for n in need]).astype(np.float64)
        recv = torch.zeros((int(recv_total[rank]), 4), dtype=torch.float64)
        dist.all_to_all_single(recv, torch.from_numpy(send), output_split_sizes=[int(x) for x in cm[:, rank]],
                               input_split_sizes=[int(x) for x in cm[rank]])
        recv = recv.numpy()
        # what src sent sits at the planned offset of my buffer
        for src in range(world):
            seg = recv[recv_off[src, rank]:recv_off[src, rank] + cm[src, rank], 3].astype(np.int64)
            sa, sb = parallel.slice_bounds(s.n, world)[src]
            assert ((seg >= sa) & (seg < sb)).all()
        gidx = recv[:, 3].astype(np.int64)
        mine = parallel.shard_indices(rank_of, rank)
        out = {"mine": mine}
        if len(mine):
            o = po.Oracle(recv[:, :3].astype(np.float32), s.mass)
            res = o.so(s.centers[mine], s.rgtp[mine], np.float32(200.0), 8)
            out.update(rvir=res["rvir"], mvir=res["mvir"], ndelta=res["ndelta"], member_offset=res["member_offset"],
                       members=gidx[res["members"]].astype(np.int64))      # back to global particle indices
        np.savez(os.path.join(out_dir, "dom%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_domain_run_protocol_over_gloo(tmp_path, world):
    """The domain-run protocol (slices, spatial halo shares, count matrix, planned offsets, all-to-all of
    {x,y,z,global index} records, per-rank solve) with numpy in place of the device kernels and the oracle as
    the per-rank solver: merged results = the single-rank oracle, member indices global."""
    from oracle import pyoracle as po
    seed = 600 + world
    mp.spawn(_domain_worker, args=(world, _free_port(), seed, str(tmp_path)), nprocs=world, join=True)
    s = synth.make_snapshot(24 ** 3, 40, seed=seed, nmax=1500)
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8)
    seen = np.zeros(s.h, bool)
    for r in range(world):
        z = np.load(str(tmp_path / ("dom%d.npz" % r)))
        mine = z["mine"]
        seen[mine] = True
        if not len(mine):
            continue
        assert z["rvir"].tobytes() == ref["rvir"][mine].tobytes()
        assert z["mvir"].tobytes() == ref["mvir"][mine].tobytes()
        assert np.array_equal(z["ndelta"], ref["ndelta"][mine])
        for k, i in enumerate(mine):
            a = z["members"][z["member_offset"][k]:z["member_offset"][k + 1]]
            b = ref["members"][ref["member_offset"][i]:ref["member_offset"][i + 1]]
            assert np.array_equal(np.sort(a), np.sort(b))
    assert seen.all()


def _worker_domain_step(rank, world, port, seed, out_dir):
    """The host side of the domain step on CPU ranks: ownership restated in numpy (every rank computes the same
    one, no exchange), each rank solves only the halos it owns over only the particles routed to it, the
    results are merged with the NOT_MINE convention by MAX reductions (gloo)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        s = synth.make_snapshot(28 ** 3, 60, seed=seed, nmax=1500)
        owner = parallel.owner_numpy(s.centers, s.rgtp, s.n, world)
        mine = np.nonzero(owner == rank)[0]
        # routing stand-in: the particles inside the cubes this rank's halos can reach after 4 schedule balls
        keep = np.zeros(s.n, bool)
        for i in mine:
            w = float(s.rgtp[i]) * 1.2 ** 4 * 1.05 + 2.0 / 32
            d = s.pos - s.centers[i]
            d -= np.rint(d)
            keep |= (np.abs(d) < w).all(axis=1)
        idx = np.nonzero(keep)[0]
        code = torch.full((s.h,), int(parallel.NOT_MINE), dtype=torch.int32)
        m = torch.full((s.h,), float("nan"), dtype=torch.float32)
        if len(mine):
            o = po.Oracle(s.pos[idx], s.mass)
            res = o.so(s.centers[mine], s.rgtp[mine], np.float32(200.0), 8, want_members=False)
            code[torch.from_numpy(mine)] = torch.from_numpy(np.where(res["ndelta"] > 0, res["ndelta"], res["rvir"].astype(np.int32)))
            m[torch.from_numpy(mine)] = torch.from_numpy(res["mvir"])
        code, m = parallel.merge_owned(code, m)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), code=code.numpy(), m=m.numpy(), owner=owner)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_domain_step_host_logic_on_cpu_ranks(tmp_path, world):
    from oracle import pyoracle as po
    seed = 700 + world
    port = _free_port()
    mp.spawn(_worker_domain_step, args=(world, port, seed, str(tmp_path)), nprocs=world, join=True)
    s = synth.make_snapshot(28 ** 3, 60, seed=seed, nmax=1500)
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8, want_members=False)
    want = np.where(ref["ndelta"] > 0, ref["ndelta"], ref["rvir"].astype(np.int32))
    owners = None
    for r in range(world):
        z = np.load(str(tmp_path / ("rank%d.npz" % r)))
        assert np.array_equal(z["code"], want)
        ok = want > 0
        assert z["m"][ok].tobytes() == ref["mvir"][ok].tobytes()
        owners = z["owner"] if owners is None else owners
        assert np.array_equal(z["owner"], owners)            # every rank derived the same ownership
    assert len(np.unique(owners)) == world


def test_owner_numpy_is_balanced_and_compact():
    s = synth.make_snapshot(64 ** 3, 400, seed=9, nmax=300)
    for world in (1, 2, 4, 8, 16):
        own = parallel.owner_numpy(s.centers, s.rgtp, s.n, world)
        assert own.min() == 0 and own.max() == world - 1
        cost = parallel.halo_cost(s.rgtp, s.n, 1.0)
        load = np.array([cost[own == r].sum() for r in range(world)])
        assert load.max() < 1.6 * cost.sum() / world + cost.max()
    assert parallel.flags_text(0) == "ok" and "receive" in parallel.flags_text(2)
    rc, sc = parallel.default_caps(1 << 30, 8)
    assert rc > (1 << 30) * 0.3 / 8 and sc > 0 and parallel.default_caps(1000, 1)[1] == 0


def test_destination_rule_of_the_domain_step_is_exact():
    """plain / listed cells (so_b200/csrc/domain_step.cuh: k_halo_cubes, k_mark_table, k_route_split): the destination
    set derived from "owner of the cell's bin" for plain halos and from the table for halos at an ownership boundary
    equals the brute-force union of the owners of all halos whose cube covers the cell — for every cell, on
    catalogs with many overlapping cubes, for 2 .. 16 ranks and for mask grids coarser and finer than the bins."""
    rng = np.random.default_rng(5)
    for mb, world, nh in [(6, 2, 300), (6, 8, 600), (7, 5, 1500), (7, 16, 1500), (4, 3, 40)]:
        nm = 1 << mb
        centers = (rng.random((nh, 3)) - 0.5).astype(np.float32)
        rgtp = (0.004 + 0.02 * rng.random(nh) ** 3).astype(np.float32)
        owner, bin_owner = parallel.owner_numpy(centers, rgtp, 1 << 24, world, return_bins=True)
        assert len(np.unique(owner)) == world
        half = np.maximum(1, np.ceil(rgtp * 2.1 * nm).astype(np.int64))               # ~ the 4-ball cube, in cells
        c0 = np.floor((centers.astype(np.float64) + 0.5) * nm).astype(np.int64)
        cubes = [(int(c0[h, 0] - half[h]), int(c0[h, 1] - half[h]), int(c0[h, 2] - half[h]),
                  int(min(nm, 2 * half[h] + 1)), int(min(nm, 2 * half[h] + 1)), int(min(nm, 2 * half[h] + 1))) for h in range(nh)]
        dest, crossing = parallel.destinations_numpy(cubes, owner, bin_owner, mb)
        brute = np.zeros((nm, nm, nm), np.uint32)
        for h, (x0, y0, z0, nx, ny, nz) in enumerate(cubes):
            brute[np.ix_(np.arange(z0, z0 + nz) % nm, np.arange(y0, y0 + ny) % nm, np.arange(x0, x0 + nx) % nm)] |= np.uint32(1 << int(owner[h]))
        assert np.array_equal(dest, brute), (mb, world)
        if mb >= 6 and world <= 8:
            assert crossing.mean() < 0.9 and not crossing.all()          # the table is the exception, not the rule
        if mb < 5:
            assert crossing.all()                                        # bins finer than cells: everything is listed


def test_conversion_free_cell_coordinate_equals_floor():
    """coarse_coord (routing kernels) == ((int)floorf(t) & (nc - 1)) >> ms (the grid build's cell_coord), incl. negative t,
    exact cell edges and the values just below them."""
    rng = np.random.default_rng(6)
    for lb, mb in [(10, 9), (11, 9), (9, 9), (6, 6), (7, 5)]:
        nc, ms = 1 << lb, lb - mb
        g0, invh = np.float32(-0.5), np.float32(nc)
        x = np.concatenate([rng.random(200000).astype(np.float32) - np.float32(0.5),
                            (rng.random(20000).astype(np.float32) * np.float32(3.0) - np.float32(1.5)),      # outside the box: wraps
                            (np.arange(-nc, 2 * nc, dtype=np.float64) / nc - 0.5).astype(np.float32)])
        x = np.concatenate([x, np.nextafter(x, np.float32(-np.inf)), np.nextafter(x, np.float32(np.inf))])
        t = ((x - g0).astype(np.float32) * invh).astype(np.float32)
        want = ((np.floor(t).astype(np.int64) & (nc - 1)) >> ms).astype(np.uint32)
        got = parallel.coarse_coord_numpy(x, g0, invh, ms, mb)
        assert np.array_equal(got, want), (lb, mb)
