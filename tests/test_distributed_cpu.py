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
