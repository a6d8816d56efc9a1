"""Strong-scaling run of one BASELINE config as a DOMAIN RUN (so_b200.parallel.DomainRun): every rank
holds 1/R of the particles and a compact share of the halos; per step: masks -> routing of the particles
to the ranks that need them (NVLink peer stores, or NCCL all-to-all) -> per-rank grid -> SO solve.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 \
        tools/domain_bench.py --config 1 --steps 10 [--transport p2p|nccl] [--scale S]

Prints one JSON line on rank 0: ms per step (max over ranks, CUDA events), its breakdown, the particles
exchanged, and a bit-exact check of rank 0's halos against the single full grid."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--balls", type=int, default=4)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from so_b200 import api, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_gen = time.time()
    if rank == 0:
        if args.config == 3 and args.scale == 1.0:
            s = synth.make_snapshot(1024 ** 3, 100000, seed=1003, shuffle=False, name="cfg3_1024^3")
        elif args.config == 2 and args.scale == 1.0:
            s = synth.make_snapshot(512 ** 3, 50000, seed=1002, shuffle=False, name="cfg2_512^3")
        else:
            s = synth.config(args.config, args.scale)
        meta = [s.n, s.h, float(s.mass), s.name, float(s.omega0)]
    else:
        s, meta = None, [0, 0, 0.0, "", 1.0]
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    n, h, mass, name, omega0 = int(meta[0]), int(meta[1]), np.float32(meta[2]), meta[3], float(meta[4])
    thr = np.float32(np.float32(200.0) * np.float32(omega0))
    cat = torch.empty((h, 4), dtype=torch.float32, device=dev)
    if rank == 0:
        cat.copy_(torch.from_numpy(np.concatenate([s.centers, s.rgtp[:, None]], axis=1).astype(np.float32)))
    if world > 1:
        dist.broadcast(cat, src=0)
    cat_h = cat.cpu().numpy()
    centers, rgtp = np.ascontiguousarray(cat_h[:, :3]), np.ascontiguousarray(cat_h[:, 3])

    # every rank receives only ITS slice of the particles (rank 0 sends the others theirs)
    bounds = parallel.slice_bounds(n, world)
    a, b = bounds[rank]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    g = api.SoGpu(device=local, stream=stream.cuda_stream)
    d_slice = torch.empty((b - a, 4), dtype=torch.float32, device=dev)
    full = None
    if rank == 0:
        full = torch.empty((n, 4), dtype=torch.float32, device=dev)
        chunk = 1 << 24
        for i0 in range(0, n, chunk):
            i1 = min(n, i0 + chunk)
            full[i0:i1, :3] = torch.from_numpy(s.pos[i0:i1]).to(dev)
        full[:, 3] = float(mass)
        d_slice.copy_(full[a:b])
        for r in range(1, world):
            ra, rb = bounds[r]
            dist.send(full[ra:rb].contiguous(), dst=r)
    elif world > 1:
        dist.recv(d_slice, src=0)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    rank_of, load = parallel.spatial_assign(centers, parallel.halo_cost(rgtp, n, 1.0), world)
    mine = parallel.shard_indices(rank_of, rank)
    nh = len(mine)
    run = parallel.DomainRun(g, n, mass, transport=args.transport, n_balls=args.balls)
    d_c = torch.from_numpy(np.ascontiguousarray(centers[mine])).to(dev)
    d_r = torch.from_numpy(np.ascontiguousarray(rgtp[mine])).to(dev)
    d_n = torch.empty(max(nh, 1), dtype=torch.int32, device=dev)
    d_m = torch.empty(max(nh, 1), dtype=torch.float32, device=dev)
    t_ex = t_so = 0.0
    n_recv_first = [0]
    outgrown_first = [0]

    def step(record=False):
        nonlocal t_ex, t_so
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        n_recv = run.exchange(d_slice.data_ptr(), b - a, a, centers[mine], rgtp[mine], args.balls)
        e[1].record(stream)
        run.solve(n_recv, d_c.data_ptr(), d_r.data_ptr(), nh, thr, 8, args.balls, d_n.data_ptr(), d_m.data_ptr())
        e[2].record(stream)
        # halos whose balls left the mask: again with more of the schedule inside the mask (all ranks take part)
        code = d_n[:nh].cpu().numpy() if nh else np.zeros(0, np.int32)
        res_n, res_m = code.copy(), (d_m[:nh].cpu().numpy() if nh else np.zeros(0, np.float32))
        left = np.nonzero(code == -103)[0]
        n_left = torch.tensor([len(left)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n_left, op=dist.ReduceOp.MAX)
        balls = args.balls
        if record:
            n_recv_first[0], outgrown_first[0] = n_recv, len(left)
        while int(n_left.item()) > 0:
            balls += 4
            c2, r2 = np.ascontiguousarray(centers[mine][left]), np.ascontiguousarray(rgtp[mine][left])
            nr = run.exchange(d_slice.data_ptr(), b - a, a, c2, r2, balls)
            if len(left):
                dc2, dr2 = torch.from_numpy(c2).to(dev), torch.from_numpy(r2).to(dev)
                dn2 = torch.empty(len(left), dtype=torch.int32, device=dev)
                dm2 = torch.empty(len(left), dtype=torch.float32, device=dev)
                run.solve(nr, dc2.data_ptr(), dr2.data_ptr(), len(left), thr, 8, balls, dn2.data_ptr(), dm2.data_ptr())
                c = dn2.cpu().numpy()
                res_n[left], res_m[left] = c, dm2.cpu().numpy()
                left = left[c == -103]
            n_left = torch.tensor([len(left)], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(n_left, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        if record:
            t_ex += e[0].elapsed_time(e[1])
            t_so += e[1].elapsed_time(e[2])
        return res_n, res_m

    for _ in range(args.warmup):
        step()
    g.profile_enable(True)
    g.profile_read(reset=True)
    step()
    prof = {k: round(v[0], 3) for k, v in g.profile_read(reset=True).items() if v[1]}
    g.profile_enable(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        res_n, res_m = step(record=True)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps, t_ex / args.steps, t_so / args.steps,
                      float(n_recv_first[0]), float(outgrown_first[0]), float(run.stats.get("sent", 0))],
                     dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)

    check = None
    if rank == 0 and not args.no_check:
        run.close_buffers() if world == 1 else None
        g1 = api.SoGpu(device=local, stream=stream.cuda_stream)
        g1.set_particles_device(full.data_ptr(), n)
        g1.build_grid()
        d_n1 = torch.empty(max(nh, 1), dtype=torch.int32, device=dev)
        d_m1 = torch.empty(max(nh, 1), dtype=torch.float32, device=dev)
        t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            g1.build_grid()
            g1.so_device(d_c.data_ptr(), d_r.data_ptr(), nh, thr, 8, d_n1.data_ptr(), d_m1.data_ptr())
        torch.cuda.synchronize()
        t1[0].record(stream)
        g1.build_grid()
        g1.so_device(d_c.data_ptr(), d_r.data_ptr(), nh, thr, 8, d_n1.data_ptr(), d_m1.data_ptr())
        t1[1].record(stream)
        torch.cuda.synchronize()
        check = {"rank0_halos": nh,
                 "n_delta_identical": bool(np.array_equal(res_n, d_n1[:nh].cpu().numpy())),
                 "m_delta_bits_identical": bool(res_m.tobytes() == d_m1[:nh].cpu().numpy().tobytes()),
                 "full_grid_on_one_gpu_ms_for_rank0_halos": t1[0].elapsed_time(t1[1])}
        g1.close()
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps({
            "mode": "domain run (strong scaling)", "workload": "%s: %d particles, %d halos" % (name, n, h),
            "n_gpus": world, "transport": args.transport, "mask_balls": args.balls, "steps": args.steps,
            "ms_per_step": float(tmax[0]), "halos_per_s": h / (float(tmax[0]) * 1e-3),
            "exchange_ms_max": float(tmax[1]), "grid_and_solve_ms_max": float(tmax[2]),
            "particles_received_total": float(tsum[3]), "particles_received_max": float(tmax[3]),
            "received_fraction_of_N": float(tsum[3]) / n, "halos_outgrown_total": float(tsum[4]),
            "halos_per_rank": [int((rank_of == r).sum()) for r in range(world)],
            "setup_s": t_gen, "rank0_kernel_ms_one_step": prof, "check_vs_single_full_grid": check}))
    run.close_buffers()
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
