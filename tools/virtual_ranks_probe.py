"""What one rank of an R-rank domain step costs, measured on ONE GPU: the R ranks run one after the other as
handles of one process (so_b200.parallel.VirtualDomainStep's protocol), with the library's per-kernel CUDA
events on.  The push is a local copy here and there is no peer barrier, everything else is the kernel sequence a
rank of the real run executes on its 1/R slice.

    python tools/virtual_ranks_probe.py [--ranks 8] [--config 3] [--steps 3]

Prints, per kernel: the mean and the maximum over ranks of the time per step, the single-rank (R = 1) time / R it
would have under perfect scaling when --with-single is given, and per rank the phases begin+route / push / solve (host clock around a
synchronised phase: includes the launch latencies an idle device exposes)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--balls", type=int, default=4)
    args = ap.parse_args()
    import torch
    from so_b200 import parallel, synth

    dev = torch.device("cuda", 0)
    s = synth.config(args.config, args.scale)
    R = args.ranks
    thr = np.float32(np.float32(200.0) * np.float32(s.omega0))
    slices, base = [], []
    for a, b in parallel.slice_bounds(s.n, R):
        t = torch.empty((b - a, 4), dtype=torch.float32, device=dev)
        for c0 in range(a, b, 1 << 26):
            c1 = min(b, c0 + (1 << 26))
            t[c0 - a:c1 - a, :3] = torch.from_numpy(s.pos[c0:c1]).to(dev)
        t[:, 3] = float(s.mass)
        slices.append(t)
        base.append(a)
    v = parallel.VirtualDomainStep(R, s.n, s.mass, frac=0.9 if args.config == 4 else 0.30)
    cen = torch.from_numpy(np.ascontiguousarray(s.centers, np.float32)).to(dev)
    rg = torch.from_numpy(np.ascontiguousarray(s.rgtp, np.float32)).to(dev)
    outs = [(torch.empty(s.h, dtype=torch.int32, device=dev), torch.empty(s.h, dtype=torch.float32, device=dev)) for _ in range(R)]
    phases = np.zeros((R, 3))

    def step(timed):
        import time
        torch.cuda.synchronize()
        t = np.zeros((R, 3))
        for r, g in enumerate(v.gs):
            t0 = time.perf_counter()
            g.domain_begin(cen.data_ptr(), rg.data_ptr(), s.h, args.balls)
            g.domain_route(slices[r].data_ptr(), len(slices[r]), base[r])
            torch.cuda.synchronize()
            t[r, 0] = time.perf_counter() - t0
        for r, g in enumerate(v.gs):
            t0 = time.perf_counter()
            g.domain_push(barrier=False)
            torch.cuda.synchronize()
            t[r, 1] = time.perf_counter() - t0
        for r, g in enumerate(v.gs):
            t0 = time.perf_counter()
            g.domain_solve(thr, 8, outs[r][0].data_ptr(), outs[r][1].data_ptr())
            torch.cuda.synchronize()
            t[r, 2] = time.perf_counter() - t0
        if timed:
            phases[:] += t * 1e3

    for _ in range(3):
        step(False)
    for g in v.gs:
        res = g.domain_result(s.h)
        assert not res["flags"], parallel.flags_text(res["flags"])
        g.profile_enable(True)
        g.profile_read(reset=True)
    for _ in range(args.steps):
        step(True)
    prof = [g.profile_read(reset=True) for g in v.gs]
    names = sorted({k for p in prof for k, x in p.items() if x[1] > 0})
    table = {}
    for k in names:
        ms = np.array([p.get(k, (0.0, 0, 0.0))[0] for p in prof]) / args.steps
        table[k] = {"mean_ms": round(float(ms.mean()), 4), "max_ms": round(float(ms.max()), 4),
                    "launches": int(prof[0][k][1] // args.steps)}
    phases /= args.steps
    n_recv = [int(g.domain_result(s.h)["n_recv"]) for g in v.gs]
    st = [g.stats() for g in v.gs]
    out = {"workload": s.name, "ranks": R, "kernels": table,
           "phase_ms_per_rank": {"begin+route": [round(x, 3) for x in phases[:, 0]], "push": [round(x, 3) for x in phases[:, 1]],
                                 "solve": [round(x, 3) for x in phases[:, 2]]},
           "sum_of_phase_maxima_ms": round(float(phases.max(axis=0).sum()), 3),
           "n_recv": n_recv, "launches_per_step": [int(x["last_kernel_launches"]) for x in st]}
    print(json.dumps(out))
    v.close()


if __name__ == "__main__":
    main()
