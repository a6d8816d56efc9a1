#!/usr/bin/env python
"""bench.py — halos/sec of the SO hot path on ONE catalog over ONE snapshot (strong scaling).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU code (oracle/_ref)

Workload: BASELINE.json configs[3] — 1024^3 periodic snapshot (1.07 G particles), 100 000 NFW halo centres,
Delta = 200 rho_crit (the configuration north_star's target is stated on; 73 of 180 GB on one GPU).
A "step" is one pass of the hot path over that input:
    kdBuildTree (so.c:515) + kdSO/kdRvir for every halo (so.c:540; R/M/N + member lists),
run as a DOMAIN STEP (so_b200/csrc/domain_step.cuh): every rank holds 1/N of the particle array (as a parallel
reader delivers it) and the whole catalog; per step each rank derives halo ownership and every cell's
destinations on its own device, keeps what some halo can reach of its slice in one streaming pass, stores those
{x,y,z,index} records straight into the owners' buffers over NVLink peer memory (k_route_split), meets the others
at a flag barrier, builds a grid over what arrived and solves its halos.
At N = 1 the same code is the best single-GPU path (routing = compaction to the 5-20 % of the particles any
halo can reach).  There is no collective in the timed region.

`value`  = halos/s with the slices and the catalog already in HBM, CUDA events, max over ranks.
`e2e`    = the same through the public API with HOST buffers: every rank's slice from its own page-locked
           memory over its own PCIe link, catalog H2D, R/M/N (merged on rank 0's side by one small NCCL
           reduction when N > 1) and the member lists D2H.
`cfg1`   = (N = 1) the round-1 workload, BASELINE configs[1] (256^3 / 10 000 halos), for continuity.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "halos/sec"
UNIT = "halos/s"
THR = 200.0
NMEM = 8
N_BALLS = 4
REF_SAMPLE = {3: 1.0 / 64.0, 2: 1.0 / 8.0, 4: 1.0 / 8.0, 1: 1.0, 0: 1.0}   # --impl reference: fraction of the workload per step


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json configs index (default 3: 1024^3)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and H together (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg1", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (default min(steps, 10))")
    ap.add_argument("--ab", action="append", default=[],
                    help="debug: KEY=VALUE environment setting to time as a variant of the device-resident leg (repeatable)")
    ap.add_argument("--ref-scale", type=float, default=0.0,
                    help="--impl reference: fraction of the workload timed per step (default: 1/64 of configs[3])")
    return ap.parse_args()


def workload_name(name, n, h, omega0=1.0):
    return "%s: %d particles, %d halo centres, Delta=200 rho_crit, z=0" % (name, n, h)


# ---------------------------------------------------------------------------------------------
# clocks: sample NVML while the GPU works
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle/_ref/so_ref_timed = the reference's own kdBuildTree +
# kdSO, unmodified objects, timed around the two calls
# ---------------------------------------------------------------------------------------------
def write_inputs(s, tmpdir):
    from so_b200 import tipsy
    snap, gtp = os.path.join(tmpdir, "snap.tipsy"), os.path.join(tmpdir, "halos.gtp")
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    return snap, gtp


def time_reference(s, steps, warmup, tmpdir, full_kdso=False):
    """Seconds per step of the reference's kdBuildTree + kdSO on snapshot s -> (list, kind).
    The reference's kdSO includes kdVcirc, kdTagParticles and _VcmParticles per halo (kd2.c:823-826,884-885)."""
    from oracle import pyoracle as po
    if po.ref_available("so_ref_timed"):
        snap, gtp = write_inputs(s, tmpdir)
        ts = []
        for i in range(warmup + steps):
            t = po.run_so_ref_timed(snap, gtp, THR * s.omega0, NMEM, 1.0)
            if i >= warmup:
                ts.append(t["t_build"] + t["t_so"])
        os.remove(snap)
        os.remove(gtp)
        return ts, "reference"
    # reference binary absent: time the oracle port (cell list build + kdSO restatement)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o = po.Oracle(s.pos, s.mass)
        o.so(s.centers, s.rgtp, np.float32(THR * s.omega0), NMEM, want_members=True)
        o.close()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return ts, "port"


def tmp_dir():
    import tempfile
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    return tempfile.TemporaryDirectory(dir=base)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from so_b200 import synth
    frac = args.ref_scale if args.ref_scale > 0 else REF_SAMPLE.get(args.config, 1.0)
    scale = args.scale * frac
    s = synth.config(args.config, scale)
    with tmp_dir() as tmp:
        ts, kind = time_reference(s, args.steps, args.warmup, tmp)
    sec = float(np.mean(ts))
    value = s.h / sec
    sample = ("%s scaled by %g (%d particles, %d halos) per step; kdBuildTree + kdSO of the reference (kdSO there "
              "also runs kdVcirc, kdTagParticles and _VcmParticles per halo), single-threaded as shipped"
              % (s.name, scale, s.n, s.h))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[%d] (%s)" % (args.config, s.name), "sample": sample,
                   "delta": THR, "n_members": NMEM,
                   "note": "the reference needs ~70 GB and tens of minutes per step on the full 1024^3 configuration; "
                           "its per-halo rate falls with N (kdBuildTree is N log N), so the sampled rate flatters it"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.config != 1 and not args.no_cfg1:
        # the round-1 workload in full, so that one key of the two arms compares identical work
        s1 = synth.config(1, args.scale)
        with tmp_dir() as tmp:
            t1, _ = time_reference(s1, min(args.steps, 3), 0, tmp)
        line["cfg1"] = {"workload": workload_name(s1.name, s1.n, s1.h), "value": s1.h / float(np.mean(t1)), "unit": UNIT,
                        "ms_per_step": float(np.mean(t1)) * 1e3, "steps": len(t1), "sample": "the full configuration"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def kernel_table(prof, steps, units):
    """per-kernel ms / launches / algorithmic bytes per step from the library's event log"""
    alg = {"k_so_query<32>": ("eval", 16.0), "k_so_query<256>": ("eval", 16.0), "k_so_query<1024>": ("eval", 16.0),
           "k_so_emit<32>": ("member", 20.0), "k_so_emit<256>": ("member", 20.0), "k_so_emit<1024>": ("member", 20.0)}
    total_ms = sum(v[0] for v in prof.values())
    kernels = {}
    for kname, (kms, launches, kbytes) in prof.items():
        if launches == 0:
            continue
        per = kms / steps
        ent = {"ms_per_step": per, "share": kms / total_ms if total_ms else 0.0, "launches_per_step": launches / steps}
        if kbytes > 0:
            ent["alg_bytes"] = kbytes / steps
        if "alg_bytes" in ent and per > 0:
            ent["gbs"] = ent["alg_bytes"] / (per * 1e-3) / 1e9
        kernels[kname] = ent
    # the query kernels run side by side and share one evaluation counter: their roofline is taken over the group
    q = [k for k in kernels if k.startswith("k_so_query")]
    if q and units.get("eval"):
        tq = max(kernels[k]["ms_per_step"] for k in q)          # they overlap: the longest one bounds the phase
        for k in q:
            kernels[k]["note"] = "runs concurrently with the other size classes; evaluations are counted for the group"
        kernels["gather(k_so_query*)"] = {"ms_per_step": tq, "alg_bytes": 16.0 * units["eval"],
                                          "gbs": 16.0 * units["eval"] / (tq * 1e-3) / 1e9 if tq > 0 else 0.0,
                                          "note": "16 B x r^2 evaluations over the longest of the concurrent query kernels"}
    del alg
    return kernels


def roofline_entry(kernels, name, peak, peak_src, extra=None):
    k = kernels.get(name)
    if not k or not k.get("gbs"):
        return None
    lps = max(k.get("launches_per_step", 1.0), 1.0)
    r = {"bound": "hbm", "kernel": name, "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
         "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": k["alg_bytes"] / lps,
         "launches_per_step": lps, "avg_launch_ms": k["ms_per_step"] / lps, "share_of_step": k.get("share"),
         "timed": "live CUDA events on the launching stream, second pass over the same K steps"}
    if extra:
        r.update(extra)
    return r


def traffic_for(name, workload):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, None
    tj = json.load(open(tpath))
    for entry in (tj if isinstance(tj, list) else [tj]):
        if entry.get("workload", "") and workload.startswith(entry["workload"]):
            hit = [v["dram_bytes_per_launch"] for k, v in entry["kernels"].items() if k == name] or \
                  [v["dram_bytes_per_launch"] for k, v in entry["kernels"].items() if k.startswith(name)]
            if hit:
                return sum(hit) / len(hit), entry.get("source")
    return None, None


def run_cfg1(torch, api, synth, dev, stream, args):
    """The round-1 step on BASELINE configs[1] (full grid build + solve, and the domain step), one GPU."""
    from so_b200 import parallel
    s = synth.config(1, args.scale)
    g = api.SoGpu(device=dev.index, stream=stream.cuda_stream)
    xyzm = torch.empty((s.n, 4), dtype=torch.float32, device=dev)
    pos_pin = torch.from_numpy(s.pos).pin_memory()
    g.upload_particles(pos_pin.numpy(), s.mass, xyzm.data_ptr())
    g.set_particles_device(xyzm.data_ptr(), s.n)
    d_c = torch.from_numpy(s.centers).to(dev)
    d_r = torch.from_numpy(s.rgtp).to(dev)
    d_n = torch.empty(s.h, dtype=torch.int32, device=dev)
    d_m = torch.empty(s.h, dtype=torch.float32, device=dev)
    thr = np.float32(THR)
    steps = max(args.steps, 5)

    def timed(fn, k):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    def full():
        g.build_grid()
        g.so_device(d_c.data_ptr(), d_r.data_ptr(), s.h, thr, NMEM, d_n.data_ptr(), d_m.data_ptr())
    ms_full = timed(full, steps)
    n_full = d_n.cpu().numpy().copy()
    st = g.stats()

    def e2e():
        g.upload_particles(pos_pin.numpy(), s.mass, xyzm.data_ptr())
        g.set_particles_device(xyzm.data_ptr(), s.n)
        g.build_grid()
        r = g.so(s.centers, s.rgtp, thr, NMEM)
        g.members(copy=False)
        return r
    for _ in range(2):
        e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    # kdSO-equivalent step (what the reference's timed kdSO does): + tagging, vcm needs velocities (not in this
    # synthetic input), kdVcirc on the device
    r = g.so(s.centers, s.rgtp, thr, NMEM)

    def full_kdso():
        g.upload_particles(pos_pin.numpy(), s.mass, xyzm.data_ptr())
        g.set_particles_device(xyzm.data_ptr(), s.n)
        g.build_grid()
        g.keep_member_d2(True)
        rr = g.so(s.centers, s.rgtp, thr, NMEM)
        g.members(sorted=True, copy=False)
        g.tag_members(np.arange(1, s.h + 1, dtype=np.int32))
        ok = rr["rvir"] > 0
        g.vcirc(s.centers[ok], rr["rvir"][ok], rr["mvir"][ok], 1.0, NMEM, profile=False)
        g.keep_member_d2(False)
    kdso_ms = None
    try:
        for _ in range(2):
            full_kdso()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            full_kdso()
        torch.cuda.synchronize()
        kdso_ms = (time.perf_counter() - t0) / 3 * 1e3
    except Exception as e:    # reported, not fatal: this leg is an extra
        kdso_ms = "failed: %s" % e
    g.close()
    # the same workload as a domain step on one GPU
    g2 = api.SoGpu(device=dev.index, stream=stream.cuda_stream)
    ds = parallel.DomainStep(g2, s.n, s.mass)
    ds.step(d_c.data_ptr(), d_r.data_ptr(), s.h, N_BALLS, xyzm.data_ptr(), s.n, 0, thr, NMEM, d_n.data_ptr(), d_m.data_ptr())
    res = ds.result()
    ms_dom = timed(lambda: ds.step(d_c.data_ptr(), d_r.data_ptr(), s.h, N_BALLS, xyzm.data_ptr(), s.n, 0, thr, NMEM,
                                   d_n.data_ptr(), d_m.data_ptr()), steps)
    n_dom = d_n.cpu().numpy()
    inside = n_dom != -103                    # (a halo whose ball leaves the 4-ball mask is re-run with a larger one)
    same = bool(np.array_equal(n_dom[inside], n_full[inside]))
    ds.close()
    g2.close()
    return {"workload": workload_name(s.name, s.n, s.h), "value": s.h / (min(ms_full, ms_dom) * 1e-3), "unit": UNIT,
            "ms_per_step_full_grid": ms_full, "ms_per_step_domain_step": ms_dom,
            "domain_step_same_n_delta_as_full_grid": same, "domain_step_outgrown_-103": int((~inside).sum()),
            "domain_step_flags": res["flags"],
            "evals_per_step": st["last_evals"], "members_per_step": st["last_members"],
            "e2e": {"value": s.h / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 12 * s.n + 16 * s.h, "d2h_bytes_per_step": 16 * s.h + 8 + 4 * st["last_members"]},
            "e2e_full_kdso_equivalent": {"ms_per_step": kdso_ms,
                                         "note": "upload + build + solve + sorted member lists + device tagging + kdVcirc: the work "
                                                 "the reference's kdBuildTree + kdSO does (kd2.c:864-895), minus _VcmParticles "
                                                 "(the synthetic input has no velocities)"}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from so_b200 import api, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    t_setup = time.time()

    # ---- inputs: rank 0 generates; every rank receives ITS slice and the whole (small) catalog ----------
    if rank == 0:
        s = synth.config(args.config, args.scale)
        meta = [s.n, s.h, float(s.mass), s.name, float(s.omega0)]
    else:
        s, meta = None, [0, 0, 0.0, "", 1.0]
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    n, h, mass, name, omega0 = int(meta[0]), int(meta[1]), np.float32(meta[2]), meta[3], float(meta[4])
    thr = np.float32(np.float32(THR) * np.float32(omega0))
    cat = torch.empty((h, 4), dtype=torch.float32, device=dev)
    if rank == 0:
        cat.copy_(torch.from_numpy(np.concatenate([s.centers, s.rgtp[:, None]], axis=1).astype(np.float32)))
    if world > 1:
        dist.broadcast(cat, src=0)
    cat_pin = torch.empty((h, 4), dtype=torch.float32).pin_memory()
    cat_pin.copy_(cat)
    d_centers = cat[:, :3].contiguous()
    d_rgtp = cat[:, 3].contiguous()
    cen_pin = d_centers.cpu().pin_memory()
    rg_pin = d_rgtp.cpu().pin_memory()

    bounds = parallel.slice_bounds(n, world)
    a, b = bounds[rank]
    n_slice = b - a
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    d_slice = torch.empty((n_slice, 4), dtype=torch.float32, device=dev)
    chunk = 1 << 24
    for r in range(world):
        ra, rb = bounds[r]
        for c0 in range(ra, rb, chunk):
            c1 = min(rb, c0 + chunk)
            if rank == 0:
                t = torch.empty((c1 - c0, 4), dtype=torch.float32, device=dev)
                t[:, :3] = torch.from_numpy(s.pos[c0:c1]).to(dev)
                t[:, 3] = float(mass)
                if r == 0:
                    d_slice[c0 - ra:c1 - ra] = t
                else:
                    dist.send(t, dst=r)
                del t
            elif rank == r:
                dist.recv(d_slice[c0 - ra:c1 - ra], src=0)
    torch.cuda.synchronize()
    if rank == 0:
        s.pos = None                       # the slices live on the devices now
    # page-locked host copy of the slice's positions: the input of the end-to-end leg
    pos_pin = torch.empty((n_slice, 3), dtype=torch.float32).pin_memory()
    for c0 in range(0, n_slice, chunk):
        c1 = min(n_slice, c0 + chunk)
        pos_pin[c0:c1].copy_(d_slice[c0:c1, :3])
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    g = api.SoGpu(device=local, stream=stream.cuda_stream)
    frac = 0.9 if args.config == 4 else 0.30          # share of the snapshot the receive buffers are sized for
    rc_, sc_ = parallel.default_caps(n, world, frac)
    ds = parallel.DomainStep(g, n, mass, recv_cap=rc_, stage_cap=sc_)
    d_out_n = torch.empty(h, dtype=torch.int32, device=dev)
    d_out_m = torch.empty(h, dtype=torch.float32, device=dev)

    def step_device():
        ds.step(d_centers.data_ptr(), d_rgtp.data_ptr(), h, N_BALLS, d_slice.data_ptr(), n_slice, a, thr, NMEM,
                d_out_n.data_ptr(), d_out_m.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(xs):
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def check_flags(res, what):
        f = allmax(float(res["flags"]))
        if f:
            raise SystemExit("bench.py: domain step failed (%s): %s" % (what, parallel.flags_text(int(res["flags"])) or f))

    # ---- device-resident leg ---------------------------------------------------------------------
    sampler = ClockSampler(local)
    g.profile_enable(False)
    for _ in range(warmup):
        step_device()
        check_flags(ds.result(), "warm-up")          # (also teaches the library how many records to expect)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_step = allmax(e0.elapsed_time(e1)) / args.steps
    res = ds.result(h)
    check_flags(res, "timed steps")
    owner = res["owner"]
    st = g.stats()
    launches = st["last_kernel_launches"]
    # the SAME K steps once more with the library's per-kernel CUDA events on (two event records per launch
    # perturb the step by a few percent, so the headline above is timed without them)
    g.profile_enable(True)
    g.profile_read(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record(stream)
    for _ in range(args.steps):
        step_device()
    p1.record(stream)
    barrier()
    ms_profiled = p0.elapsed_time(p1) / args.steps
    prof = g.profile_read(reset=True)
    g.profile_enable(False)
    st = g.stats()
    n_recv = ds.result()["n_recv"]
    evals_total, members_total, recv_total, recv_max = allsum([float(st["last_evals"]), float(st["last_members"]),
                                                               float(n_recv), 0.0])[:3] + [allmax(float(n_recv))]

    # results of the whole catalog, merged: every halo is solved by exactly one rank
    code = d_out_n.clone()
    mm = torch.where(code == int(parallel.NOT_MINE), torch.full_like(d_out_m, -float("inf")), d_out_m)
    if world > 1:
        dist.all_reduce(code, op=dist.ReduceOp.MAX)
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    res_n, res_m = code.cpu().numpy(), mm.cpu().numpy()
    ok = res_n > 0
    outgrown = int((res_n == -103).sum())
    k = res_n[ok].astype(np.int64)
    m_ok = bool(np.array_equal(res_m[ok], (api.mass_prefix(mass, k + 1) - mass).astype(np.float32)))
    codes = {int(c): int((res_n == c).sum()) for c in np.unique(res_n[~ok])}

    # ---- e2e: host buffers in, host results out, every step ----------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    out_pin = torch.empty((2, h), dtype=torch.float32).pin_memory()

    def step_e2e():
        d_centers.copy_(cen_pin, non_blocking=True)
        d_rgtp.copy_(rg_pin, non_blocking=True)
        ds.step(d_centers.data_ptr(), d_rgtp.data_ptr(), h, N_BALLS, d_slice.data_ptr(), n_slice, a, thr, NMEM,
                d_out_n.data_ptr(), d_out_m.data_ptr(), host_xyz=pos_pin.data_ptr())
        c2 = d_out_n.clone()
        m2 = torch.where(c2 == int(parallel.NOT_MINE), torch.full_like(d_out_m, -float("inf")), d_out_m)
        if world > 1:
            dist.all_reduce(c2, op=dist.ReduceOp.MAX)
            dist.all_reduce(m2, op=dist.ReduceOp.MAX)
        out_pin[0].copy_(c2.view(torch.float32), non_blocking=True)
        out_pin[1].copy_(m2, non_blocking=True)
        r = ds.result()                              # synchronises
        off, mem = g.members(copy=False)             # member lists of this rank's halos (global indices)
        return r, off

    for _ in range(2):
        r, off = step_e2e()
        check_flags(r, "end-to-end warm-up")
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r, off = step_e2e()
    barrier()
    e2e_s = allmax((time.perf_counter() - t0) / e2e_steps)
    check_flags(r, "end-to-end steps")
    clocks = sampler.stop()
    n_members_mine = int(off[-1])
    h2d, d2h = allsum([12.0 * n_slice + 16.0 * h, 8.0 * h + 8.0 * (h + 1) + 4.0 * n_members_mine])
    halos_per_rank = allsum([float((owner == r_).sum()) if rank == 0 else 0.0 for r_ in range(world)])

    ds.close()
    g.close()
    # ---- debug: variants of the device-resident leg under other tuning switches (same inputs) ----------
    ab = {}
    for spec in args.ab:
        keep = {}
        for kv in spec.split("+"):
            k_, v_ = kv.split("=", 1)
            keep[k_] = os.environ.get(k_)
            os.environ[k_] = v_
        g2 = api.SoGpu(device=local, stream=stream.cuda_stream)
        ds2 = parallel.DomainStep(g2, n, mass, recv_cap=rc_, stage_cap=sc_)

        balls2 = int(os.environ.get("SO_BENCH_BALLS", N_BALLS))

        def step2():
            ds2.step(d_centers.data_ptr(), d_rgtp.data_ptr(), h, balls2, d_slice.data_ptr(), n_slice, a, thr, NMEM,
                     d_out_n.data_ptr(), d_out_m.data_ptr())
        for _ in range(warmup):
            step2()
            check_flags(ds2.result(), "variant warm-up")
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        for _ in range(args.steps):
            step2()
        q1.record(stream)
        barrier()
        ab[spec] = allmax(q0.elapsed_time(q1)) / args.steps
        check_flags(ds2.result(), "variant")
        same = bool(torch.equal(torch.where(d_out_n == int(parallel.NOT_MINE), torch.zeros_like(d_out_n), d_out_n),
                                torch.where(code == int(parallel.NOT_MINE), torch.zeros_like(code), code)) if world == 1 else True)
        ab[spec + " same_n_delta"] = same
        ab[spec + " n_recv"] = allsum([float(ds2.result()["n_recv"])])[0]
        ds2.close()
        g2.close()
        for k_, v_ in keep.items():
            if v_ is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v_
    del d_slice, pos_pin
    torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline entries (rank 0's launches, live CUDA-event times) -------------------------------
    peak, peak_src = peak_hbm()
    units = {"eval": float(st["last_evals"]), "member": float(st["last_members"])}
    kernels = kernel_table(prof, args.steps, units)
    timed_kernels = {k: v for k, v in kernels.items() if not k.startswith("gather(")}
    dom = max(timed_kernels, key=lambda k_: timed_kernels[k_]["ms_per_step"])
    wl = workload_name(name, n, h)
    roofline = roofline_entry(kernels, dom, peak, peak_src) or {"bound": "hbm", "kernel": dom, "achieved": None, "peak": peak,
                                                                 "unit": "GB/s", "frac": None, "traffic": None}
    tr, tr_src = traffic_for(dom, name)
    if tr is not None and args.scale == 1.0:
        roofline["traffic"], roofline["traffic_source"] = tr, tr_src
    e_min = members_total + h
    roof_gather = roofline_entry(kernels, "gather(k_so_query*)", peak, peak_src,
                                 {"evals": units["eval"], "evals_min": float(st["last_members"]) + float((owner == 0).sum()),
                                  "note": "rank 0's halos; E_min = sum(N_Delta + 1)"})
    build_names = [k_ for k_ in ("k_lvl_hist", "k_scan", "k_lvl_partition", "k_bucket_live", "k_bucket_sort", "k_bucket_sort(big)")
                   if k_ in kernels]
    t_build = sum(kernels[k_]["ms_per_step"] for k_ in build_names)
    if "k_bucket_sort" in kernels and "k_bucket_sort(big)" in kernels:
        # the two bucket kernels run side by side on two streams: the longer one is what the step waits for
        t_build -= min(kernels["k_bucket_sort"]["ms_per_step"], kernels["k_bucket_sort(big)"]["ms_per_step"])
    roof_build = None
    if t_build > 0:
        gbs = 36.0 * n_recv / (t_build * 1e-3) / 1e9
        roof_build = {"bound": "hbm", "kernel": "+".join(build_names), "achieved": gbs, "peak": peak, "unit": "GB/s",
                      "frac": gbs / peak, "alg_bytes": 36.0 * n_recv, "ms_per_step": t_build,
                      "note": "SURVEY 8(d): 36 B per particle sorted for the whole build; particles = records this rank "
                              "received; the two concurrent bucket kernels count once (the longer one)"}

    if args.scale == 1.0:
        if roof_gather:
            tg, src = traffic_for("k_so_query_fused", name)
            if tg is not None:
                roof_gather["traffic"], roof_gather["traffic_source"] = tg, src
                roof_gather["traffic_over_16B_evals_min"] = tg / (16.0 * max(roof_gather["evals_min"], 1.0))
        if roof_build:
            parts = [traffic_for(k_, name)[0] for k_ in ("k_lvl_hist<0>", "k_lvl_hist<1>", "k_lvl_partition_rt<0, 0>",
                                                            "k_lvl_partition_rt<1, 1>", "k_bucket_sort_sparse",
                                                            "k_bucket_sort_sparse_warp")]
            if all(p_ is not None for p_ in parts):
                roof_build["traffic"] = sum(parts)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        frac = REF_SAMPLE.get(args.config, 1.0)
        sb = synth.config(args.config, args.scale * frac)
        with tmp_dir() as tmp:
            ts, kind = time_reference(sb, 1, 0, tmp)
        cpu_baseline = {"value": sb.h / ts[0], "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": "%s scaled by %g (%d particles, %d halos) once: kdBuildTree + kdSO of the reference, "
                                  "%.1f s, single-threaded as shipped" % (sb.name, args.scale * frac, sb.n, sb.h, ts[0])}
    cfg1 = None
    if world == 1 and not args.no_cfg1 and args.config != 1:
        cfg1 = run_cfg1(torch, api, synth, dev, stream, args)

    line = {
        "metric": METRIC, "value": h / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "baseline_config": args.config, "delta": THR, "n_members": NMEM,
                   "cells_per_axis": st["cells_per_axis"], "mask_balls": N_BALLS,
                   "halos_per_rank": [int(x) for x in halos_per_rank],
                   "l2": "inputs (%.1f GB of float4 particles per rank) exceed the 126 MB L2; no flush between steps"
                         % (16 * n_slice / 1e9),
                   "parallelism": "domain step over %d rank(s): particle slices, halos owned by spatially compact "
                                  "cost-balanced shares, records pushed over NVLink peer memory, no collective" % world},
        "records_received_total": recv_total, "records_received_max": recv_max, "received_fraction_of_N": recv_total / n,
        "evals_per_s": evals_total / (ms_step * 1e-3), "evals_per_step": evals_total,
        "members_per_step": members_total,
        "check": {"halos_resolved": int(ok.sum()), "codes": codes, "outgrown_-103": outgrown,
                  "m_delta_equals_mass_table_at_n_delta": m_ok},
        # work inflation: r^2 evaluations over the minimum sum(N_Delta + 1)
        "evals_over_min": evals_total / max(e_min, 1.0),
        "e2e": {"value": h / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "roofline": roofline, "roofline_gather": roof_gather, "roofline_build": roof_build,
        "kernels": kernels, "ms_per_step_with_kernel_events": ms_profiled,
        "cpu_baseline": cpu_baseline, "cfg1": cfg1, "setup_s": t_setup, "ab_ms_per_step": ab or None,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if outgrown or not m_ok:
        line["invalid"] = "results failed the built-in checks"
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    # exactly ONE line on stdout: anything a library prints there (e.g. NCCL's version banner)
    # is diverted to stderr while the benchmark runs
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = []
    import builtins
    real_print = builtins.print

    def capture(*args, **kw):
        if kw.get("file") in (None, sys.stdout):
            buf.append(" ".join(str(x) for x in args))
        else:
            real_print(*args, **kw)
    builtins.print = capture
    try:
        rc = run_reference(a) if a.impl == "reference" else run_ours(a)
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in buf:
        print(line)
    sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
