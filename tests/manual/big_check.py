"""Large-configuration check (run under gpurun): BASELINE configs[2..4] through the C-ABI, with
size-independent properties and the oracle on a random subsample of halos."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from so_b200 import api, synth


def run(idx, scale, oracle_sample=200, thr=200.0):
    t0 = time.time()
    kw = {}
    s = synth.config(idx, scale) if scale != 1.0 or idx < 2 else None
    if s is None:
        # big ones: skip the global shuffle (memory/time); halo blocks + uniform background
        if idx == 2:
            s = synth.make_snapshot(512 ** 3, 50000, seed=1002, omega0=0.3, z=0.5, shuffle=False, name="cfg2_512^3")
        elif idx == 3:
            s = synth.make_snapshot(1024 ** 3, 100000, seed=1003, shuffle=False, name="cfg3_1024^3")
        elif idx == 4:
            sizes = np.concatenate([np.full(64, 1.0e6), np.full(436, 3.0e4)])
            s = synth.make_snapshot(512 ** 3, 500, seed=1004, sizes=sizes, nmax=1e6, trunc=1.3, shuffle=False, name="cfg4_cluster")
    thr = np.float32(np.float32(thr) * np.float32(s.omega0))
    print("[%s] generated N=%d H=%d in %.1fs" % (s.name, s.n, s.h, time.time() - t0), flush=True)
    g = api.SoGpu()
    g.profile_enable(True)
    t1 = time.time()
    g.set_particles(s.pos, s.mass)
    t2 = time.time()
    g.build_grid()
    st0 = g.stats()
    t3 = time.time()
    r = g.so(s.centers, s.rgtp, thr)
    t4 = time.time()
    off, mem = g.members(copy=False)
    t5 = time.time()
    st = g.stats()
    prof = g.profile_read()
    print("   upload %.3fs build %.3fs so %.3fs members %.3fs | cells/axis %d evals %d members %d" %
          (t2 - t1, t3 - t2, t4 - t3, t5 - t4, st["cells_per_axis"], st["last_evals"], st["last_members"]))
    for k, (ms, ln, by) in prof.items():
        if ln:
            print("     %-20s %9.3f ms  %3d launches  %s" % (k, ms, ln, ("%.0f GB/s" % (by / ms / 1e6)) if by else ""))
    ok = r["ndelta"] > 0
    print("   resolved %d / %d ; codes %s ; N_delta max %d" %
          (ok.sum(), s.h, dict(zip(*np.unique(r["rvir"][~ok], return_counts=True))), r["ndelta"].max()))
    assert np.array_equal(np.diff(off), np.where(ok, r["ndelta"], 0))
    k = r["ndelta"][ok].astype(np.int64)
    assert np.array_equal(r["mvir"][ok], (api.mass_prefix(s.mass, k + 1) - s.mass).astype(np.float32)), "M != S[N+1]-m"
    # members: unique, and every member closer than every non-member is implied by the key test; check counts
    for i in np.nonzero(ok)[0][:: max(1, int(ok.sum() // 300))]:
        seg = mem[off[i]:off[i + 1]]
        assert len(np.unique(seg)) == len(seg)
    if oracle_sample:
        from oracle import pyoracle as po
        t6 = time.time()
        o = po.Oracle(s.pos, s.mass)
        pick = np.random.default_rng(0).choice(s.h, min(oracle_sample, s.h), replace=False)
        ref = o.so(s.centers[pick], s.rgtp[pick], thr, 8)
        bad = 0
        for n, i in enumerate(pick):
            same = (r["ndelta"][i] == ref["ndelta"][n] and r["mvir"][i].tobytes() == ref["mvir"][n].tobytes() and
                    r["rvir"][i].tobytes() == ref["rvir"][n].tobytes())
            if same and ref["ndelta"][n] > 0:
                a = np.sort(mem[off[i]:off[i + 1]])
                b = np.sort(ref["members"][ref["member_offset"][n]:ref["member_offset"][n + 1]])
                same = np.array_equal(a, b)
            bad += not same
        print("   oracle subsample: %d mismatches of %d (%.1fs)" % (bad, len(pick), time.time() - t6))
        assert bad == 0
    g.close()
    print("   OK", flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        run(int(parts[0]), float(parts[1]) if len(parts) > 1 else 1.0,
            int(parts[2]) if len(parts) > 2 else 200)
