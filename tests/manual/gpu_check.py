"""Development check: CUDA path vs the oracle on scaled BASELINE configs (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from so_b200 import api, synth
from oracle import pyoracle as po


def compare(s, thr=200.0, nmem=8, label=""):
    t0 = time.time()
    g = api.SoGpu()
    g.set_particles(s.pos, s.mass)
    t1 = time.time()
    g.build_grid()
    t2 = time.time()
    g.keep_member_d2(True)
    r = g.so(s.centers, s.rgtp, thr, nmem)
    t3 = time.time()
    off, mem = g.members(sorted=True)
    t4 = time.time()
    st = g.stats()
    print("[%s] N=%d H=%d upload %.3fs build %.3fs so %.3fs members %.3fs stats %s" %
          (label, s.n, s.h, t1 - t0, t2 - t1, t3 - t2, t4 - t3, st), flush=True)
    o = po.Oracle(s.pos, s.mass)
    t5 = time.time()
    ref = o.so(s.centers, s.rgtp, np.float32(thr), nmem)
    t6 = time.time()
    print("   oracle so %.3fs evals %d" % (t6 - t5, ref["nevals"]))
    bad = 0
    for i in range(s.h):
        ok = (r["ndelta"][i] == ref["ndelta"][i] and
              r["mvir"][i].tobytes() == ref["mvir"][i].tobytes() and
              r["rvir"][i].tobytes() == ref["rvir"][i].tobytes())
        if ok and ref["ndelta"][i] > 0:
            a = mem[off[i]:off[i + 1]]
            b = ref["members"][ref["member_offset"][i]:ref["member_offset"][i + 1]]
            ok = np.array_equal(a, b)
        if not ok:
            bad += 1
            if bad <= 10:
                print("   MISMATCH halo %d gpu (n=%d m=%g r=%g) ref (n=%d m=%g r=%g)" %
                      (i, r["ndelta"][i], r["mvir"][i], r["rvir"][i], ref["ndelta"][i], ref["mvir"][i], ref["rvir"][i]))
    print("   mismatches: %d / %d ; codes gpu: %s" % (bad, s.h, np.unique(r["rvir"][r["rvir"] < 0], return_counts=True)), flush=True)
    g.close()
    return bad


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "cfg0"]
    tot = 0
    if "tiny" in which:
        tot += compare(synth.make_snapshot(32 ** 3, 20, seed=5, nmax=2000), label="tiny")
    if "small" in which:
        tot += compare(synth.config(0, 0.125), label="cfg0/8")
    if "cfg0" in which:
        tot += compare(synth.config(0), label="cfg0")
    if "cfg1" in which:
        tot += compare(synth.config(1), label="cfg1")
    if "big" in which:
        tot += compare(synth.make_snapshot(128 ** 3, 8, seed=7, sizes=[3e5, 2e5, 1e5, 5e4, 2e4, 1e4, 5e3, 100], nmax=1e6), label="bighalos")
    print("TOTAL MISMATCHES", tot)
    sys.exit(1 if tot else 0)
