"""Generate the golden fixtures in tests/golden/*.npz by running the REFERENCE built from
/root/reference (oracle/_ref/so_ref, so_ref_inst; recipe: oracle/Makefile).

Run here (the authoring container), commit the .npz files.  Each fixture stores the generator
parameters (inputs are re-created from the seed, guarded by a SHA-1 of the position bytes), the
catalog actually passed, and what the reference produced:
    rvir, mvir      from <out>.sogtp  (kdWriteGTP, kd2.c:1299-1321): eps / mass fields
    ndelta, members from so_ref_inst's dump at kd2.c:823 (j and the first j sorted iOrder)
    igrp            from <out>.sogrp  (kdWriteArray, kd2.c:1256-1258)
    removed/slurped group counters from the .sovcirc stats block (kd2.c:1408-1409)
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from so_b200 import synth, tipsy  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (make_snapshot kwargs, delta, extra catalog spec)
    "basic": dict(gen=dict(n_particles=48 ** 3, n_halos=60, seed=21, nmax=3000), delta=200.0),
    "conflict": dict(gen=dict(n_particles=48 ** 3, n_halos=40, seed=22, nmax=4000, overlap_pairs=8), delta=200.0),
    "omega03": dict(gen=dict(n_particles=60 ** 3, n_halos=6, seed=23, omega0=0.3,
                             sizes=[40000, 20000, 9000, 5000, 800, 60], nmax=1e5), delta=200.0),
    "errors": dict(gen=dict(n_particles=32 ** 3, n_halos=10, seed=24, nmax=1500), delta=200.0, voids=True),
    "never": dict(gen=dict(n_particles=24 ** 3, n_halos=4, seed=25, nmax=600), delta=0.5),
    "members4": dict(gen=dict(n_particles=40 ** 3, n_halos=30, seed=26, nmax=2500), delta=178.0, n_members=4),
}


def catalog(case, s):
    c, r, m = s.centers.copy(), s.rgtp.copy(), s.gtp_mass.copy()
    if case.get("voids"):
        rng = np.random.default_rng(99)
        vc = (rng.random((10, 3)) - 0.5).astype(np.float32)
        vr = np.concatenate([np.full(5, 0.02), np.full(5, 0.06)]).astype(np.float32)
        vm = (np.arange(10) + 1).astype(np.float32) * np.float32(1e-7)
        c, r, m = np.concatenate([c, vc]), np.concatenate([r, vr]), np.concatenate([m, vm])
    return c, r, m


def run_case(name, case, tmp):
    s = synth.make_snapshot(**case["gen"])
    c, r, m = catalog(case, s)
    snap, gtp, out = (os.path.join(tmp, name + e) for e in (".tipsy", ".gtp", ".out"))
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, c, r, m)
    extra = ["-grp", "-gtp", "-O", repr(float(s.omega0))]
    nmem = case.get("n_members", 8)
    if nmem != 8:
        extra += ["-m", str(nmem)]
    inst = os.path.join(tmp, name + ".inst")
    res = po.run_so_ref(snap, gtp, out, delta=case["delta"], extra=extra, inst=True, inst_file=inst)
    rec = po.read_inst_file(inst)
    _, star = tipsy.read_gtp(out + ".sogtp")
    igrp = tipsy.read_sogrp(out + ".sogrp")
    hdr, rows = tipsy.parse_sovcirc(out + ".sovcirc")
    removed = slurped = None
    for line in hdr:
        if "Groups subsumed into larger groups" in line:
            removed = int(line.split(":")[1])
        if "Groups 'slurped'" in line:
            slurped = int(line.split(":")[1])
    h = len(r)
    ndelta = np.zeros(h, np.int32)
    off = np.zeros(h + 1, np.int64)
    mem, md2 = [], []
    for i in range(h):
        if (i + 1) in rec:
            j, order, d2 = rec[i + 1]
            ndelta[i] = j
            mem.append(order)
            md2.append(d2)
        off[i + 1] = off[i] + ndelta[i]
    # fThreshold exactly as so.c:319,480: (float)atof(delta) then *= fOmega (float)
    thr = np.float32(np.float32(case["delta"]) * np.float32(s.omega0))
    rows_full = np.array(rows, dtype=np.float64)          # all 15 columns of every .sovcirc row
    rows = np.array([row[:3] for row in rows], dtype=np.float64)
    sogtp_bytes = np.frombuffer(open(out + ".sogtp", "rb").read(), np.uint8)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        gen=repr(case["gen"]), pos_sha1=hashlib.sha1(s.pos.tobytes()).hexdigest(), mass=s.mass,
        centers=c, rgtp=r, gtp_mass=m, thr=thr, n_members=np.int32(nmem), omega0=np.float64(s.omega0),
        delta=np.float64(case["delta"]),
        rvir=star["eps"].astype(np.float32), mvir_sogtp=star["mass"].astype(np.float32),
        vcm=star["vel"].astype(np.float32),
        sovcirc_idx_m_r=rows, sovcirc_rows=rows_full, sogtp_bytes=sogtp_bytes, ndelta=ndelta, member_offset=off,
        members=np.concatenate(mem) if mem else np.zeros(0, np.int32),
        members_d2=np.concatenate(md2) if md2 else np.zeros(0, np.float32),
        igrp=igrp.astype(np.int32), groups_removed=np.int32(removed), groups_slurped=np.int32(slurped),
        ndist=np.int64(res["ndist"]), ngather=np.int64(res["ngather"]))
    print("%-9s N=%d H=%d  ok=%d  codes=%s removed=%s slurped=%s ndelta max=%d" %
          (name, s.n, h, int((star["eps"] > 0).sum()),
           dict(zip(*np.unique(star["eps"][star["eps"] < 0], return_counts=True))), removed, slurped,
           ndelta.max()))


if __name__ == "__main__":
    if not po.ref_available("so_ref_inst"):
        sys.exit("oracle/_ref/so_ref_inst missing: run `make -C oracle` where /root/reference exists")
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, case in CASES.items():
            if only and name not in only:
                continue
            run_case(name, case, tmp)
