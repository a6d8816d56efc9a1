"""Tipsy snapshot / .gtp catalog / SO output file formats (numpy, host side).

Byte layouts follow the reference's own structs and readers/writers:
  * header ``struct dump``  : /root/reference/tipsydefs.h:41-48 (32 bytes native: f8 time,
    5 x i4, 4 pad bytes); ``-std`` XDR header = big-endian f8 + 6 x i4, kd2.c:32-44.
  * gas / dark / star records: tipsydefs.h:6-37 (12 / 9 / 11 float32).
  * .gtp catalog = star-only tipsy file, mass = catalog mass, pos = centre, eps = fRgtp
    (kd2.c:220-223, 268-271).
  * .sogtp written by kdWriteGTP (kd2.c:1267-1332), .sogrp by kdWriteArray (kd2.c:1244-1264).

Used by the synthetic-data generator, the tests and bench.py.  The compiled drop-in `so`
(so_b200/host) has its own C readers/writers; the two are cross-checked in tests.
"""
from __future__ import annotations

import numpy as np

GAS_DT = np.dtype([("mass", "<f4"), ("pos", "<f4", 3), ("vel", "<f4", 3), ("rho", "<f4"),
                   ("temp", "<f4"), ("hsmooth", "<f4"), ("metals", "<f4"), ("phi", "<f4")])
DARK_DT = np.dtype([("mass", "<f4"), ("pos", "<f4", 3), ("vel", "<f4", 3), ("eps", "<f4"),
                    ("phi", "<f4")])
STAR_DT = np.dtype([("mass", "<f4"), ("pos", "<f4", 3), ("vel", "<f4", 3), ("metals", "<f4"),
                    ("tform", "<f4"), ("eps", "<f4"), ("phi", "<f4")])
HEADER_DT = np.dtype([("time", "<f8"), ("nbodies", "<i4"), ("ndim", "<i4"), ("nsph", "<i4"),
                      ("ndark", "<i4"), ("nstar", "<i4"), ("pad", "<i4")])
assert HEADER_DT.itemsize == 32 and DARK_DT.itemsize == 36 and STAR_DT.itemsize == 44
assert GAS_DT.itemsize == 48


def _be(dt: np.dtype) -> np.dtype:
    return dt.newbyteorder(">")


def write_tipsy(path, time, gas=None, dark=None, star=None, standard=False):
    """Write a tipsy snapshot.  gas/dark/star are structured arrays (or None)."""
    ng = 0 if gas is None else len(gas)
    nd = 0 if dark is None else len(dark)
    ns = 0 if star is None else len(star)
    h = np.zeros(1, HEADER_DT)
    h["time"] = time
    h["nbodies"] = ng + nd + ns
    h["ndim"] = 3
    h["nsph"], h["ndark"], h["nstar"] = ng, nd, ns
    with open(path, "wb") as f:
        if standard:
            f.write(h.astype(_be(HEADER_DT)).tobytes())
        else:
            f.write(h.tobytes())
        for arr, dt in ((gas, GAS_DT), (dark, DARK_DT), (star, STAR_DT)):
            if arr is None or len(arr) == 0:
                continue
            arr = np.ascontiguousarray(arr, dtype=dt)
            if standard:
                arr.astype(_be(dt)).tofile(f)
            else:
                arr.tofile(f)


def read_tipsy(path, standard=False):
    """Return (header dict, gas, dark, star) structured arrays (native byte order)."""
    with open(path, "rb") as f:
        hd = _be(HEADER_DT) if standard else HEADER_DT
        h = np.frombuffer(f.read(32), hd, 1)[0]
        out = []
        for n, dt in ((int(h["nsph"]), GAS_DT), (int(h["ndark"]), DARK_DT),
                      (int(h["nstar"]), STAR_DT)):
            d = _be(dt) if standard else dt
            a = np.fromfile(f, d, n)
            out.append(a.astype(dt) if standard else a)
    hdr = {k: h[k].item() for k in ("time", "nbodies", "ndim", "nsph", "ndark", "nstar")}
    return hdr, out[0], out[1], out[2]


def dark_from_arrays(pos, mass, vel=None, eps=0.0, phi=None):
    n = len(pos)
    d = np.zeros(n, DARK_DT)
    d["mass"] = mass
    d["pos"] = pos
    if vel is not None:
        d["vel"] = vel
    d["eps"] = eps
    if phi is not None:
        d["phi"] = phi
    return d


def write_gtp(path, time, centers, rgtp, gtp_mass, standard=False):
    """Halo catalog in the form kdReadGTPList expects (kd2.c:171-284)."""
    s = np.zeros(len(centers), STAR_DT)
    s["mass"] = gtp_mass
    s["pos"] = centers
    s["eps"] = rgtp
    write_tipsy(path, time, star=s, standard=standard)


def read_gtp(path, standard=False):
    hdr, gas, dark, star = read_tipsy(path, standard)
    if len(gas) or len(dark):
        raise ValueError("FILE TYPE MISMATCH: GTP file contains non-star particles!")
    return hdr, star


def read_sogrp(path):
    """`.sogrp` tipsy array: N then one iGrp per particle in file order (kd2.c:1256-1258)."""
    a = np.loadtxt(path, dtype=np.int64)
    n = int(a[0])
    g = a[1:].astype(np.int32)
    assert len(g) == n
    return g


def parse_sovcirc(path):
    """Return (header lines, rows) of a `.sovcirc` file; rows = list of float lists."""
    hdr, rows = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                hdr.append(line.rstrip("\n"))
            elif line.strip():
                rows.append([float(t) for t in line.split()])
    return hdr, rows
