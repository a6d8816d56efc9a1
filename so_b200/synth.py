"""Seeded synthetic inputs: NFW halos in a uniform periodic background.

Box L=1, positions in [-0.5, 0.5), N equal-mass dark particles with m = fl32(Omega0/N) so the
mean density is Omega0 in the reference's rho_crit,0 = 1 units (so.c:477-481).  The halo catalog
(`.gtp` content) carries mass = true M200, pos = centre (+ jitter <= 0.05 R200) and
eps = fRgtp = 0.8 R200, so that the first ball of kdRvir's schedule (1.2 fRgtp, kd2.c:745,767)
does not already contain the answer and the -1 path is not hit.

The generator is plain numpy (PCG64) and is used identically by tests, bench.py and the
reference arm, so every implementation sees bit-identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Snapshot:
    pos: np.ndarray        # (N,3) float32 in [-0.5,0.5)
    mass: np.float32       # per-particle mass
    omega0: float
    time: float            # tipsy header time = 1/(1+z)
    centers: np.ndarray    # (H,3) float32 catalog centres
    rgtp: np.ndarray       # (H,) float32 catalog radii (fRgtp)
    gtp_mass: np.ndarray   # (H,) float32 catalog masses, all distinct
    n200: np.ndarray       # (H,) int64 intended particle count inside R200
    r200: np.ndarray       # (H,) float64 intended R200
    name: str = ""

    @property
    def n(self) -> int:
        return len(self.pos)

    @property
    def h(self) -> int:
        return len(self.centers)


def _mu(x):
    return np.log1p(x) - x / (1.0 + x)


def _nfw_radii(rng, n, c, rmax_over_rs):
    """Radii (in units of rs) of n particles from an NFW profile truncated at rmax_over_rs."""
    xs = np.concatenate([[0.0], np.geomspace(1e-4, rmax_over_rs, 4095)])
    cdf = _mu(xs)
    cdf /= cdf[-1]
    u = rng.random(n)
    return np.interp(u, cdf, xs)


def _unit_vectors(rng, n):
    z = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([s * np.cos(ph), s * np.sin(ph), z], axis=1)


def _wrap(x):
    x = x - np.floor(x + 0.5)
    x32 = x.astype(np.float32)
    x32[x32 >= np.float32(0.5)] = np.float32(-0.5)
    x32[x32 < np.float32(-0.5)] = np.float32(-0.5)
    return x32


def halo_sizes(rng, h, nmin, nmax, slope=-0.9):
    """N200 drawn from dn/dlnM ∝ M^slope between nmin and nmax particles (continuous)."""
    u = rng.random(h)
    if abs(slope) < 1e-9:
        return nmin * (nmax / nmin) ** u
    a, b = float(nmin) ** slope, float(nmax) ** slope
    return (a + u * (b - a)) ** (1.0 / slope)


def make_snapshot(n_particles, n_halos, seed, *, nmin=20, nmax=2.0e4, slope=-0.9, omega0=1.0,
                  delta=200.0, z=0.0, sizes=None, trunc=2.0, overlap_pairs=0, shuffle=True,
                  rgtp_factor=0.8, name=""):
    """Build a Snapshot.  `sizes` (array of N200) overrides the mass function.

    overlap_pairs > 0 moves that many halos next to a bigger neighbour (centre distance
    0.6..1.6 R200 of the bigger one) to exercise the subsume / ignore / slurp rules of
    kdTagParticles (kd2.c:663-720)."""
    rng = np.random.default_rng(seed)
    n_particles = int(n_particles)
    m = np.float32(omega0 / n_particles)
    if sizes is None:
        msz = halo_sizes(rng, n_halos, nmin, nmax, slope)
    else:
        msz = np.asarray(sizes, dtype=np.float64)
        n_halos = len(msz)
    n200 = np.maximum(nmin, np.rint(msz)).astype(np.int64)
    rho_bar = omega0  # total mass omega0 in unit volume
    r200 = (3.0 * n200 * float(m) / (4.0 * np.pi * delta * rho_bar)) ** (1.0 / 3.0)
    conc = np.clip(10.0 * (n200 / 100.0) ** -0.1, 4.0, 10.0)
    ntot = np.rint(n200 * _mu(trunc * conc) / _mu(conc)).astype(np.int64)
    if ntot.sum() > 0.8 * n_particles:
        raise ValueError("halos hold %d of %d particles; lower n_halos/nmax" % (ntot.sum(), n_particles))

    # centres on a jittered lattice: spacing guarantees > 2 (R_i + R_j) separation of the
    # truncated halos unless two of the very largest land on adjacent sites (checked below)
    ns = int(np.ceil(n_halos ** (1.0 / 3.0)))
    while ns ** 3 < n_halos:
        ns += 1
    spacing = 1.0 / ns
    sites = rng.permutation(ns ** 3)[:n_halos]
    ijk = np.stack([sites // (ns * ns), (sites // ns) % ns, sites % ns], axis=1).astype(np.float64)
    rmax = trunc * r200
    room = np.maximum(0.0, 0.5 * spacing - 2.0 * rmax)  # jitter that keeps the separation
    jit = (rng.random((n_halos, 3)) * 2.0 - 1.0) * room[:, None]
    true_c = (ijk + 0.5) * spacing - 0.5 + jit

    if overlap_pairs:
        order = np.argsort(-n200)
        big = order[:overlap_pairs]
        small = order[len(order) // 2: len(order) // 2 + overlap_pairs]
        frac = rng.uniform(0.6, 1.6, overlap_pairs)
        true_c[small] = true_c[big] + _unit_vectors(rng, overlap_pairs) * (frac * r200[big])[:, None]

    # particles
    parts = []
    for i in range(n_halos):
        k = int(ntot[i])
        rs = r200[i] / conc[i]
        r = _nfw_radii(rng, k, conc[i], trunc * conc[i]) * rs
        parts.append(true_c[i] + _unit_vectors(rng, k) * r[:, None])
    n_bg = n_particles - int(ntot.sum())
    pos = np.empty((n_particles, 3), dtype=np.float32)
    off = 0
    for p in parts:
        pos[off:off + len(p)] = _wrap(p)
        off += len(p)
    del parts
    step = 1 << 22
    while off < n_particles:
        k = min(step, n_particles - off)
        pos[off:off + k] = _wrap(rng.random((k, 3)) - 0.5)
        off += k
    if shuffle == "blocks":
        # file order with spatially coherent runs: blocks of 2^16 particles in a seeded random order (the halo
        # blocks end up spread over the whole file, hence over the slices of a multi-GPU run); a full
        # particle-level permutation of 10^9 records costs minutes of host time and changes nothing downstream
        bs = 1 << 16
        nblk = n_particles // bs
        order = np.random.default_rng(seed + 77).permutation(nblk)
        out = np.empty_like(pos)
        for k, src in enumerate(order):
            out[k * bs:(k + 1) * bs] = pos[src * bs:(src + 1) * bs]
        out[nblk * bs:] = pos[nblk * bs:]
        pos = out
    elif shuffle:
        perm = rng.permutation(n_particles)
        pos = pos[perm]

    cat_c = true_c + _unit_vectors(rng, n_halos) * (rng.random(n_halos) * 0.05 * r200)[:, None]
    centers = _wrap(cat_c)
    rgtp = (rgtp_factor * r200).astype(np.float32)
    gmass = (n200 * float(m) * (1.0 + 1e-3 * rng.random(n_halos))).astype(np.float32)
    # indexx (nr.c:91-151) is unstable for equal keys: make the catalogue masses distinct
    o = np.argsort(gmass, kind="stable")
    for a, b in zip(o[:-1], o[1:]):
        if gmass[b] <= gmass[a]:
            gmass[b] = np.nextafter(gmass[a], np.float32(np.inf))
    return Snapshot(pos=pos, mass=m, omega0=omega0, time=1.0 / (1.0 + z), centers=centers,
                    rgtp=rgtp, gtp_mass=gmass, n200=n200, r200=r200, name=name)


# --- the BASELINE.json configurations ---------------------------------------

def config(idx: int, scale: float = 1.0, big_shuffle="blocks") -> Snapshot:
    """BASELINE.json configs[idx]; `scale` < 1 shrinks N and H together for smoke runs.  Snapshots of 512^3
    particles and more are shuffled block-wise (see make_snapshot) instead of particle-wise."""
    if idx in (2, 3, 4) and 512 ** 3 * (8 if idx == 3 else 1) * scale >= 512 ** 3:
        if idx == 2:
            return make_snapshot(int(512 ** 3 * scale), max(1, int(50000 * scale)), seed=1002, omega0=0.3, z=0.5,
                                 shuffle=big_shuffle, name="cfg2_512^3_50000halos")
        if idx == 3:
            return make_snapshot(int(1024 ** 3 * scale), max(1, int(100000 * scale)), seed=1003,
                                 shuffle=big_shuffle, name="cfg3_1024^3_100000halos")
        sizes = np.concatenate([np.full(64, 1.0e6 * scale), np.full(436, 3.0e4 * scale)])
        return make_snapshot(int(512 ** 3 * scale), 500, seed=1004, sizes=np.maximum(sizes, 20), nmax=1e6, trunc=1.3,
                             shuffle=big_shuffle, name="cfg4_512^3_clusterheavy")
    if idx == 0:
        return make_snapshot(int(128 ** 3 * scale), max(1, int(1000 * scale)), seed=1000,
                             name="cfg0_128^3_1000halos")
    if idx == 1:
        return make_snapshot(int(256 ** 3 * scale), max(1, int(10000 * scale)), seed=1001,
                             name="cfg1_256^3_10000halos")
    if idx == 2:
        return make_snapshot(int(512 ** 3 * scale), max(1, int(50000 * scale)), seed=1002,
                             omega0=0.3, z=0.5, name="cfg2_512^3_50000halos")
    if idx == 3:
        return make_snapshot(int(1024 ** 3 * scale), max(1, int(100000 * scale)), seed=1003,
                             name="cfg3_1024^3_100000halos")
    if idx == 4:  # 5a of SURVEY §8d: 512^3, 64 x 1e6 + 436 x 3e4
        n = int(512 ** 3 * scale)
        sizes = np.concatenate([np.full(64, 1.0e6 * scale), np.full(436, 3.0e4 * scale)])
        return make_snapshot(n, 500, seed=1004, sizes=np.maximum(sizes, 20), nmax=1e6, trunc=1.3,
                             name="cfg4_512^3_clusterheavy")
    raise ValueError(idx)
