"""Parity of the CUDA path (through the C-ABI) with the oracle and with the reference's golden
outputs.  Integer/index results and the fp32 outputs are compared BIT-EXACT; the north-star
tolerance for R_Delta / M_Delta is 1e-6 relative, which bit equality satisfies trivially."""
import numpy as np
import pytest

from oracle import pyoracle as po
from so_b200 import api, synth
from tests.util import GOLDEN_CASES, assert_so_equal, load_golden

pytestmark = pytest.mark.gpu


def run_gpu(pos, mass, centers, rgtp, thr, n_members=8, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0),
            ppc=None, first_ball=None, tma=False):
    g = api.SoGpu()
    if tma:
        g.set_tma_staging(True)
    if ppc:
        g.set_cell_occupancy(ppc)
    if first_ball:
        g.set_first_ball(first_ball)
    g.set_particles(pos, mass, period, center)
    g.build_grid()
    g.keep_member_d2(True)
    r = g.so(centers, rgtp, thr, n_members)
    r["member_offset"], r["members"], r["members_d2"] = g.members(want_d2=True, sorted=True)
    r["stats"] = g.stats()
    g.close()
    return r


def check_against_oracle(pos, mass, centers, rgtp, thr, n_members=8, period=(1.0, 1.0, 1.0), **kw):
    r = run_gpu(pos, mass, centers, rgtp, thr, n_members, period, **kw)
    ref = po.Oracle(pos, mass, period).so(centers, rgtp, np.float32(thr), n_members)
    assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
    assert np.array_equal(r["member_offset"], ref["member_offset"])
    assert np.array_equal(r["members"], ref["members"])          # same (r^2, index) order
    assert r["stats"]["last_members"] == int(ref["member_offset"][-1])
    return r, ref


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_reference_outputs(name):
    """CUDA path vs what the reference binary itself produced (tests/golden)."""
    s, g = load_golden(name)
    r = run_gpu(s.pos, s.mass, g["centers"], g["rgtp"], g["thr"], int(g["n_members"]))
    sub = g["rvir"] <= -10.0
    err = (g["rvir"] < 0) & ~sub
    ok = ~err & ~sub
    assert np.array_equal(r["rvir"][err], g["rvir"][err]) and np.array_equal(r["mvir"][err], g["rvir"][err])
    assert r["rvir"][ok].tobytes() == g["rvir"][ok].tobytes()
    assert r["mvir"][ok].tobytes() == g["mvir_sogtp"][ok].tobytes()
    assert np.array_equal(r["ndelta"], g["ndelta"])
    for i in range(len(g["rgtp"])):
        a = r["members"][r["member_offset"][i]:r["member_offset"][i + 1]]
        b = g["members"][g["member_offset"][i]:g["member_offset"][i + 1]]
        assert np.array_equal(np.sort(a), np.sort(b)), "halo %d" % i


@pytest.mark.parametrize("seed,n,h,nmax", [(1, 32 ** 3, 20, 2000), (2, 64 ** 3, 200, 8000), (3, 100000, 64, 20000)])
def test_seeded_snapshots_vs_oracle(seed, n, h, nmax):
    s = synth.make_snapshot(n, h, seed=seed, nmax=nmax)
    check_against_oracle(s.pos, s.mass, s.centers, s.rgtp, 200.0)


def test_config0_full_size_vs_oracle():
    """BASELINE.json configs[0]: 128^3, 1000 halos, Delta = 200 rho_crit, z = 0."""
    s = synth.config(0)
    r, ref = check_against_oracle(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    assert (ref["rvir"] > 0).all()


def test_non_power_of_two_mass_sequential_sum():
    """m = 0.3/N is not a power of two: M_Delta must be the SEQUENTIAL fp32 sum (SURVEY trap #1)."""
    s = synth.make_snapshot(60 ** 3, 5, seed=31, omega0=0.3, sizes=[50000, 20000, 7000, 900, 50], nmax=1e5)
    thr = np.float32(np.float32(200.0) * np.float32(0.3))
    r, ref = check_against_oracle(s.pos, s.mass, s.centers, s.rgtp, thr)
    assert ref["ndelta"].max() > 30000


@pytest.mark.parametrize("tma", [False, True])
def test_large_halos_use_block_kernel_and_refinement(tma):
    """Cluster-size halos: the 1024-thread class with per-thread loads, and with the TMA bulk-copy ring
    (sogpu_set_tma_staging), incl. a halo on the periodic boundary (row segments that wrap)."""
    s = synth.make_snapshot(128 ** 3, 6, seed=32, sizes=[400000, 150000, 60000, 20000, 3000, 100], nmax=1e6)
    shift = np.array([0.4999 - s.centers[0, 0], -0.4999 - s.centers[0, 1], 0.0], np.float64)
    pos = synth._wrap(s.pos.astype(np.float64) + shift).astype(np.float32)
    centers = synth._wrap(s.centers.astype(np.float64) + shift).astype(np.float32)
    r, ref = check_against_oracle(pos, s.mass, centers, s.rgtp, 200.0, tma=tma)
    assert ref["ndelta"].max() > 300000


def test_centers_on_the_periodic_boundary():
    s = synth.make_snapshot(48 ** 3, 30, seed=33, nmax=3000)
    # shift everything so that halos straddle the box faces, re-wrap into [-0.5,0.5)
    shift = np.array([0.5 - s.centers[0, 0], 0.5 - s.centers[1, 1], 0.5 - s.centers[2, 2]], np.float64)
    pos = synth._wrap(s.pos.astype(np.float64) + shift)
    cen = synth._wrap(s.centers.astype(np.float64) + shift)
    check_against_oracle(pos, s.mass, cen, s.rgtp, 200.0)
    # centres given outside the box (one period away) must give the same member sets
    cen2 = cen.copy()
    cen2[::2, 0] += np.float32(1.0)
    check_against_oracle(pos, s.mass, cen2, s.rgtp, 200.0)


def test_other_period_and_grid_center():
    rng = np.random.default_rng(5)
    s = synth.make_snapshot(40 ** 3, 20, seed=34, nmax=2500)
    L = 2.5
    pos = (s.pos * np.float32(L)).astype(np.float32)
    cen = (s.centers * np.float32(L)).astype(np.float32)
    rg = (s.rgtp * np.float32(L)).astype(np.float32)
    thr = np.float32(200.0 / L ** 3)
    check_against_oracle(pos, s.mass, cen, rg, thr, period=(L, L, L))
    # data in [0,L): grid centred at L/2, and (second run) left at the default 0 -> cells wrap
    pos2 = (pos + np.float32(L / 2)).astype(np.float32)
    pos2[pos2 >= np.float32(L)] -= np.float32(L)
    cen2 = (cen + np.float32(L / 2)).astype(np.float32)
    check_against_oracle(pos2, s.mass, cen2, rg, thr, period=(L, L, L), center=(L / 2, L / 2, L / 2))
    check_against_oracle(pos2, s.mass, cen2, rg, thr, period=(L, L, L))
    del rng


@pytest.mark.parametrize("nmem", [2, 4, 8, 16, 64])
def test_nmembers_and_error_codes(nmem):
    s = synth.make_snapshot(32 ** 3, 12, seed=35, nmax=1500)
    rng = np.random.default_rng(7)
    vc = (rng.random((12, 3)) - 0.5).astype(np.float32)
    vr = np.concatenate([np.full(6, 0.015), np.full(6, 0.07)]).astype(np.float32)
    centers = np.concatenate([s.centers, vc])
    rgtp = np.concatenate([s.rgtp, vr])
    r, ref = check_against_oracle(s.pos, s.mass, centers, rgtp, 200.0, n_members=nmem)
    assert set(np.unique(ref["rvir"][ref["rvir"] < 0])) <= {-1.0, -2.0, -3.0}


@pytest.mark.parametrize("first_ball", [1, 2, 3, 7, 40])
def test_ball_schedule_subset_gives_the_reference_results(first_ball):
    """The library gathers a subset of kdRvir's ball schedule (kd2.c:765-768): starting at any ball, and
    jumping ahead, must reproduce what the reference finds walking every ball, including the -1 test on
    the schedule's FIRST ball, the -2 test and -3 at the schedule's LAST ball."""
    s = synth.make_snapshot(40 ** 3, 30, seed=51, nmax=3000)
    rng = np.random.default_rng(9)
    vc = (rng.random((16, 3)) - 0.5).astype(np.float32)
    vr = np.concatenate([np.full(6, 0.004), np.full(5, 0.02), np.full(5, 0.08)]).astype(np.float32)
    centers = np.concatenate([s.centers, vc, s.centers[:8]])
    rgtp = np.concatenate([s.rgtp, vr, s.rgtp[:8] * np.float32(0.3)])     # small first guesses: many steps
    for thr in (200.0, 20.0, 0.7):
        r, ref = check_against_oracle(s.pos, s.mass, centers, rgtp, thr, first_ball=first_ball)
    assert (ref["rvir"] == -3.0).any()


def test_threshold_never_reached_gives_minus3():
    s = synth.make_snapshot(24 ** 3, 4, seed=36, nmax=600)
    r, ref = check_against_oracle(s.pos, s.mass, s.centers, s.rgtp, 0.5)
    assert (r["rvir"] == -3.0).all() and (r["mvir"] == -3.0).all() and (r["ndelta"] == 0).all()


def test_cell_size_does_not_change_results():
    s = synth.make_snapshot(48 ** 3, 40, seed=37, nmax=5000)
    base = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    for ppc in (0.25, 8.0, 64.0):
        r = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0, ppc=ppc)
        assert_so_equal(r, base["rvir"], base["mvir"], base["ndelta"])
        assert np.array_equal(r["members"], base["members"])


def test_multi_block_table_scan(monkeypatch):
    """Bucket tables above 2^18 entries (1024^3-size builds) are scanned by three kernels instead of one
    block: forced here on a small snapshot (SOGPU_SCAN1_MAX=0)."""
    s = synth.make_snapshot(64 ** 3, 80, seed=43, nmax=6000)
    base = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    monkeypatch.setenv("SOGPU_SCAN1_MAX", "0")
    r = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    assert_so_equal(r, base["rvir"], base["mvir"], base["ndelta"])
    assert np.array_equal(r["members"], base["members"])


def test_build_strategies_agree():
    """Single counting sort vs coarse-partition-first build: identical results."""
    s = synth.make_snapshot(80 ** 3, 60, seed=42, nmax=8000)
    out = []
    for mode in (0, 1, 2):
        g = api.SoGpu()
        g.set_build_mode(mode)
        g.set_particles(s.pos, s.mass)
        g.build_grid()
        g.keep_member_d2(True)
        r = g.so(s.centers, s.rgtp, 200.0)
        r["off"], r["mem"] = g.members(sorted=True)
        out.append(r)
        g.close()
    for o in out[1:]:
        assert_so_equal(o, out[0]["rvir"], out[0]["mvir"], out[0]["ndelta"])
        assert np.array_equal(out[0]["mem"], o["mem"])
    assert np.array_equal(out[0]["mem"], out[1]["mem"])
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8)
    assert np.array_equal(out[1]["mem"], ref["members"])


@pytest.mark.parametrize("n_balls", [1, 2, 4])
def test_focused_build_gives_identical_results(n_balls):
    """sogpu_build_grid_for: only the neighbourhood of the halos is sorted; results are those of the
    full grid, including halos that outgrow the planned reach (library falls back by itself)."""
    s = synth.make_snapshot(64 ** 3, 120, seed=46, nmax=8000)
    rng = np.random.default_rng(5)
    vc = (rng.random((6, 3)) - 0.5).astype(np.float32)                 # void centres: -1 / -2
    centers = np.concatenate([s.centers, vc])
    rgtp = np.concatenate([s.rgtp, np.full(3, 0.004, np.float32), np.full(3, 0.03, np.float32)])
    ref = po.Oracle(s.pos, s.mass).so(centers, rgtp, np.float32(200.0), 8)
    g = api.SoGpu()
    g.set_particles(s.pos, s.mass)
    g.build_grid_for(centers, rgtp, n_balls)
    g.keep_member_d2(True)
    r = g.so(centers, rgtp, 200.0)
    off, mem = g.members(sorted=True)
    assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
    assert np.array_equal(mem, ref["members"])
    # a ball gather on a focused grid silently rebuilds the full grid
    idx, d2, n = g.ball_gather((0.0, 0.0, 0.0), np.float32(0.01))
    oi, od = po.Oracle(s.pos, s.mass).ball((0.0, 0.0, 0.0), np.float32(0.01))
    assert n == len(oi) and np.array_equal(idx, oi)
    g.close()


def test_focused_build_threshold_never_reached():
    """Every halo outgrows any focus (-3 after the whole schedule): fallback to the full grid."""
    s = synth.make_snapshot(24 ** 3, 4, seed=36, nmax=600)
    g = api.SoGpu()
    g.set_particles(s.pos, s.mass)
    g.build_grid_for(s.centers, s.rgtp, 2)
    r = g.so(s.centers, s.rgtp, 0.5)
    assert (r["rvir"] == -3.0).all()
    g.close()


def test_records_layout_and_tiny_inputs():
    """AoS input with stride (tipsy dark records) and N smaller than the reference's nSmooth."""
    from so_b200 import tipsy
    s = synth.make_snapshot(900, 2, seed=38, nmin=20, nmax=120)
    d = tipsy.dark_from_arrays(s.pos, s.mass)
    g = api.SoGpu()
    g.set_particles_records(d)
    g.build_grid()
    g.keep_member_d2(True)
    r = g.so(s.centers, s.rgtp, 200.0)
    off, mem = g.members(sorted=True)
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8)
    assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
    assert np.array_equal(mem, ref["members"])
    g.close()


@pytest.mark.parametrize("big_endian", [False, True])
def test_streaming_ingest_of_raw_tipsy_records(big_endian):
    """sogpu_ingest_records: raw gas / dark / star records (12 / 9 / 11 floats, mass first, then x y z;
    XDR byte order for -std files) unpacked on the device give the results of the packed upload."""
    s = synth.make_snapshot(40 ** 3, 25, seed=71, nmax=3000)
    n = s.n
    ng, nd = n // 5, n // 2
    cuts = [(0, ng, 12), (ng, ng + nd, 9), (ng + nd, n, 11)]
    rng = np.random.default_rng(3)
    blocks = []
    for a, b, nf in cuts:
        rec = rng.random((b - a, nf)).astype(np.float32)          # junk in the fields the path ignores
        rec[:, 0] = s.mass
        rec[:, 1:4] = s.pos[a:b]
        blocks.append(rec.astype(">f4").view(np.float32) if big_endian else rec)
    ref = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    g = api.SoGpu()
    g.ingest_records(blocks, big_endian=big_endian, chunk=7777)
    g.build_grid()
    g.keep_member_d2(True)
    r = g.so(s.centers, s.rgtp, 200.0)
    off, mem = g.members(sorted=True)
    g.close()
    assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
    assert np.array_equal(mem, ref["members"]) and np.array_equal(off, ref["member_offset"])


def test_vcm_on_the_device_matches_oracle():
    """sogpu_vcm (_VcmParticles, kd2.c:595-609): the sequential fp32 sum over the sorted members, bit-exact
    against the oracle's replay (which is pinned to the reference's .sogtp velocities)."""
    s = synth.make_snapshot(40 ** 3, 30, seed=72, nmax=4000)
    rng = np.random.default_rng(5)
    vel = rng.normal(size=(s.n, 3)).astype(np.float32)
    rec = np.zeros((s.n, 9), np.float32)
    rec[:, 0] = s.mass
    rec[:, 1:4] = s.pos
    rec[:, 4:7] = vel
    o = po.Oracle(s.pos, s.mass)
    ref = o.so(s.centers, s.rgtp, np.float32(200.0), 8)
    t = o.tag(np.arange(1, s.h + 1), s.centers, s.gtp_mass, ref["rvir"], ref["mvir"], ref["member_offset"],
              ref["members"], vel=vel)
    g = api.SoGpu()
    g.ingest_records([rec], keep_velocities=True, chunk=50000)
    g.build_grid()
    g.keep_member_d2(True)
    r = g.so(s.centers, s.rgtp, 200.0)
    vcm = g.vcm(r["mvir"])
    g.close()
    ok = ref["rvir"] > 0
    assert ok.sum() > 20
    assert vcm[ok].tobytes() == t["vcm"][ok].tobytes()
    assert not vcm[~ok].any()


def test_pinned_host_memory_fast_paths():
    """Page-locked caller memory is DMA'd directly (xyz triplets / float4) and unpacked on the GPU."""
    import torch
    s = synth.make_snapshot(30 ** 3, 10, seed=43, nmax=1500)
    ref = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, np.float32(200.0), 8)
    pin3 = torch.from_numpy(s.pos).pin_memory()
    xyzm = np.concatenate([s.pos, np.full((s.n, 1), s.mass, np.float32)], axis=1)
    pin4 = torch.from_numpy(xyzm).pin_memory()
    for which in ("xyz", "xyzm"):
        g = api.SoGpu()
        if which == "xyz":
            g.set_particles(pin3.numpy(), s.mass)
        else:
            a = pin4.numpy()
            g.set_particles(a[:, :3], a[:, 3])
        g.build_grid()
        g.keep_member_d2(True)
        r = g.so(s.centers, s.rgtp, 200.0)
        off, mem = g.members(sorted=True)
        assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
        assert np.array_equal(mem, ref["members"])
        g.close()


def test_ball_gather_matches_oracle_ball():
    """smBallGather + qsort replacement (smooth2.c:58-114, kd2.c:781)."""
    s = synth.make_snapshot(40 ** 3, 10, seed=39, nmax=4000)
    o = po.Oracle(s.pos, s.mass)
    g = api.SoGpu()
    g.set_particles(s.pos, s.mass)
    g.build_grid()
    for i in range(s.h):
        for f in (0.5, 1.2, 3.0):
            b = np.float32(s.rgtp[i] * f)
            b2 = np.float32(b * b)
            idx, d2, n = g.ball_gather(s.centers[i], b2)
            oi, od = o.ball(s.centers[i], b2)
            assert n == len(oi) and np.array_equal(idx, oi) and d2.tobytes() == od.tobytes()
    # empty ball, and a ball wider than half the box
    idx, d2, n = g.ball_gather((0.1234, -0.3, 0.2), np.float32(1e-12))
    assert n == len(o.ball((0.1234, -0.3, 0.2), np.float32(1e-12))[0])
    g.close()


@pytest.mark.parametrize("nmem", [8, 3])
def test_vcirc_and_mass_profile_match_oracle(nmem):
    """sogpu_vcirc (kdVcirc + kdMassProfile, kd2.c:498-586): bit-exact against the oracle's literal walk of
    the sorted 2*Rvir lists, including a group on the periodic boundary and Omega0 = 0.3 (a particle mass
    that is not a power of two, so the sequential fp32 mass sum matters)."""
    for seed, omega0 in ((61, 1.0), (62, 0.3)):
        s = synth.make_snapshot(48 ** 3, 40, seed=seed, nmax=6000, omega0=omega0)
        centers = s.centers.copy()
        centers[0] = (0.4995, -0.4995, 0.1)
        thr = np.float32(np.float32(200.0) * np.float32(omega0))
        o = po.Oracle(s.pos, s.mass)
        ref = o.so(centers, s.rgtp, thr, nmem)
        ok = ref["rvir"] > 0
        assert ok.sum() > 20
        want = o.vcirc(centers[ok], ref["rvir"][ok], ref["mvir"][ok], 1.0, nmem)
        g = api.SoGpu()
        g.set_particles(s.pos, s.mass)
        g.build_grid()
        got = g.vcirc(centers[ok], ref["rvir"][ok], ref["mvir"][ok], 1.0, nmem)
        for k in ("vcirc", "rmass", "rmax", "vmax", "profile"):
            assert got[k].tobytes() == want[k].tobytes(), k
        # the sorted 2 Rvir lists stay available
        off, mem, d2 = g.members(want_d2=True, sorted=True)
        oi, od = o.ball(centers[ok][3], np.float32(np.float32(2.0 * ref["rvir"][ok][3]) ** 2))
        assert np.array_equal(mem[off[3]:off[4]], oi) and d2[off[3]:off[4]].tobytes() == od.tobytes()
        g.close()


def test_vcirc_species_with_mixed_masses_matches_oracle():
    """sogpu_vcirc_species: kdVcirc + the per-species kdMassProfile sums (kd2.c:458-496, 498-586) for a gas + dark +
    star snapshot with three different masses — the cumulative mass is the reference's sequential fp32 sum in
    sorted order, evaluated on the device; bit-exact against the oracle's literal walk (r^2 ties between different
    masses are out of contract and do not occur in this seeded input)."""
    s = synth.make_snapshot(40 ** 3, 30, seed=64, nmax=5000)
    rng = np.random.default_rng(64)
    ptype = rng.choice(np.array([1, 2, 4], np.uint8), size=s.n, p=[0.6, 0.3, 0.1])      # dark / gas / star bits
    mass = np.where(ptype == 1, s.mass, np.where(ptype == 2, s.mass * np.float32(0.37), s.mass * np.float32(0.11))).astype(np.float32)
    mark = (rng.random(s.n) < 0.05)
    ptype = (ptype | (mark.astype(np.uint8) << 3)).astype(np.uint8)
    thr = np.float32(120.0)
    o = po.Oracle(s.pos, mass)
    ref = o.so(s.centers, s.rgtp, thr, 8)
    ok = ref["rvir"] > 0
    assert ok.sum() > 15
    g = api.SoGpu()
    g.set_particles(s.pos, mass)
    g.build_grid()
    got = g.vcirc_species(s.centers[ok], ref["rvir"][ok], ref["mvir"][ok], ptype, masks=(1, 2, 4, 8))
    g.close()
    want = o.vcirc(s.centers[ok], ref["rvir"][ok], ref["mvir"][ok], 1.0, 8)
    for k in ("vcirc", "rmass", "rmax", "vmax"):
        assert got[k].tobytes() == want[k].tobytes(), k
    for m, bit in enumerate((1, 2, 4, 8)):
        w = o.vcirc(s.centers[ok], ref["rvir"][ok], ref["mvir"][ok], 1.0, 8, ptype_of=ptype, ptype_mask=bit)
        assert got["profiles"][m].tobytes() == w["profile"].tobytes(), bit


def test_vcirc_refuses_mixed_masses():
    s = synth.make_snapshot(20 ** 3, 4, seed=63, nmax=500)
    mass = np.full(s.n, s.mass, np.float32)
    mass[::7] *= np.float32(3.0)
    g = api.SoGpu()
    g.set_particles(s.pos, mass)
    g.build_grid()
    with pytest.raises(api.SoGpuError):
        g.vcirc(s.centers, np.full(s.h, 0.02, np.float32), np.full(s.h, 1e-4, np.float32))
    g.close()


@pytest.mark.parametrize("name", ["conflict", "basic"])
def test_device_tagging_of_conflict_free_groups(name):
    """sogpu_tag_members against the sequential kdTagParticles replay of the oracle (kd2.c:663-720): every
    particle the device tags carries the tag the full replay ends with, every group the replay subsumes,
    slurps or lets ignore particles is flagged, and flagged groups leave their particles untagged."""
    s, g = load_golden(name)
    h = len(g["rgtp"])
    ids = np.arange(1, h + 1, dtype=np.int32)
    o = po.Oracle(s.pos, s.mass)
    ref = o.so(g["centers"], g["rgtp"], g["thr"], int(g["n_members"]))
    t = o.tag(ids, g["centers"], g["gtp_mass"], ref["rvir"].copy(), ref["mvir"].copy(), ref["member_offset"],
              ref["members"])
    gpu = api.SoGpu()
    gpu.set_particles(s.pos, s.mass)
    gpu.build_grid()
    gpu.so(g["centers"], g["rgtp"], g["thr"], int(g["n_members"]))
    dirty, igrp = gpu.tag_members(ids, s.n)
    gpu.close()
    off, mem = ref["member_offset"], ref["members"]
    owners = np.zeros(s.n, np.int32)
    np.add.at(owners, mem, 1)
    for i in range(h):
        seg = mem[off[i]:off[i + 1]]
        shares = bool((owners[seg] > 1).any())
        assert dirty[i] == shares, i
        if not shares:
            assert (igrp[seg] == ids[i]).all() and (t["igrp"][seg] == ids[i]).all()
        else:
            assert not (igrp[seg] == ids[i]).any()
    tagged = igrp != 0
    assert np.array_equal(igrp[tagged], t["igrp"][tagged])
    changed = t["rvir"] != ref["rvir"]                     # subsumed or slurped by the replay
    assert dirty[changed].all()
    if name == "conflict":
        assert dirty.any() and changed.any()
    else:
        assert not dirty.any() and np.array_equal(igrp, t["igrp"])


def _replay_case(s, centers, rgtp, gtp_mass, thr, nmem=8):
    h = len(rgtp)
    ids = np.arange(1, h + 1, dtype=np.int32)
    o = po.Oracle(s.pos, s.mass)
    ref = o.so(centers, rgtp, thr, nmem)
    t = o.tag(ids, centers, gtp_mass, ref["rvir"].copy(), ref["mvir"].copy(), ref["member_offset"], ref["members"])
    gpu = api.SoGpu()
    gpu.set_particles(s.pos, s.mass)
    gpu.build_grid()
    gpu.keep_member_d2(True)
    r = gpu.so(centers, rgtp, thr, nmem)
    gpu.members(sorted=True)
    dirty, _ = gpu.tag_members(ids, s.n)
    order = [g for g in (po.indexx(gtp_mass) - 1) if dirty[g] and r["rvir"][g] > 0]     # kdSortMass order (kd2.c:843-861)
    out = gpu.tag_replay(order, ids, centers, r["rvir"], r["mvir"], s.n)
    gpu.close()
    assert np.array_equal(out["igrp"], t["igrp"]), "PINIT.iGrp differs"
    assert np.array_equal(out["nsub"], t["nsub"]) and np.array_equal(out["nign"], t["nign"])
    assert out["rvir"].tobytes() == t["rvir"].tobytes() and out["mvir"].tobytes() == t["mvir"].tobytes()
    assert (out["groups_removed"], out["groups_slurped"]) == (t["groups_removed"], t["groups_slurped"])
    return t, dirty


def test_ordered_conflict_replay_on_the_device_golden():
    """sogpu_tag_replay (kdTagParticles incl. subsume / slurp / ignore and kdZeroGroup, kd2.c:617-720) on the
    `conflict` golden: PINIT.iGrp / nSubsumed / nIgnored, the -10*index / -Mvir marks and both counters equal the
    oracle's sequential replay, which is pinned to the reference's .sogrp."""
    s, g = load_golden("conflict")
    t, dirty = _replay_case(s, g["centers"], g["rgtp"], g["gtp_mass"], g["thr"], int(g["n_members"]))
    assert np.array_equal(t["igrp"], g["igrp"])              # the reference binary's own .sogrp
    assert t["groups_removed"] == int(g["groups_removed"]) and dirty.any()


def test_ordered_conflict_replay_with_a_thousand_conflicts():
    """A catalog built to collide: 1200 of 2500 halos sit next to a bigger neighbour (subsume, slurp and ignore
    all occur, chains of events inside one group's walk included)."""
    s = synth.make_snapshot(96 ** 3, 2500, seed=83, nmax=3000, overlap_pairs=1200)
    t, dirty = _replay_case(s, s.centers, s.rgtp, s.gtp_mass, np.float32(200.0))
    assert dirty.sum() >= 1000 and t["groups_removed"] > 100 and (t["nign"] > 0).sum() > 1000


def test_unequal_masses_general_path():
    """Mixed particle masses: the enclosed mass is the SEQUENTIAL fp32 sum in sorted order, so the
    library switches to the full-sort path; results must still be bit-exact."""
    s = synth.make_snapshot(40 ** 3, 24, seed=40, nmax=6000)
    rng = np.random.default_rng(11)
    m = np.full(s.n, s.mass, np.float32)
    sel = rng.random(s.n)
    m[sel < 0.3] *= np.float32(0.37)          # "gas"
    m[sel > 0.9] *= np.float32(2.9)           # "stars"
    vc = (rng.random((6, 3)) - 0.5).astype(np.float32)          # -1 / -2 cases
    centers = np.concatenate([s.centers, vc])
    rgtp = np.concatenate([s.rgtp, np.full(3, 0.01, np.float32), np.full(3, 0.06, np.float32)])
    g = api.SoGpu()
    g.set_particles(s.pos, m)
    g.build_grid()
    g.keep_member_d2(True)
    r = g.so(centers, rgtp, 150.0)
    off, mem = g.members(sorted=True)
    assert g.stats()["equal_mass"] == 0
    ref = po.Oracle(s.pos, m).so(centers, rgtp, np.float32(150.0), 8)
    assert_so_equal(r, ref["rvir"], ref["mvir"], ref["ndelta"])
    assert np.array_equal(off, ref["member_offset"]) and np.array_equal(mem, ref["members"])
    assert (ref["rvir"] > 0).sum() >= 20 and (ref["rvir"] < 0).sum() >= 3
    # never-reached threshold on the general path: -3 for everything
    r3 = g.so(s.centers[:3], s.rgtp[:3], 0.01)
    assert (r3["rvir"] == -3.0).all()
    g.close()


def test_thousands_of_particles_at_identical_radius():
    """More coincident particles than any histogram level can split: those halos are re-run through
    the general full-sort path instead of failing."""
    s = synth.make_snapshot(32 ** 3, 6, seed=44, nmax=2500)
    pos = s.pos.copy()
    rng = np.random.default_rng(4)
    bg = rng.choice(s.n, 12000, replace=False)
    pos[bg[:6000]] = s.centers[0]                                    # r^2 = 0 for 6000 particles of halo 0
    off = np.array([0.3, -0.2, 0.1], np.float32) * np.float32(s.rgtp[1])
    pos[bg[6000:]] = s.centers[1] + off                              # one identical r^2 > 0 around halo 1
    r, ref = check_against_oracle(pos, s.mass, s.centers, s.rgtp, 200.0)
    assert ref["ndelta"][0] > 6000 or ref["rvir"][0] < 0
    assert r["stats"]["equal_mass"] == 1


def test_member_buffer_grows_when_halos_overlap_heavily():
    """Sum of N_Delta far above N (the same big halo listed hundreds of times)."""
    s = synth.make_snapshot(40 ** 3, 3, seed=45, sizes=[6000, 3000, 500], nmax=1e4)
    reps = 300
    centers = np.repeat(s.centers[:1], reps, axis=0)
    rgtp = np.repeat(s.rgtp[:1], reps)
    r, ref = check_against_oracle(s.pos, s.mass, centers, rgtp, 200.0)
    assert int(ref["member_offset"][-1]) > (1 << 20) > s.n


def test_bad_arguments_return_errors():
    g = api.SoGpu()
    with pytest.raises(api.SoGpuError):
        g.build_grid()                       # no particles yet
    s = synth.make_snapshot(3000, 2, seed=41, nmax=200)
    with pytest.raises(api.SoGpuError):
        g.set_particles(s.pos, s.mass, period=(0.0, 1.0, 1.0))
    g.set_particles(s.pos, s.mass)
    with pytest.raises(api.SoGpuError):
        g.so(s.centers, s.rgtp, 200.0)       # grid not built
    g.build_grid()
    with pytest.raises(api.SoGpuError):
        g.so(s.centers, s.rgtp, 200.0, n_members=1)
    g.close()


def test_config1_full_size_properties():
    """BASELINE.json configs[1] (256^3, 10 000 halos): too big for the oracle to be the only
    check, so: size-independent properties + the oracle on a random subsample of the halos."""
    s = synth.config(1)
    r = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    off, mem, d2 = r["member_offset"], r["members"], r["members_d2"]
    ok = r["ndelta"] > 0
    assert ok.sum() > 0.95 * s.h
    # counts, sortedness, uniqueness, M = sequential sum of N equal masses, R from M
    assert np.array_equal(np.diff(off), np.where(ok, r["ndelta"], 0))
    for i in np.nonzero(ok)[0][:2000]:
        seg = d2[off[i]:off[i + 1]]
        assert np.all(np.diff(seg) >= 0)
        assert len(np.unique(mem[off[i]:off[i + 1]])) == r["ndelta"][i]
    k = r["ndelta"][ok].astype(np.int64)
    s_next = api.mass_prefix(s.mass, k + 1)
    assert np.array_equal(r["mvir"][ok], (s_next - s.mass).astype(np.float32))
    assert np.array_equal(r["rvir"][ok], np.array([api.rdelta(m, 200.0) for m in r["mvir"][ok]], np.float32))
    # idempotence: a second context gives identical bits
    r2 = run_gpu(s.pos, s.mass, s.centers, s.rgtp, 200.0)
    assert_so_equal(r2, r["rvir"], r["mvir"], r["ndelta"])
    assert np.array_equal(r2["members"], mem)
    # oracle on a subsample
    pick = np.random.default_rng(0).choice(s.h, 300, replace=False)
    ref = po.Oracle(s.pos, s.mass).so(s.centers[pick], s.rgtp[pick], np.float32(200.0), 8)
    assert np.array_equal(r["ndelta"][pick], ref["ndelta"])
    assert r["mvir"][pick].tobytes() == ref["mvir"].tobytes()
    assert r["rvir"][pick].tobytes() == ref["rvir"].tobytes()
    for n, i in enumerate(pick):
        assert np.array_equal(mem[off[i]:off[i + 1]], ref["members"][ref["member_offset"][n]:ref["member_offset"][n + 1]])


@pytest.mark.parametrize("n_ranks,n_balls", [(1, 4), (2, 1), (4, 2), (8, 4)])
def test_domain_run_matches_single_grid(n_ranks, n_balls):
    """Domain run: slices of the snapshot, focus masks, routing of {x,y,z,global index}
    records, one grid per rank over what it received — all ranks simulated on this one device.  Results
    (incl. -1/-2/-3 codes, halos on the periodic boundary, halos whose balls outgrow the first masks and
    are re-run with larger ones) are those of the single full grid, member indices global."""
    from so_b200 import parallel
    s = synth.make_snapshot(64 ** 3, 150, seed=81, nmax=6000)
    rng = np.random.default_rng(11)
    vc = (rng.random((10, 3)) - 0.5).astype(np.float32)
    centers = np.concatenate([s.centers, vc, np.array([[0.4999, 0.4999, -0.4999]], np.float32)])
    rgtp = np.concatenate([s.rgtp, np.full(5, 0.004, np.float32), np.full(5, 0.03, np.float32), [np.float32(0.01)]])
    ref = run_gpu(s.pos, s.mass, centers, rgtp, 200.0)
    out = parallel.VirtualDomainRun(n_ranks, n_balls).run(s.pos, s.mass, centers, rgtp, np.float32(200.0))
    assert_so_equal(out, ref["rvir"], ref["mvir"], ref["ndelta"])
    for i in range(len(rgtp)):
        a = ref["members"][ref["member_offset"][i]:ref["member_offset"][i + 1]]
        b = out["members"][i] if out["members"][i] is not None else np.zeros(0, np.int32)
        assert np.array_equal(np.sort(a), np.sort(b)), i
    if n_ranks > 1:
        assert out["sent"] < n_ranks * s.n        # far from "everything to everyone"


def _domain_step_inputs():
    s = synth.make_snapshot(64 ** 3, 150, seed=81, nmax=6000)
    rng = np.random.default_rng(11)
    vc = (rng.random((10, 3)) - 0.5).astype(np.float32)
    centers = np.concatenate([s.centers, vc, np.array([[0.4999, 0.4999, -0.4999]], np.float32)])
    rgtp = np.concatenate([s.rgtp, np.full(5, 0.004, np.float32), np.full(5, 0.03, np.float32), [np.float32(0.01)]])
    return s, centers, rgtp


def _check_domain_step(out, ref, n_ranks):
    assert_so_equal(out, ref["rvir"], ref["mvir"], ref["ndelta"])
    for i in range(len(ref["ndelta"])):
        a = ref["members"][ref["member_offset"][i]:ref["member_offset"][i + 1]]
        b = out["members"][i] if out["members"][i] is not None else np.zeros(0, np.int32)
        assert np.array_equal(a, b), i            # same (r^2, index) order, global indices
    if n_ranks > 1:
        assert len(np.unique(out["owner"])) == n_ranks          # every rank owns some halos


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranks,n_balls,direct", [(1, 4, "1"), (2, 1, "1"), (3, 2, "1"), (3, 2, "0"), (8, 4, "1"), (8, 4, "0")])
def test_domain_step_matches_single_grid(n_ranks, n_balls, direct, monkeypatch):
    """The stream-ordered domain step (device-side ownership, plain / listed cells and the destination table of the
    halos at ownership boundaries, one-pass routing into the hit list, k_route_split into the receivers' buffers
    — direct = "0": into staging runs shipped by k_push_copy —, receiver-side reservations, grid build with the
    particle count read on the device) with all ranks simulated on one device: results of the single full grid bit
    for bit, incl. error codes, the periodic boundary and halos whose balls outgrow the first masks."""
    import torch
    monkeypatch.setenv("SOGPU_DIRECT_PUSH", direct)
    from so_b200 import parallel
    s, centers, rgtp = _domain_step_inputs()
    ref = run_gpu(s.pos, s.mass, centers, rgtp, 200.0)
    dev = torch.device("cuda", 0)
    full = torch.empty((s.n, 4), dtype=torch.float32, device=dev)
    full[:, :3] = torch.from_numpy(s.pos).to(dev)
    full[:, 3] = float(s.mass)
    bounds = parallel.slice_bounds(s.n, n_ranks)
    slices = [full[a:b].contiguous() for a, b in bounds]
    torch.cuda.synchronize()
    run = parallel.VirtualDomainStep(n_ranks, s.n, s.mass)
    try:
        out = run.run(slices, centers, rgtp, np.float32(200.0), 8, n_balls)
    finally:
        run.close()
    _check_domain_step(out, ref, n_ranks)
    assert sum(out["n_recv"][0]) < n_ranks * s.n
    # the ownership every rank derived on its device == the numpy restatement the CPU tests use
    assert np.array_equal(out["owner"], parallel.owner_numpy(centers, rgtp, s.n, n_ranks))


@pytest.mark.gpu
def test_domain_step_reports_overflow_instead_of_writing_past_the_buffers(monkeypatch):
    import torch
    from so_b200 import parallel
    s, centers, rgtp = _domain_step_inputs()
    dev = torch.device("cuda", 0)
    full = torch.empty((s.n, 4), dtype=torch.float32, device=dev)
    full[:, :3] = torch.from_numpy(s.pos).to(dev)
    full[:, 3] = float(s.mass)
    torch.cuda.synchronize()
    for recv_cap, stage_cap, what, direct in [(2000, 1 << 16, "receive", "1"), (2000, 1 << 16, "receive", "0"),
                                              (1 << 18, 500, "staging", "0")]:
        # SOGPU_DIRECT_PUSH=0: runs are staged locally and shipped by k_push_copy (default: k_route_split stores
        # into the receivers' buffers itself, so only the receive buffers can overflow)
        monkeypatch.setenv("SOGPU_DIRECT_PUSH", direct)
        run = parallel.VirtualDomainStep(2, s.n, s.mass, recv_cap=recv_cap, stage_cap=stage_cap)
        try:
            with pytest.raises(RuntimeError, match=what):
                run.run([full[: s.n // 2].contiguous(), full[s.n // 2:].contiguous()], centers, rgtp, np.float32(200.0))
        finally:
            run.close()


@pytest.mark.gpu
def test_domain_step_across_devices():
    """The same step with one rank per visible GPU inside this process: pushes travel over NVLink peer memory
    and the ranks meet at the flag barrier.  Runs when more than one device is visible."""
    import torch
    from so_b200 import parallel
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("one device visible")
    R = min(nd, 8)
    s, centers, rgtp = _domain_step_inputs()
    ref = run_gpu(s.pos, s.mass, centers, rgtp, 200.0)
    bounds = parallel.slice_bounds(s.n, R)
    slices = []
    for r, (a, b) in enumerate(bounds):
        dev = torch.device("cuda", r)
        t = torch.empty((b - a, 4), dtype=torch.float32, device=dev)
        t[:, :3] = torch.from_numpy(s.pos[a:b]).to(dev)
        t[:, 3] = float(s.mass)
        slices.append(t)
    run = parallel.VirtualDomainStep(R, s.n, s.mass, devices=list(range(R)))
    try:
        out = run.run(slices, centers, rgtp, np.float32(200.0), 8, 4)
    finally:
        run.close()
    _check_domain_step(out, ref, R)
