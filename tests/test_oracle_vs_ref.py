"""The oracle against the reference binary itself, on fresh seeded inputs (runs wherever
oracle/_ref/so_ref_inst exists: it is built from /root/reference by oracle/Makefile and
travels to the GPU box as a prebuilt file)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from so_b200 import synth, tipsy

pytestmark = pytest.mark.skipif(not po.ref_available("so_ref_inst"),
                                reason="reference binary not built (oracle/_ref/so_ref_inst)")


@pytest.mark.parametrize("seed,omega0,delta", [(101, 1.0, 200.0), (102, 0.3, 200.0), (103, 1.0, 500.0)])
def test_oracle_bit_exact_against_reference_binary(tmp_path, seed, omega0, delta):
    s = synth.make_snapshot(40 ** 3, 25, seed=seed, nmax=6000, omega0=omega0)
    snap, gtp, out, inst = (str(tmp_path / n) for n in ("s.tipsy", "h.gtp", "o", "o.inst"))
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    po.run_so_ref(snap, gtp, out, delta=delta, extra=["-gtp", "-O", repr(omega0)], inst=True, inst_file=inst)
    rec = po.read_inst_file(inst)
    _, star = tipsy.read_gtp(out + ".sogtp")
    thr = np.float32(np.float32(delta) * np.float32(omega0))
    res = po.Oracle(s.pos, s.mass).so(s.centers, s.rgtp, thr, 8)
    for i in range(s.h):
        if res["rvir"][i] > 0 and star["eps"][i] > -10:
            assert res["rvir"][i].tobytes() == star["eps"][i].tobytes()
            assert res["mvir"][i].tobytes() == star["mass"][i].tobytes()
        if res["ndelta"][i] > 0:
            j, order, _ = rec[i + 1]
            assert j == res["ndelta"][i]
            a = res["members"][res["member_offset"][i]:res["member_offset"][i + 1]]
            assert np.array_equal(np.sort(a), np.sort(order))
        else:
            assert (i + 1) not in rec and star["eps"][i] == res["rvir"][i]


def test_timed_driver_agrees_with_reference_main(tmp_path):
    """so_ref_timed (pristine reference objects, our timing main) = so_ref's .sogtp."""
    if not po.ref_available("so_ref_timed"):
        pytest.skip("so_ref_timed not built")
    s = synth.make_snapshot(32 ** 3, 12, seed=104, nmax=2000)
    snap, gtp = str(tmp_path / "s.tipsy"), str(tmp_path / "h.gtp")
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    po.run_so_ref(snap, gtp, str(tmp_path / "a"), delta=200.0, extra=["-gtp"])
    t = po.run_so_ref_timed(snap, gtp, 200.0, 8, 1.0, str(tmp_path / "b"))
    assert t["n"] == s.n and t["h"] == s.h and t["t_build"] > 0 and t["t_so"] > 0
    a = open(str(tmp_path / "a.sogtp"), "rb").read()
    b = open(str(tmp_path / "b.sogtp"), "rb").read()
    # bytes 28..31 are the uninitialised padding of `struct dump` (kd2.c:1272,1297)
    assert a[:28] == b[:28] and a[32:] == b[32:]


@pytest.mark.skipif(not po.ref_available("so_ref"), reason="reference binary not built")
def test_oracle_vcirc_and_species_profiles_against_reference_binary(tmp_path):
    """kdVcirc + kdMassProfile restated in the oracle (so_oracle_vcirc, with the per-species mask) against a
    live run of the reference on a gas + dark + star snapshot with three different particle masses:
    .sovcirc columns and the .sodark/.sogas/.sostar rows, which the reference prints with %g."""
    s = synth.make_snapshot(36 ** 3, 16, seed=79, nmax=3000)
    ng, ns = s.n // 4, s.n // 10
    nd = s.n - ng - ns
    gas = np.zeros(ng, tipsy.GAS_DT)
    gas["mass"], gas["pos"] = s.mass * np.float32(0.4), s.pos[:ng]
    dark = tipsy.dark_from_arrays(s.pos[ng:ng + nd], s.mass)
    star = np.zeros(ns, tipsy.STAR_DT)
    star["mass"], star["pos"] = s.mass * np.float32(0.13), s.pos[ng + nd:]
    snap, gtp, out = (str(tmp_path / n) for n in ("s.tipsy", "h.gtp", "o"))
    tipsy.write_tipsy(snap, s.time, gas=gas, dark=dark, star=star)
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    po.run_so_ref(snap, gtp, out, delta=120.0, extra=["-gtp", "-all"])
    mass = np.concatenate([gas["mass"], dark["mass"], star["mass"]]).astype(np.float32)
    ptype = np.concatenate([np.full(ng, 2), np.full(nd, 1), np.full(ns, 4)]).astype(np.uint8)   # GAS 2, DARK 1, STAR 4
    o = po.Oracle(s.pos, mass)
    res = o.so(s.centers, s.rgtp, np.float32(120.0), 8)
    _, rows = tipsy.parse_sovcirc(out + ".sovcirc")
    rows = np.array(rows)
    ok = rows[:, 2] > 0
    assert ok.sum() >= 12
    # Mvir / Rvir of the general (mixed-mass) solver, then the Vcirc block
    np.testing.assert_allclose(res["mvir"][ok], rows[ok, 1], rtol=6e-6)
    np.testing.assert_allclose(res["rvir"][ok], rows[ok, 2], rtol=6e-6)
    v = o.vcirc(s.centers, np.where(ok, res["rvir"], 0).astype(np.float32), res["mvir"], 1.0, 8)
    ours = np.concatenate([v["rmass"], v["rmax"][:, None], v["vmax"][:, None], v["vcirc"]], axis=1)
    np.testing.assert_allclose(ours[ok], rows[ok, 3:15], rtol=2e-5)      # %g; ties between unequal masses: SURVEY H3
    for ext, bit in ((".sodark", 1), (".sogas", 2), (".sostar", 4)):
        _, prow = tipsy.parse_sovcirc(out + ext)
        prow = np.array(prow)
        p = o.vcirc(s.centers, np.where(ok, res["rvir"], 0).astype(np.float32), res["mvir"], 1.0, 8,
                    ptype_of=ptype, ptype_mask=bit)["profile"]
        np.testing.assert_allclose(p[ok], prow[ok, 1:17], rtol=2e-5, err_msg=ext)
        assert (p[ok, -1] > 0).all()


@pytest.mark.skipif(not po.ref_available("so_ref"), reason="reference binary not built")
@pytest.mark.parametrize("omega0,flat,extra", [(0.3, True, ["-L"]), (0.3, False, []), (1.0, False, []), (0.25, True, ["-L", "-z", "1.5"])])
def test_virial_threshold_restatement_matches_the_reference_header(tmp_path, omega0, flat, extra):
    """po.virial_threshold restates so.c:57-86,470-481; the reference prints the value it used in the .sovcirc
    header ('# fThreshold = %g  (VIRIAL DENSITY)'): same 6 significant digits for both cosmology branches."""
    from so_b200 import synth, tipsy
    s = synth.make_snapshot(20 ** 3, 3, seed=5, nmax=300, omega0=omega0, z=0.5)
    snap, gtp, out = str(tmp_path / "s.tipsy"), str(tmp_path / "h.gtp"), str(tmp_path / "ref")
    tipsy.write_tipsy(snap, s.time, dark=tipsy.dark_from_arrays(s.pos, s.mass))
    tipsy.write_gtp(gtp, s.time, s.centers, s.rgtp, s.gtp_mass)
    po.run_so_ref(snap, gtp, out, delta=None, extra=["-O", repr(omega0), "-s", "64"] + extra)
    hdr, _ = tipsy.parse_sovcirc(out + ".sovcirc")
    line = [l for l in hdr if "fThreshold" in l][0]
    assert "VIRIAL DENSITY" in line
    z = 1.5 if "-z" in extra else None
    mine = po.virial_threshold(omega0, flat, time=np.float32(s.time), z=z)
    assert line.split("=")[1].split()[0] == "%g" % mine, (line, mine)
