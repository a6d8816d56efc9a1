"""CPU-side checks of the product: C-ABI surface, exact-arithmetic helpers, file formats.
No compute call is made (there is no GPU here and no CPU fallback to call)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as po
from so_b200 import api, synth, tipsy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sogpu.h")).read()
    declared = set(re.findall(r"\b(sogpu_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("sogpu_t")
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    L = api.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = api.lib().sogpu_create(ctypes.byref(h), -1)
    assert rc != 0 and not h.value
    assert len(api.lib().sogpu_last_error()) > 0
    with pytest.raises(api.SoGpuError):
        api.SoGpu()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "so_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "so_oracle" not in txt and "liboracle" not in txt, f


@pytest.mark.parametrize("m", [2.0 ** -21, 0.3 / 2 ** 21, 0.3 / 2 ** 27, 1.0 / 3e6, 1e-7, 0.1, 1.0, 3.0,
                               1.5e-9, 7.7e-5, 1.0 / 1024 ** 3])
def test_mass_prefix_table_equals_literal_fp32_loop(m):
    """S[k] = fl(S[k-1] + m) (kd2.c:787,807) vs the compressed table the kernels evaluate."""
    kmax = 200_000
    s = np.float32(0.0)
    mm = np.float32(m)
    lit = np.zeros(kmax + 1, np.float32)
    for k in range(1, kmax + 1):
        s = np.float32(s + mm)
        lit[k] = s
    got = api.mass_prefix(m, np.arange(kmax + 1))
    assert got.tobytes() == lit.tobytes()


def test_mass_prefix_far_range_is_consistent():
    """Spot-check very large k (beyond what a literal Python loop can reach) with a chunked
    numpy recurrence started from a table value."""
    m = np.float32(1.0 / 1024 ** 3)
    k0 = 900_000_000
    base = api.mass_prefix(m, [k0])[0]
    s = base
    for _ in range(1000):
        s = np.float32(s + m)
    assert api.mass_prefix(m, [k0 + 1000])[0] == s


def test_ball_schedule_and_rdelta_match_oracle():
    for rgtp in (1e-4, 0.0066, 0.05, 0.2, 0.44):
        assert np.array_equal(api.ball_schedule(rgtp), po.schedule(rgtp))
    per = (2.0, 1.0, 0.5)
    assert np.array_equal(api.ball_schedule(0.01, per), po.schedule(0.01, per))
    rng = np.random.default_rng(3)
    for _ in range(1000):
        mv = np.float32(10 ** rng.uniform(-7, 0))
        thr = np.float32(10 ** rng.uniform(0, 3))
        assert np.float32(api.rdelta(mv, thr)) == np.float32(po.rdelta(mv, thr))


def test_tipsy_roundtrip_native_and_standard(tmp_path):
    s = synth.make_snapshot(4096, 3, seed=2, nmax=200)
    d = tipsy.dark_from_arrays(s.pos, s.mass)
    for std in (False, True):
        p = str(tmp_path / ("snap%d" % std))
        tipsy.write_tipsy(p, s.time, dark=d, standard=std)
        hdr, gas, dark, star = tipsy.read_tipsy(p, standard=std)
        assert hdr["nbodies"] == s.n and hdr["ndark"] == s.n and len(gas) == 0 and len(star) == 0
        assert dark.tobytes() == d.tobytes()
        assert os.path.getsize(p) == 32 + 36 * s.n
        q = str(tmp_path / ("gtp%d" % std))
        tipsy.write_gtp(q, s.time, s.centers, s.rgtp, s.gtp_mass, standard=std)
        _, st = tipsy.read_gtp(q, standard=std)
        assert np.array_equal(st["pos"], s.centers) and np.array_equal(st["eps"], s.rgtp)


def test_generator_is_deterministic_and_in_box():
    a = synth.make_snapshot(20000, 10, seed=9, nmax=500)
    b = synth.make_snapshot(20000, 10, seed=9, nmax=500)
    assert a.pos.tobytes() == b.pos.tobytes() and a.centers.tobytes() == b.centers.tobytes()
    assert a.pos.min() >= -0.5 and a.pos.max() < 0.5
    assert len(np.unique(a.gtp_mass)) == a.h
    assert a.mass == np.float32(1.0 / 20000)
