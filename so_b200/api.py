"""ctypes binding of the C-ABI library (include/sogpu.h -> so_b200/libsogpu.so).

This is the call a Python user makes; tests and bench.py go through it, i.e. through the
C-ABI.  There is NO fallback: if the CUDA library is missing or no B200 is usable the calls
raise SoGpuError (the oracle under oracle/ is test infrastructure and is never imported here).

Names mirror the reference's hot-path API (/root/reference/kd2.h:258-276):
    KD.kdBuildTree()            kd2.c:1096-1185 -> sogpu_build_grid
    KD.kdSO(rhovir, ...)        kd2.c:864-895   -> sogpu_so  (+ host replay of kdTagParticles)
    KD.smBallGather(ball2, ri)  smooth2.c:58-114 -> sogpu_ball_gather
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOGPU_LIB") or os.path.join(HERE, "libsogpu.so")   # SOGPU_LIB: A/B builds of the library

SYMBOLS = [
    "sogpu_create", "sogpu_destroy", "sogpu_last_error", "sogpu_set_stream",
    "sogpu_set_cell_occupancy", "sogpu_set_particles_host", "sogpu_set_particles_device",
    "sogpu_build_grid", "sogpu_so", "sogpu_so_device", "sogpu_members", "sogpu_ball_gather",
    "sogpu_get_stats", "sogpu_mass_prefix", "sogpu_ball_schedule", "sogpu_rdelta",
    "sogpu_finish_host", "sogpu_keep_member_d2", "sogpu_profile_enable", "sogpu_profile_kernels",
    "sogpu_profile_name", "sogpu_profile_read", "sogpu_upload_particles",
    "sogpu_set_build_mode", "sogpu_ball_gather_batch",
    "sogpu_profile_bytes", "sogpu_build_grid_for", "sogpu_build_grid_for_device", "sogpu_set_first_ball", "sogpu_set_tma_staging", "sogpu_debug_timeline", "sogpu_vcirc", "sogpu_tag_members", "sogpu_vcm", "sogpu_ingest_keep_velocities", "sogpu_host_alloc", "sogpu_host_free", "sogpu_ingest_begin",
    "sogpu_ingest_records", "sogpu_ingest_end", "sogpu_domain_mask_words", "sogpu_domain_mask",
    "sogpu_domain_route_count", "sogpu_domain_route_scatter", "sogpu_set_particles_device_indexed",
    "sogpu_peer_alloc", "sogpu_peer_open", "sogpu_peer_close", "sogpu_peer_free",
    "sogpu_domain_open", "sogpu_domain_connect", "sogpu_enable_peer_access", "sogpu_domain_begin",
    "sogpu_domain_route", "sogpu_domain_route_host", "sogpu_domain_push", "sogpu_domain_solve",
    "sogpu_domain_result", "sogpu_domain_close", "sogpu_domain_pointers", "sogpu_vcirc_species", "sogpu_tag_replay",
    "sogpu_particles_device", "sogpu_copy", "sogpu_set_members",
]


class SoGpuError(RuntimeError):
    pass


class DomainCfg(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("n_total", C.c_int64), ("mass", C.c_float),
                ("period", C.c_float * 3), ("center", C.c_float * 3), ("recv_cap", C.c_int64), ("stage_cap", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("n_particles", C.c_int64), ("cells_per_axis", C.c_int32), ("equal_mass", C.c_int32),
                ("last_evals", C.c_int64), ("last_evals_first", C.c_int64), ("last_members", C.c_int64),
                ("last_kernel_launches", C.c_int32), ("last_deferred", C.c_int32), ("n_in_grid", C.c_int64)]


_lib = None


def lib():
    """Load libsogpu.so (built by __graft_entry__.build() / make -C so_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SoGpuError("CUDA extension missing: %s (run `python -c 'import __graft_entry__ as g; "
                         "g.build()'` or `make -C so_b200/csrc`); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.sogpu_create.restype = C.c_int
    L.sogpu_create.argtypes = [C.POINTER(vp), C.c_int]
    L.sogpu_destroy.restype = None
    L.sogpu_destroy.argtypes = [vp]
    L.sogpu_last_error.restype = C.c_char_p
    L.sogpu_last_error.argtypes = []
    L.sogpu_set_stream.argtypes = [vp, vp]
    L.sogpu_set_cell_occupancy.argtypes = [vp, C.c_float]
    L.sogpu_vcirc.argtypes = [vp, fp, fp, fp, C.c_int32, C.c_float, C.c_int32, fp, fp, fp, fp, fp]
    L.sogpu_vcirc.restype = C.c_int
    L.sogpu_vcirc_species.argtypes = [vp, fp, fp, fp, C.c_int32, C.c_float, C.c_int32, C.POINTER(C.c_ubyte), i32p, C.c_int32,
                                      fp, fp, fp, fp, fp]
    L.sogpu_vcirc_species.restype = C.c_int
    L.sogpu_tag_replay.argtypes = [vp, i32p, C.c_int32, i32p, fp, fp, fp, C.c_int32, C.c_int32, i32p, i32p, i32p, i32p, i32p,
                                   C.POINTER(C.c_ubyte)]
    L.sogpu_tag_replay.restype = C.c_int
    L.sogpu_tag_members.argtypes = [vp, i32p, C.c_int32, C.POINTER(C.c_ubyte), i32p]
    L.sogpu_tag_members.restype = C.c_int
    L.sogpu_vcm.argtypes = [vp, fp, C.c_int32, fp]
    L.sogpu_vcm.restype = C.c_int
    L.sogpu_ingest_keep_velocities.argtypes = [vp, C.c_int]
    L.sogpu_ingest_keep_velocities.restype = C.c_int
    L.sogpu_ingest_begin.argtypes = [vp, C.c_int64, fp, fp]
    L.sogpu_ingest_begin.restype = C.c_int
    L.sogpu_ingest_records.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32]
    L.sogpu_ingest_records.restype = C.c_int
    L.sogpu_ingest_end.argtypes = [vp]
    L.sogpu_ingest_end.restype = C.c_int
    L.sogpu_host_alloc.argtypes = [C.c_size_t]
    L.sogpu_host_alloc.restype = C.c_void_p
    L.sogpu_host_free.argtypes = [vp]
    L.sogpu_domain_mask_words.argtypes = [vp, C.c_int64, i64p]
    L.sogpu_domain_mask.argtypes = [vp, C.c_int64, fp, fp, fp, fp, C.c_int32, C.c_int32, vp]
    L.sogpu_domain_route_count.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int32, i64p]
    L.sogpu_domain_route_scatter.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int64, vp, C.c_int32,
                                             C.POINTER(C.c_void_p), i64p]
    L.sogpu_set_particles_device_indexed.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_float, fp, fp]
    L.sogpu_peer_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p), vp]
    L.sogpu_peer_open.argtypes = [vp, vp, C.POINTER(C.c_void_p)]
    L.sogpu_peer_close.argtypes = [vp, vp]
    L.sogpu_peer_free.argtypes = [vp, vp]
    for f in (L.sogpu_domain_mask_words, L.sogpu_domain_mask, L.sogpu_domain_route_count, L.sogpu_domain_route_scatter,
              L.sogpu_set_particles_device_indexed, L.sogpu_peer_alloc, L.sogpu_peer_open, L.sogpu_peer_close,
              L.sogpu_peer_free):
        f.restype = C.c_int
    L.sogpu_domain_open.argtypes = [vp, C.POINTER(DomainCfg), vp]
    L.sogpu_domain_connect.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sogpu_enable_peer_access.argtypes = [vp, C.c_int]
    L.sogpu_domain_begin.argtypes = [vp, vp, vp, C.c_int32, C.c_int32]
    L.sogpu_domain_route.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.sogpu_domain_route_host.argtypes = [vp, vp, C.c_int64, C.c_int64, vp]
    L.sogpu_domain_push.argtypes = [vp, C.c_int]
    L.sogpu_domain_solve.argtypes = [vp, C.c_float, C.c_int32, vp, vp]
    L.sogpu_domain_result.argtypes = [vp, i64p, i64p, C.POINTER(C.c_uint32), C.POINTER(C.c_ubyte)]
    L.sogpu_domain_close.argtypes = [vp]
    L.sogpu_domain_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sogpu_domain_pointers.restype = C.c_int
    for f in (L.sogpu_domain_open, L.sogpu_domain_connect, L.sogpu_enable_peer_access, L.sogpu_domain_begin,
              L.sogpu_domain_route, L.sogpu_domain_route_host, L.sogpu_domain_push, L.sogpu_domain_solve,
              L.sogpu_domain_result, L.sogpu_domain_close):
        f.restype = C.c_int
    L.sogpu_debug_timeline.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.sogpu_debug_timeline.restype = C.c_int
    L.sogpu_set_tma_staging.argtypes = [vp, C.c_int]
    L.sogpu_set_tma_staging.restype = C.c_int
    L.sogpu_set_first_ball.argtypes = [vp, C.c_int]
    L.sogpu_set_first_ball.restype = C.c_int
    L.sogpu_set_build_mode.argtypes = [vp, C.c_int]
    L.sogpu_set_build_mode.restype = C.c_int
    L.sogpu_set_particles_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int64, fp, fp]
    L.sogpu_set_particles_device.argtypes = [vp, vp, C.c_int64, fp, fp]
    L.sogpu_upload_particles.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int64, vp]
    L.sogpu_upload_particles.restype = C.c_int
    L.sogpu_build_grid.argtypes = [vp]
    L.sogpu_build_grid_for.argtypes = [vp, fp, fp, C.c_int32, C.c_int32]
    L.sogpu_build_grid_for.restype = C.c_int
    L.sogpu_build_grid_for_device.argtypes = [vp, vp, vp, C.c_int32, C.c_int32]
    L.sogpu_build_grid_for_device.restype = C.c_int
    L.sogpu_so.argtypes = [vp, fp, fp, C.c_int32, C.c_float, C.c_int32, fp, fp, i32p]
    L.sogpu_so_device.argtypes = [vp, vp, vp, C.c_int32, C.c_float, C.c_int32, vp, vp]
    L.sogpu_members.argtypes = [vp, i64p, C.POINTER(i32p), C.POINTER(fp), C.c_int]
    L.sogpu_keep_member_d2.argtypes = [vp, C.c_int]
    L.sogpu_finish_host.argtypes = [i32p, fp, C.c_int32, C.c_float, fp, fp, i32p]
    L.sogpu_profile_enable.argtypes = [vp, C.c_int]
    L.sogpu_profile_kernels.restype = C.c_int
    L.sogpu_profile_kernels.argtypes = []
    L.sogpu_profile_name.restype = C.c_char_p
    L.sogpu_profile_name.argtypes = [C.c_int]
    L.sogpu_profile_read.argtypes = [vp, C.POINTER(C.c_double), i64p, C.c_int, C.c_int]
    L.sogpu_profile_bytes.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int]
    L.sogpu_profile_bytes.restype = C.c_int
    L.sogpu_ball_gather.argtypes = [vp, fp, C.c_float, i32p, fp, C.c_int64, i64p]
    L.sogpu_ball_gather_batch.argtypes = [vp, fp, fp, C.c_int32]
    L.sogpu_ball_gather_batch.restype = C.c_int
    L.sogpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.sogpu_mass_prefix.argtypes = [C.c_float, C.c_int64, i64p, C.c_int64, fp]
    L.sogpu_ball_schedule.restype = C.c_int
    L.sogpu_ball_schedule.argtypes = [C.c_float, fp, fp, C.c_int]
    L.sogpu_rdelta.restype = C.c_float
    L.sogpu_rdelta.argtypes = [C.c_float, C.c_float]
    for name in ("sogpu_set_stream", "sogpu_set_cell_occupancy", "sogpu_set_particles_host",
                 "sogpu_set_particles_device", "sogpu_build_grid", "sogpu_so", "sogpu_so_device",
                 "sogpu_members", "sogpu_ball_gather", "sogpu_get_stats", "sogpu_mass_prefix",
                 "sogpu_finish_host", "sogpu_keep_member_d2", "sogpu_profile_enable", "sogpu_profile_read"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check(rc):
    if rc != 0:
        raise SoGpuError("sogpu error %d: %s" % (rc, lib().sogpu_last_error().decode()))


# ---- host-only helpers (no GPU needed) ----------------------------------------------------------

def mass_prefix(m, k):
    """S[k] = sequential fp32 sum of k equal masses m (kd2.c:787,807) from the kernels' table."""
    k = np.ascontiguousarray(k, np.int64)
    out = np.zeros(len(k), np.float32)
    kmax = int(k.max()) if len(k) else 0
    _check(lib().sogpu_mass_prefix(C.c_float(m), kmax, k.ctypes.data_as(C.POINTER(C.c_int64)), len(k), _fp(out)))
    return out


def ball_schedule(rgtp, period=(1.0, 1.0, 1.0)):
    per = np.asarray(period, np.float32)
    out = np.zeros(256, np.float32)
    k = lib().sogpu_ball_schedule(C.c_float(rgtp), _fp(per), _fp(out), 256)
    return out[:k].copy()


def rdelta(mvir, thr):
    return float(lib().sogpu_rdelta(C.c_float(mvir), C.c_float(thr)))


# ---- the handle ---------------------------------------------------------------------------------

class SoGpu:
    """One GPU context: particles -> cell grid -> SO queries."""

    def __init__(self, device=-1, stream=None):
        self._h = C.c_void_p()
        _check(lib().sogpu_create(C.byref(self._h), int(device)))
        if stream is not None:
            _check(lib().sogpu_set_stream(self._h, C.c_void_p(int(stream))))
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().sogpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, stream):
        _check(lib().sogpu_set_stream(self._h, C.c_void_p(int(stream) if stream else 0)))

    def debug_timeline(self):
        out = (C.c_uint64 * 16)()
        _check(lib().sogpu_debug_timeline(self._h, out))
        return list(out)

    def set_tma_staging(self, on=True):
        """1024-thread class: TMA bulk-copy staging instead of per-thread loads (results identical)."""
        _check(lib().sogpu_set_tma_staging(self._h, 1 if on else 0))

    def set_first_ball(self, k):
        """First ball of the reference's schedule that is gathered (results do not depend on it)."""
        _check(lib().sogpu_set_first_ball(self._h, int(k)))

    def set_build_mode(self, mode):
        """-1 auto, 0 single counting sort, 1 coarse partition first."""
        _check(lib().sogpu_set_build_mode(self._h, int(mode)))

    def set_cell_occupancy(self, ppc):
        _check(lib().sogpu_set_cell_occupancy(self._h, C.c_float(ppc)))

    def set_particles(self, pos, mass, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
        """pos: (N,3) float32 (any row stride); mass: scalar or (N,) float32."""
        pos = np.asarray(pos)
        if pos.dtype != np.float32 or pos.ndim != 2 or pos.shape[1] != 3 or pos.strides[1] != 4:
            pos = np.ascontiguousarray(pos, np.float32)
        m = np.asarray(mass, np.float32)
        if m.ndim == 0:
            m = np.full(1, m, np.float32)
            ms = 0
        else:
            if len(m) != len(pos):
                raise ValueError("mass length")
            ms = m.strides[0]
        per = np.asarray(period, np.float32).copy()
        cen = np.asarray(center, np.float32).copy()
        _check(lib().sogpu_set_particles_host(self._h, C.c_void_p(pos.ctypes.data), pos.strides[0],
                                              C.c_void_p(m.ctypes.data), ms, len(pos), _fp(per), _fp(cen)))
        self.n = len(pos)

    def upload_particles(self, pos, mass, d_xyzm_dst):
        """Pack + H2D into a caller-owned device float4 array (e.g. a torch tensor's data_ptr())."""
        pos = np.asarray(pos)
        if pos.dtype != np.float32 or pos.ndim != 2 or pos.shape[1] != 3 or pos.strides[1] != 4:
            pos = np.ascontiguousarray(pos, np.float32)
        m = np.asarray(mass, np.float32)
        if m.ndim == 0:
            m = np.full(1, m, np.float32)
            ms = 0
        else:
            ms = m.strides[0]
        _check(lib().sogpu_upload_particles(self._h, C.c_void_p(pos.ctypes.data), pos.strides[0],
                                            C.c_void_p(m.ctypes.data), ms, len(pos), C.c_void_p(int(d_xyzm_dst))))

    def set_particles_records(self, rec, pos_field="pos", mass_field="mass", period=(1.0, 1.0, 1.0),
                              center=(0.0, 0.0, 0.0)):
        """Structured records (tipsy dark_particle, PINIT, ...) used in place, no repacking."""
        rec = np.ascontiguousarray(rec)
        base = rec.ctypes.data
        po = rec.dtype.fields[pos_field][1]
        mo = rec.dtype.fields[mass_field][1]
        per = np.asarray(period, np.float32).copy()
        cen = np.asarray(center, np.float32).copy()
        _check(lib().sogpu_set_particles_host(self._h, C.c_void_p(base + po), rec.dtype.itemsize,
                                              C.c_void_p(base + mo), rec.dtype.itemsize, len(rec),
                                              _fp(per), _fp(cen)))
        self.n = len(rec)

    def set_particles_device(self, dev_ptr, n, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
        per = np.asarray(period, np.float32).copy()
        cen = np.asarray(center, np.float32).copy()
        _check(lib().sogpu_set_particles_device(self._h, C.c_void_p(int(dev_ptr)), int(n), _fp(per), _fp(cen)))
        self.n = int(n)

    def build_grid(self):
        _check(lib().sogpu_build_grid(self._h))

    def build_grid_for(self, centers, rgtp, n_balls=3):
        """Focused build: keep only what these halos can reach within n_balls schedule steps."""
        centers = np.ascontiguousarray(centers, np.float32)
        rgtp = np.ascontiguousarray(rgtp, np.float32)
        _check(lib().sogpu_build_grid_for(self._h, _fp(centers), _fp(rgtp), len(rgtp), int(n_balls)))

    def build_grid_for_device(self, d_centers, d_rgtp, nh, n_balls=3):
        _check(lib().sogpu_build_grid_for_device(self._h, C.c_void_p(int(d_centers)), C.c_void_p(int(d_rgtp)),
                                                 int(nh), int(n_balls)))

    def so(self, centers, rgtp, thr, n_members=8):
        centers = np.ascontiguousarray(centers, np.float32)
        rgtp = np.ascontiguousarray(rgtp, np.float32)
        nh = len(rgtp)
        assert centers.shape == (nh, 3)
        rv = np.zeros(nh, np.float32)
        mv = np.zeros(nh, np.float32)
        nd = np.zeros(nh, np.int32)
        _check(lib().sogpu_so(self._h, _fp(centers), _fp(rgtp), nh, C.c_float(thr), int(n_members),
                              _fp(rv), _fp(mv), nd.ctypes.data_as(C.POINTER(C.c_int32))))
        self._last_h = nh
        return dict(rvir=rv, mvir=mv, ndelta=nd)

    def so_device(self, d_centers, d_rgtp, nh, thr, n_members=8, d_out_n=0, d_out_m=0):
        _check(lib().sogpu_so_device(self._h, C.c_void_p(int(d_centers)), C.c_void_p(int(d_rgtp)), int(nh),
                                     C.c_float(thr), int(n_members), C.c_void_p(int(d_out_n)),
                                     C.c_void_p(int(d_out_m))))
        self._last_h = int(nh)

    def keep_member_d2(self, on=True):
        """Ask the next so() call to also keep r^2 of every member (needed for sorted lists)."""
        _check(lib().sogpu_keep_member_d2(self._h, 1 if on else 0))

    def members(self, want_d2=False, sorted=False, copy=True):
        """CSR member lists of the last so(): (offsets, members[, d2])."""
        nh = self._last_h
        off = np.zeros(nh + 1, np.int64)
        mp = C.POINTER(C.c_int32)()
        dp = C.POINTER(C.c_float)()
        _check(lib().sogpu_members(self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(mp),
                                   C.byref(dp) if want_d2 else None, 1 if sorted else 0))
        tot = int(off[-1])
        mem = np.ctypeslib.as_array(mp, (max(tot, 1),))[:tot]
        if copy:
            mem = mem.copy()
        if want_d2:
            d2 = np.ctypeslib.as_array(dp, (max(tot, 1),))[:tot]
            return off, mem, (d2.copy() if copy else d2)
        return off, mem

    def finish_host(self, code_or_n, m, thr):
        """rvir/mvir/ndelta from the packed device outputs of so_device()."""
        code_or_n = np.ascontiguousarray(code_or_n, np.int32)
        m = np.ascontiguousarray(m, np.float32)
        nh = len(m)
        rv, mv, nd = np.zeros(nh, np.float32), np.zeros(nh, np.float32), np.zeros(nh, np.int32)
        _check(lib().sogpu_finish_host(code_or_n.ctypes.data_as(C.POINTER(C.c_int32)), _fp(m), nh,
                                       C.c_float(thr), _fp(rv), _fp(mv), nd.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(rvir=rv, mvir=mv, ndelta=nd)

    def profile_enable(self, on=True):
        _check(lib().sogpu_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        """{kernel name: (milliseconds, launches)} accumulated since the last reset."""
        nk = lib().sogpu_profile_kernels()
        ms = (C.c_double * nk)()
        ln = (C.c_int64 * nk)()
        by = (C.c_double * nk)()
        _check(lib().sogpu_profile_read(self._h, ms, ln, nk, 1 if reset else 0))
        _check(lib().sogpu_profile_bytes(self._h, by, nk, 1 if reset else 0))
        return {lib().sogpu_profile_name(k).decode(): (ms[k], ln[k], by[k]) for k in range(nk)}

    def ball_gather(self, center, ball2, cap=None):
        c = np.asarray(center, np.float32).copy()
        n = C.c_int64(0)
        if cap is None:
            _check(lib().sogpu_ball_gather(self._h, _fp(c), C.c_float(ball2), None, None, 0, C.byref(n)))
            cap = n.value
        idx = np.zeros(max(cap, 1), np.int32)
        d2 = np.zeros(max(cap, 1), np.float32)
        _check(lib().sogpu_ball_gather(self._h, _fp(c), C.c_float(ball2),
                                       idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(d2), cap, C.byref(n)))
        k = min(cap, n.value)
        return idx[:k], d2[:k], n.value

    def ball_gather_batch(self, centers, ball2, sorted=True):
        """All particles with fDist2 <= ball2[i] around centers[i]: (offsets, indices, d2)."""
        centers = np.ascontiguousarray(centers, np.float32)
        ball2 = np.ascontiguousarray(ball2, np.float32)
        _check(lib().sogpu_ball_gather_batch(self._h, _fp(centers), _fp(ball2), len(ball2)))
        self._last_h = len(ball2)
        return self.members(want_d2=True, sorted=sorted)

    def vcirc(self, centers, rvir, mvir, G=1.0, n_members=8, profile=True):
        """kdVcirc / kdMassProfile (kd2.c:498-586) for groups with rvir > 0, on the device."""
        centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        rvir = np.ascontiguousarray(rvir, np.float32)
        mvir = np.ascontiguousarray(mvir, np.float32)
        h = len(rvir)
        out = {"vcirc": np.zeros((h, 8), np.float32), "rmass": np.zeros((h, 2), np.float32),
               "rmax": np.zeros(h, np.float32), "vmax": np.zeros(h, np.float32),
               "profile": np.zeros((h, 16), np.float32) if profile else None}
        _check(lib().sogpu_vcirc(self._h, _fp(centers), _fp(rvir), _fp(mvir), h, C.c_float(G), int(n_members),
                                 _fp(out["vcirc"]), _fp(out["rmass"]), _fp(out["rmax"]), _fp(out["vmax"]),
                                 _fp(out["profile"]) if profile else None))
        self._last_h = h
        return out

    def vcirc_species(self, centers, rvir, mvir, ptype=None, masks=(), G=1.0, n_members=8):
        """kdVcirc / kdMassProfile for any particle masses and several species: ptype = species bits per particle
        (uint8[N]), masks = up to 4 bit masks, one mass profile each.  profiles: (len(masks), h, 16)."""
        centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        rvir = np.ascontiguousarray(rvir, np.float32)
        mvir = np.ascontiguousarray(mvir, np.float32)
        h, nm = len(rvir), len(masks)
        out = {"vcirc": np.zeros((h, 8), np.float32), "rmass": np.zeros((h, 2), np.float32),
               "rmax": np.zeros(h, np.float32), "vmax": np.zeros(h, np.float32),
               "profiles": np.zeros((max(nm, 1), h, 16), np.float32)}
        pt = np.ascontiguousarray(ptype, np.uint8) if ptype is not None else None
        mk = np.ascontiguousarray(masks, np.int32) if nm else None
        _check(lib().sogpu_vcirc_species(self._h, _fp(centers), _fp(rvir), _fp(mvir), h, C.c_float(G), int(n_members),
                                         pt.ctypes.data_as(C.POINTER(C.c_ubyte)) if pt is not None else None,
                                         mk.ctypes.data_as(C.POINTER(C.c_int32)) if nm else None, nm,
                                         _fp(out["vcirc"]), _fp(out["rmass"]), _fp(out["rmax"]), _fp(out["vmax"]),
                                         _fp(out["profiles"]) if nm else None))
        self._last_h = h
        return out

    def vcm(self, mvir):
        """_VcmParticles of the last so() call (velocities kept by ingest_records(keep_velocities=True))."""
        mvir = np.ascontiguousarray(mvir, np.float32)
        out = np.zeros((len(mvir), 3), np.float32)
        _check(lib().sogpu_vcm(self._h, _fp(mvir), len(mvir), _fp(out)))
        return out

    def ingest_records(self, blocks, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0), big_endian=False, chunk=1 << 16,
                       keep_velocities=False):
        """Raw TIPSY record blocks [(float32 array (n, floats_per_record)), ...] in file order -> device."""
        _check(lib().sogpu_ingest_keep_velocities(self._h, 1 if keep_velocities else 0))
        n = sum(len(b) for b in blocks)
        per = (C.c_float * 3)(*period)
        cen = (C.c_float * 3)(*center)
        _check(lib().sogpu_ingest_begin(self._h, n, per, cen))
        for b in blocks:
            b = np.ascontiguousarray(b)
            for i0 in range(0, len(b), chunk):
                part = np.ascontiguousarray(b[i0:i0 + chunk])
                _check(lib().sogpu_ingest_records(self._h, C.c_void_p(part.ctypes.data), len(part), b.shape[1],
                                                  1 if big_endian else 0))
        _check(lib().sogpu_ingest_end(self._h))
        self.n = n

    # ---- domain runs (several GPUs) -------------------------------------------------------------
    def domain_mask_words(self, n_total):
        w = C.c_int64()
        _check(lib().sogpu_domain_mask_words(self._h, int(n_total), C.byref(w)))
        return int(w.value)

    def domain_mask(self, n_total, centers, rgtp, n_balls, d_mask, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
        centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        rgtp = np.ascontiguousarray(rgtp, np.float32)
        per, cen = (C.c_float * 3)(*period), (C.c_float * 3)(*center)
        _check(lib().sogpu_domain_mask(self._h, int(n_total), per, cen, _fp(centers) if len(rgtp) else None,
                                       _fp(rgtp) if len(rgtp) else None, len(rgtp), int(n_balls), C.c_void_p(int(d_mask))))

    def domain_route_count(self, n_total, d_slice, n_slice, d_masks, n_ranks):
        counts = np.zeros(n_ranks, np.int64)
        _check(lib().sogpu_domain_route_count(self._h, int(n_total), C.c_void_p(int(d_slice)), int(n_slice),
                                              C.c_void_p(int(d_masks)), int(n_ranks),
                                              counts.ctypes.data_as(C.POINTER(C.c_int64))))
        return counts

    def domain_route_scatter(self, n_total, d_slice, n_slice, index_base, d_masks, dst_ptrs, dst_offsets):
        n_ranks = len(dst_ptrs)
        ptrs = (C.c_void_p * n_ranks)(*[int(p) for p in dst_ptrs])
        offs = np.ascontiguousarray(dst_offsets, np.int64)
        _check(lib().sogpu_domain_route_scatter(self._h, int(n_total), C.c_void_p(int(d_slice)), int(n_slice),
                                                int(index_base), C.c_void_p(int(d_masks)), n_ranks, ptrs,
                                                offs.ctypes.data_as(C.POINTER(C.c_int64))))

    def set_particles_device_indexed(self, d_xyzi, n_local, n_total, mass, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
        per, cen = (C.c_float * 3)(*period), (C.c_float * 3)(*center)
        _check(lib().sogpu_set_particles_device_indexed(self._h, C.c_void_p(int(d_xyzi)), int(n_local), int(n_total),
                                                        C.c_float(float(mass)), per, cen))
        self.n = int(n_local)

    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte handle another process of the node can open)"""
        p = C.c_void_p()
        hbuf = (C.c_ubyte * 64)()
        _check(lib().sogpu_peer_alloc(self._h, int(nbytes), C.byref(p), hbuf))
        return int(p.value), bytes(hbuf)

    def peer_open(self, handle):
        p = C.c_void_p()
        hbuf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _check(lib().sogpu_peer_open(self._h, hbuf, C.byref(p)))
        return int(p.value)

    def peer_close(self, ptr):
        _check(lib().sogpu_peer_close(self._h, C.c_void_p(int(ptr))))

    def peer_free(self, ptr):
        _check(lib().sogpu_peer_free(self._h, C.c_void_p(int(ptr))))

    # ---- domain STEP: stream-ordered from the slice to the results -----------------------------------
    def domain_open(self, rank, n_ranks, n_total, mass, recv_cap, stage_cap, period=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
        """Allocates this rank's buffers; returns the three 64-byte handles (recv0, recv1, ctrl) other processes open."""
        cfg = DomainCfg(int(rank), int(n_ranks), int(n_total), float(mass), (C.c_float * 3)(*period),
                        (C.c_float * 3)(*center), int(recv_cap), int(stage_cap))
        hb = (C.c_ubyte * 192)()
        _check(lib().sogpu_domain_open(self._h, C.byref(cfg), hb))
        raw = bytes(hb)
        return [raw[0:64], raw[64:128], raw[128:192]]

    def domain_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().sogpu_domain_pointers(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def domain_connect(self, recv0, recv1, ctrl):
        n = len(ctrl)
        arr = lambda v: (C.c_void_p * n)(*[int(p) if p else None for p in v])
        _check(lib().sogpu_domain_connect(self._h, arr(recv0), arr(recv1), arr(ctrl)))

    def enable_peer_access(self, peer_device):
        _check(lib().sogpu_enable_peer_access(self._h, int(peer_device)))

    def domain_begin(self, d_centers, d_rgtp, nh, n_balls):
        _check(lib().sogpu_domain_begin(self._h, C.c_void_p(int(d_centers)), C.c_void_p(int(d_rgtp)), int(nh), int(n_balls)))
        self._last_h = int(nh)

    def domain_route(self, d_chunk, n, index_base):
        _check(lib().sogpu_domain_route(self._h, C.c_void_p(int(d_chunk)), int(n), int(index_base)))

    def domain_route_host(self, xyz_pinned_ptr, n, index_base, d_slice_dst):
        _check(lib().sogpu_domain_route_host(self._h, C.c_void_p(int(xyz_pinned_ptr)), int(n), int(index_base),
                                             C.c_void_p(int(d_slice_dst))))

    def domain_push(self, barrier=True):
        _check(lib().sogpu_domain_push(self._h, 1 if barrier else 0))

    def domain_solve(self, thr, n_members=8, d_out_n=0, d_out_m=0):
        _check(lib().sogpu_domain_solve(self._h, C.c_float(float(thr)), int(n_members),
                                        C.c_void_p(int(d_out_n)) if d_out_n else None,
                                        C.c_void_p(int(d_out_m)) if d_out_m else None))

    def domain_result(self, nh=0):
        """Synchronises.  Returns dict(n_recv, n_sent, flags, owner[nh] or None)."""
        nr, ns, fl = C.c_int64(), C.c_int64(), C.c_uint32()
        owner = np.zeros(int(nh), np.uint8) if nh else None
        _check(lib().sogpu_domain_result(self._h, C.byref(nr), C.byref(ns), C.byref(fl),
                                         owner.ctypes.data_as(C.POINTER(C.c_ubyte)) if nh else None))
        return {"n_recv": int(nr.value), "n_sent": int(ns.value), "flags": int(fl.value), "owner": owner}

    def domain_close(self):
        _check(lib().sogpu_domain_close(self._h))

    def tag_members(self, index, n_particles=None):
        """Order-independent part of kdTagParticles: (in_conflict[nh], igrp[N]) for the last so() call."""
        index = np.ascontiguousarray(index, np.int32)
        dirty = np.zeros(len(index), np.uint8)
        igrp = np.zeros(int(n_particles), np.int32) if n_particles else None
        _check(lib().sogpu_tag_members(self._h, index.ctypes.data_as(C.POINTER(C.c_int32)), len(index),
                                       dirty.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                       igrp.ctypes.data_as(C.POINTER(C.c_int32)) if igrp is not None else None))
        return dirty.astype(bool), igrp

    def tag_replay(self, order, index, centers, rvir, mvir, n_particles):
        """Ordered replay of kdTagParticles for the groups in conflict, on the device (after members(sorted=True)
        and tag_members).  Returns dict(rvir, mvir, igrp, nsub, nign, groups_removed, groups_slurped, still_valid)."""
        order = np.ascontiguousarray(order, np.int32)
        index = np.ascontiguousarray(index, np.int32)
        centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        rv = np.array(rvir, np.float32, copy=True)
        mv = np.array(mvir, np.float32, copy=True)
        nh = len(index)
        igrp, nsub, nign = (np.zeros(int(n_particles), np.int32) for _ in range(3))
        rem, slu = C.c_int32(), C.c_int32()
        valid = np.zeros(nh, np.uint8)
        ip = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
        _check(lib().sogpu_tag_replay(self._h, ip(order), len(order), ip(index), _fp(centers), _fp(rv), _fp(mv), nh,
                                      int(index.max()), ip(igrp), ip(nsub), ip(nign), C.byref(rem), C.byref(slu),
                                      valid.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return dict(rvir=rv, mvir=mv, igrp=igrp, nsub=nsub, nign=nign, groups_removed=int(rem.value),
                    groups_slurped=int(slu.value), still_valid=valid.astype(bool))

    def stats(self):
        s = Stats()
        _check(lib().sogpu_get_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_}


# ---- mirror of the reference's hot-path interface (kd2.h:258-276) ----------------------------------

class KD:
    """Python twin of the `KD` handle for the hot path only: particles + halo list in, the fields
    kdSO fills out.  `grps` arrays follow GRPNODE (kd2.h:86-102)."""

    def __init__(self, nMembers=8, fPeriod=(1.0, 1.0, 1.0), fCenter=(0.0, 0.0, 0.0), device=-1, stream=None):
        self.nMembers = int(nMembers)
        self.fPeriod = tuple(float(x) for x in fPeriod)
        self.fCenter = tuple(float(x) for x in fCenter)
        self.gpu = SoGpu(device, stream)
        self.nParticles = 0
        self.nGrps = 0

    def set_particles(self, pos, mass):
        """What kdReadTipsy leaves in kd->pInit (kd2.c:352-416), positions and masses only."""
        self.gpu.set_particles(pos, mass, self.fPeriod, self.fCenter)
        self.nParticles = len(pos)

    def set_groups(self, index, pos, fRgtp, fGTPMass):
        """What kdReadGTPList leaves in kd->grps (kd2.c:245-281)."""
        self.index = np.ascontiguousarray(index, np.int32)
        self.pos = np.ascontiguousarray(pos, np.float32)
        self.fRgtp = np.ascontiguousarray(fRgtp, np.float32)
        self.fGTPMass = np.ascontiguousarray(fGTPMass, np.float32)
        self.nGrps = len(self.index)

    def kdBuildTree(self):
        self.gpu.build_grid()
        return 1

    def smBallGather(self, fBall2, ri):
        return self.gpu.ball_gather(ri, fBall2)

    def kdSO(self, rhovir, nSmooth=1028, sorted_members=False):
        """kdRvir for every group (kd2.c:875-879).  Fills fRvir/fMvir/nDelta and the member lists;
        nSmooth is accepted for signature parity and unused (it only sizes the reference's nnList)."""
        self.gpu.keep_member_d2(sorted_members)
        r = self.gpu.so(self.pos, self.fRgtp, rhovir, self.nMembers)
        self.fRvir, self.fMvir, self.nDelta = r["rvir"], r["mvir"], r["ndelta"]
        self.member_offset, self.members = self.gpu.members(sorted=sorted_members)
        return r

    def close(self):
        self.gpu.close()
