"""ctypes binding of oracle/_ref/liboracle.so (the CPU restatement) and helpers to run the
reference binaries in oracle/_ref.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
product package so_b200.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "liboracle.so")


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/so_ref*."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


class _Res(C.Structure):
    _fields_ = [("rvir", C.c_float), ("mvir", C.c_float), ("ndelta", C.c_int32),
                ("ngather", C.c_int32), ("nevals", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        fp, i64, i32p = C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_int32)
        L.so_oracle_create.restype = C.c_void_p
        L.so_oracle_create.argtypes = [fp, i64, fp, i64, i64, fp]
        L.so_oracle_destroy.argtypes = [C.c_void_p]
        L.so_oracle_rho_enclosed.restype = C.c_float
        L.so_oracle_rho_enclosed.argtypes = [C.c_float, C.c_float]
        L.so_oracle_dist2.restype = C.c_float
        L.so_oracle_dist2.argtypes = [fp, fp, fp]
        L.so_oracle_rdelta.restype = C.c_float
        L.so_oracle_rdelta.argtypes = [C.c_float, C.c_float]
        L.so_oracle_schedule.restype = C.c_int
        L.so_oracle_schedule.argtypes = [C.c_float, fp, fp, C.c_int]
        L.so_oracle_ball.restype = i64
        L.so_oracle_ball.argtypes = [C.c_void_p, fp, C.c_float]
        L.so_oracle_ball_index.restype = i32p
        L.so_oracle_ball_index.argtypes = [C.c_void_p]
        L.so_oracle_ball_d2.restype = fp
        L.so_oracle_ball_d2.argtypes = [C.c_void_p]
        L.so_oracle_rvir.restype = C.c_int
        L.so_oracle_rvir.argtypes = [C.c_void_p, fp, C.c_float, C.c_float, C.c_int, C.POINTER(_Res)]
        L.so_oracle_so.restype = C.c_int
        L.so_oracle_so.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_float, C.c_int, fp, fp, i32p,
                                   C.POINTER(i64), C.POINTER(i32p), C.POINTER(i64)]
        L.so_oracle_free.argtypes = [C.c_void_p]
        L.so_oracle_indexx.argtypes = [C.c_int, fp, i32p]
        L.so_oracle_tag.restype = C.c_int
        L.so_oracle_tag.argtypes = [C.c_void_p, C.c_int, i32p, fp, fp, fp, fp, C.POINTER(i64), i32p,
                                    fp, i64, i32p, i32p, i32p, fp, i32p]
        L.so_oracle_vcirc.restype = C.c_int
        L.so_oracle_vcirc.argtypes = [C.c_void_p, fp, C.c_float, C.c_float, C.c_float, C.c_int, fp, fp, fp, fp, fp,
                                      C.POINTER(C.c_ubyte), C.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def rho_enclosed(mass, r2):
    return float(lib().so_oracle_rho_enclosed(C.c_float(mass), C.c_float(r2)))


def rdelta(mvir, thr):
    return float(lib().so_oracle_rdelta(C.c_float(mvir), C.c_float(thr)))


def schedule(rgtp, period=(1.0, 1.0, 1.0)):
    per = np.asarray(period, np.float32)
    out = np.zeros(256, np.float32)
    k = lib().so_oracle_schedule(C.c_float(rgtp), _fp(per), _fp(out), 256)
    return out[:k].copy()


def indexx(arr):
    arr = np.ascontiguousarray(arr, np.float32)
    out = np.zeros(len(arr), np.int32)
    lib().so_oracle_indexx(len(arr), _fp(arr), _ip(out))
    return out


class Oracle:
    """Particle set + the reference's kdRvir / kdSO loop restated on the CPU."""

    def __init__(self, pos, mass, period=(1.0, 1.0, 1.0)):
        self.pos = np.ascontiguousarray(pos, np.float32)
        assert self.pos.ndim == 2 and self.pos.shape[1] == 3
        self.n = len(self.pos)
        m = np.asarray(mass, np.float32)
        if m.ndim == 0:
            self.mass = np.full(1, m, np.float32)
            ms = 0
        else:
            self.mass = np.ascontiguousarray(m)
            ms = 1
        self.period = np.asarray(period, np.float32).copy()
        self._h = lib().so_oracle_create(_fp(self.pos), 3, _fp(self.mass), ms, self.n, _fp(self.period))
        if not self._h:
            raise MemoryError("so_oracle_create")

    def close(self):
        if self._h:
            lib().so_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dist2(self, c, p):
        c = np.asarray(c, np.float32)
        p = np.asarray(p, np.float32)
        return float(lib().so_oracle_dist2(_fp(c), _fp(p), _fp(self.period)))

    def ball(self, c, ball2):
        c = np.asarray(c, np.float32)
        n = lib().so_oracle_ball(self._h, _fp(c), C.c_float(ball2))
        if n < 0:
            raise MemoryError
        idx = np.ctypeslib.as_array(lib().so_oracle_ball_index(self._h), (max(n, 1),))[:n].copy()
        d2 = np.ctypeslib.as_array(lib().so_oracle_ball_d2(self._h), (max(n, 1),))[:n].copy()
        return idx, d2

    def vcirc(self, centers, rvir, mvir, G=1.0, n_members=8, ptype_of=None, ptype_mask=0xFF):
        """kdVcirc + kdMassProfile (kd2.c:498-586, 458-496) for every group with rvir > 0."""
        centers = np.ascontiguousarray(centers, np.float32).reshape(-1, 3)
        h = len(centers)
        out = {"vcirc": np.zeros((h, 8), np.float32), "rmass": np.zeros((h, 2), np.float32),
               "rmax": np.zeros(h, np.float32), "vmax": np.zeros(h, np.float32),
               "profile": np.zeros((h, 16), np.float32)}
        pt = None
        if ptype_of is not None:
            ptype_of = np.ascontiguousarray(ptype_of, np.uint8)
            pt = ptype_of.ctypes.data_as(C.POINTER(C.c_ubyte))
        for i in range(h):
            if not rvir[i] > 0:
                continue
            rm, vm = C.c_float(), C.c_float()
            rc = lib().so_oracle_vcirc(self._h, _fp(centers[i]), C.c_float(rvir[i]), C.c_float(mvir[i]), C.c_float(G),
                                       int(n_members), _fp(out["vcirc"][i]), _fp(out["rmass"][i]),
                                       C.cast(C.byref(rm), C.POINTER(C.c_float)), C.cast(C.byref(vm), C.POINTER(C.c_float)),
                                       _fp(out["profile"][i]), pt, int(ptype_mask))
            if rc:
                raise RuntimeError("so_oracle_vcirc failed")
            out["rmax"][i], out["vmax"][i] = rm.value, vm.value
        return out

    def rvir(self, c, rgtp, thr, n_members=8):
        c = np.asarray(c, np.float32)
        r = _Res()
        rc = lib().so_oracle_rvir(self._h, _fp(c), C.c_float(rgtp), C.c_float(thr), n_members, C.byref(r))
        if rc:
            raise RuntimeError("so_oracle_rvir rc=%d" % rc)
        mem = None
        if r.ndelta > 0:
            mem = np.ctypeslib.as_array(lib().so_oracle_ball_index(self._h), (r.ndelta,)).copy()
        return dict(rvir=r.rvir, mvir=r.mvir, ndelta=r.ndelta, ngather=r.ngather, nevals=r.nevals,
                    members=mem)

    def so(self, centers, rgtp, thr, n_members=8, want_members=True):
        centers = np.ascontiguousarray(centers, np.float32)
        rgtp = np.ascontiguousarray(rgtp, np.float32)
        h = len(rgtp)
        rv = np.zeros(h, np.float32)
        mv = np.zeros(h, np.float32)
        nd = np.zeros(h, np.int32)
        off = np.zeros(h + 1, np.int64)
        memp = C.POINTER(C.c_int32)()
        nev = C.c_int64(0)
        rc = lib().so_oracle_so(self._h, _fp(centers), _fp(rgtp), h, C.c_float(thr), n_members,
                                _fp(rv), _fp(mv), _ip(nd), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                C.byref(memp) if want_members else None, C.byref(nev))
        if rc:
            raise RuntimeError("so_oracle_so rc=%d" % rc)
        members = None
        if want_members:
            tot = int(off[-1])
            members = np.ctypeslib.as_array(memp, (max(tot, 1),))[:tot].copy()
            lib().so_oracle_free(memp)
        return dict(rvir=rv, mvir=mv, ndelta=nd, member_offset=off, members=members, nevals=nev.value)

    def tag(self, index, centers, gtp_mass, rvir, mvir, member_offset, members, vel=None):
        """Replay kdTagParticles in kdSO order.  rvir/mvir are copied and returned modified."""
        h = len(index)
        index = np.ascontiguousarray(index, np.int32)
        centers = np.ascontiguousarray(centers, np.float32)
        gtp_mass = np.ascontiguousarray(gtp_mass, np.float32)
        rv = np.array(rvir, np.float32, copy=True)
        mv = np.array(mvir, np.float32, copy=True)
        off = np.ascontiguousarray(member_offset, np.int64)
        mem = np.ascontiguousarray(members, np.int32)
        igrp = np.zeros(self.n, np.int32)
        nsub = np.zeros(self.n, np.int32)
        nign = np.zeros(self.n, np.int32)
        vcm = np.zeros((h, 3), np.float32)
        counts = np.zeros(2, np.int32)
        if vel is not None:
            vel = np.ascontiguousarray(vel, np.float32)
        rc = lib().so_oracle_tag(self._h, h, _ip(index), _fp(centers), _fp(gtp_mass), _fp(rv), _fp(mv),
                                 off.ctypes.data_as(C.POINTER(C.c_int64)), _ip(mem),
                                 _fp(vel) if vel is not None else None, 3, _ip(igrp), _ip(nsub),
                                 _ip(nign), _fp(vcm), _ip(counts))
        if rc:
            raise RuntimeError("so_oracle_tag rc=%d" % rc)
        return dict(rvir=rv, mvir=mv, igrp=igrp, nsub=nsub, nign=nign, vcm=vcm,
                    groups_removed=int(counts[0]), groups_slurped=int(counts[1]))


# ---- running the reference binaries (oracle/_ref) -----------------------------------------------

def ref_available(which="so_ref"):
    return os.path.exists(os.path.join(REF_DIR, which))


def run_so_ref(snap_path, gtp_path, out_base, delta=None, extra=(), inst=False, inst_file=None,
               cwd=None):
    """Run the reference program: so_ref -i gtp -o out [-delta D] -grp -gtp < snap.
    Returns dict(stderr=..., so_cpu_time=..., ndist=..., ngather=...)."""
    exe = os.path.join(REF_DIR, "so_ref_inst" if inst else "so_ref")
    cmd = [exe, "-i", gtp_path, "-o", out_base]
    if delta is not None:
        cmd += ["-delta", repr(float(delta))]
    cmd += list(extra)
    env = dict(os.environ)
    if inst_file:
        env["SO_INST_FILE"] = inst_file
    else:
        env.pop("SO_INST_FILE", None)
    with open(snap_path, "rb") as fin:
        r = subprocess.run(cmd, stdin=fin, capture_output=True, text=True, env=env, cwd=cwd)
    if r.returncode != 0:
        raise RuntimeError("so_ref failed rc=%d\n%s" % (r.returncode, r.stderr[-2000:]))
    out = dict(stderr=r.stderr, so_cpu_time=None, ndist=None, ngather=None)
    for line in r.stderr.splitlines():
        if line.startswith("SO CPU Time:"):
            out["so_cpu_time"] = float(line.split(":")[1])
        if line.startswith("SO_INST"):
            for tok in line.split()[1:]:
                k, v = tok.split("=")
                out[k] = int(v)
    return out


def read_inst_file(path):
    """Records written by so_ref_inst: {index: (j, iOrder[j], fDist2[j])}."""
    out = {}
    if not os.path.exists(path):      # no halo succeeded: the hook never opened the file
        return out
    with open(path, "rb") as f:
        buf = f.read()
    off = 0
    while off < len(buf):
        index, j = np.frombuffer(buf, np.int32, 2, off)
        off += 8
        order = np.frombuffer(buf, np.int32, int(j), off).copy()
        off += 4 * int(j)
        d2 = np.frombuffer(buf, np.float32, int(j), off).copy()
        off += 4 * int(j)
        out[int(index)] = (int(j), order, d2)
    return out


def run_so_ref_timed(snap_path, gtp_path, thr, n_members=8, period=1.0, out_base=None):
    exe = os.path.join(REF_DIR, "so_ref_timed")
    cmd = [exe, snap_path, gtp_path, repr(float(thr)), str(int(n_members)), repr(float(period))]
    if out_base:
        cmd.append(out_base)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("so_ref_timed failed rc=%d\n%s" % (r.returncode, r.stderr[-2000:]))
    return json.loads(r.stdout.strip().splitlines()[-1])


# ---- the reference's default threshold (so.c:57-86, 470-481): restated here as test infrastructure ----------

def virial_threshold(omega0, flat_lambda, time=None, z=None):
    """fThreshold exactly as so.c computes it when no -delta is given: Kitayama & Suto (1996) virial overdensity
    (so.c:57-86) times Omega0 (so.c:478), with z = 1/h.time - 1 from the snapshot header unless -z sets it
    (so.c:470-472).  fOmega, fRedshift, kd->fTime and fThreshold are floats in the reference (so.c:200, kd2.h:119);
    the formula itself runs in double."""
    import math
    f_omega = np.float32(omega0)
    if z is None:
        f_time = np.float32(time)
        z = np.float32(1.0 / float(f_time) - 1.0)             # (1.0/kd->fTime)-1.0 in double, stored in a float
    else:
        z = np.float32(z)
    om, zz = float(f_omega), float(z)

    def omegaf(omega, lam, zv):                               # so.c:57-66
        zp2 = (1.0 + zv) * (1.0 + zv)
        zp3 = zp2 * (1.0 + zv)
        return omega * zp3 / (omega * zp3 + (1.0 - omega - lam) * zp2 + lam)

    if om == 1.0:
        ratio = 178.0
    elif flat_lambda:
        wf = 1.0 / omegaf(om, 1.0 - om, zz) - 1.0
        ratio = 18.0 * (math.pi * math.pi) * (1.0 + 0.4093 * math.pow(wf, 0.9052))
    else:
        etaf = math.acosh(2.0 / omegaf(om, 0.0, zz) - 1.0)
        ratio = 4.0 * (math.pi * math.pi) / math.pow(math.sinh(etaf) - etaf, 2)
        ratio *= math.pow(math.cosh(etaf) - 1.0, 3)
    return np.float32(ratio * om)                             # double product, stored in a float
