"""ctypes front-end of the CPU oracle (oracle/_build/liborc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  Nothing under d2slam_b200/ imports this module.
The Python surface mirrors d2slam_b200.solver.Solver (minus the window argument) so parity
tests drive both sides with the same calls.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from d2slam_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "liborc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_factors.c", "orc_solver.c", "orc_oracle.h", "orc_math.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "d2ba.h"))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if force or stale:
        # the GPU box has gcc too, but normally the prebuilt .so travels with the snapshot
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.argtypes = [C.POINTER(abi.Config), C.POINTER(C.c_void_p)]
        for name in ("orc_solve", "orc_solve_fixed", "orc_admm_solve", "orc_solve_many"):
            getattr(_LIB, name).restype = C.c_int
    return _LIB


class OrcImuConst(C.Structure):
    _fields_ = [("sum_dt", C.c_double), ("delta_p", C.c_double * 3), ("delta_q", C.c_double * 4),
                ("delta_v", C.c_double * 3), ("linearized_ba", C.c_double * 3),
                ("linearized_bg", C.c_double * 3), ("jacobian", C.c_double * 225),
                ("covariance", C.c_double * 225), ("sqrt_info", C.c_double * 225)]


class OrcObsConst(C.Structure):
    _fields_ = [("pts_i", C.c_double * 3), ("pts_j", C.c_double * 3), ("vel_i", C.c_double * 3),
                ("vel_j", C.c_double * 3), ("td_i", C.c_double), ("td_j", C.c_double),
                ("tangent_base", C.c_double * 6), ("inv_depth_j", C.c_double)]


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


def preintegrate(dt, acc, gyr, ba, bg, acc_n, gyr_n, acc_w, gyr_w):
    """IntegrationBase midpoint pre-integration (integration_base.h:95-199).
    acc/gyr: (n+1, 3) with row 0 = acc_0/gyr_0. Returns dict of numpy arrays."""
    L = lib()
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    gyr = np.ascontiguousarray(gyr, dtype=np.float64)
    ba = np.ascontiguousarray(ba, dtype=np.float64)
    bg = np.ascontiguousarray(bg, dtype=np.float64)
    out = OrcImuConst()
    L.orc_preintegrate(C.c_int(len(dt)), abi.ptr(dt), abi.ptr(acc), abi.ptr(gyr), abi.ptr(ba), abi.ptr(bg),
                       C.c_double(acc_n), C.c_double(gyr_n), C.c_double(acc_w), C.c_double(gyr_w), C.byref(out))
    return {
        "sum_dt": out.sum_dt, "delta_p": np.array(out.delta_p), "delta_q": np.array(out.delta_q),
        "delta_v": np.array(out.delta_v), "linearized_ba": np.array(out.linearized_ba),
        "linearized_bg": np.array(out.linearized_bg), "jacobian": np.array(out.jacobian),
        "covariance": np.array(out.covariance), "sqrt_info": np.array(out.sqrt_info),
    }


class Oracle:
    """One window, CPU, mirrors the C ABI."""

    def __init__(self, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else abi.default_config(**kw)
        self.h = C.c_void_p()
        _chk(lib().orc_create(C.byref(self.cfg), C.byref(self.h)), "create")

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _chk(lib().orc_reset(self.h), "reset")

    def set_blocks(self, kind, ids, values, is_const=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        c = None if is_const is None else np.ascontiguousarray(is_const, dtype=np.uint8)
        _chk(lib().orc_set_blocks(self.h, C.c_int32(kind), C.c_int32(len(ids)), abi.ptr(ids), abi.ptr(values), abi.ptr(c)), "set_blocks")

    def add_proj(self, obs):
        obs = np.ascontiguousarray(obs, dtype=abi.proj_obs_dtype)
        _chk(lib().orc_add_proj(self.h, C.c_int32(len(obs)), abi.ptr(obs)), "add_proj")

    def add_landmark_tracks(self, landmark_ids, track_ptr, tobs, fuse_dep=0, min_d=0.3, max_d=5.0, ignore_frames=()):
        landmark_ids = np.ascontiguousarray(landmark_ids, dtype=np.int64)
        track_ptr = np.ascontiguousarray(track_ptr, dtype=np.int32)
        tobs = np.ascontiguousarray(tobs, dtype=abi.track_obs_dtype)
        ign = np.ascontiguousarray(ignore_frames, dtype=np.int64)
        _chk(lib().orc_add_landmark_tracks(self.h, C.c_int32(len(landmark_ids)), abi.ptr(landmark_ids), abi.ptr(track_ptr),
                                           abi.ptr(tobs), C.c_int32(fuse_dep), C.c_double(min_d), C.c_double(max_d),
                                           C.c_int32(len(ign)), abi.ptr(ign) if len(ign) else None), "add_landmark_tracks")

    def add_imu(self, imu):
        imu = np.ascontiguousarray(imu, dtype=abi.imu_dtype)
        _chk(lib().orc_add_imu(self.h, C.c_int32(len(imu)), abi.ptr(imu)), "add_imu")

    def set_prior(self, J, e0, refs, x0):
        J = np.ascontiguousarray(J, dtype=np.float64); e0 = np.ascontiguousarray(e0, dtype=np.float64)
        refs = np.ascontiguousarray(refs, dtype=abi.blockref_dtype); x0 = np.ascontiguousarray(x0, dtype=np.float64)
        _chk(lib().orc_set_prior(self.h, C.c_int32(len(e0)), abi.ptr(J), abi.ptr(e0), C.c_int32(len(refs)), abi.ptr(refs), abi.ptr(x0)), "set_prior")

    def set_prior_info(self, A, b, refs, x0):
        A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        refs = np.ascontiguousarray(refs, dtype=abi.blockref_dtype); x0 = np.ascontiguousarray(x0, dtype=np.float64)
        _chk(lib().orc_set_prior_info(self.h, C.c_int32(len(b)), abi.ptr(A), abi.ptr(b), C.c_int32(len(refs)), abi.ptr(refs), abi.ptr(x0)), "set_prior_info")

    def set_consensus(self, refs, slots, n_slots_global):
        refs = np.ascontiguousarray(refs, dtype=abi.blockref_dtype)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        _chk(lib().orc_set_consensus(self.h, C.c_int32(len(refs)), abi.ptr(refs), abi.ptr(slots), C.c_int32(n_slots_global)), "set_consensus")

    def finalize(self):
        pass

    def solve(self):
        r = abi.Report()
        _chk(lib().orc_solve(self.h, C.byref(r)), "solve")
        return r

    def solve_fixed(self, iters):
        r = abi.Report()
        _chk(lib().orc_solve_fixed(self.h, C.c_int32(iters), C.byref(r)), "solve_fixed")
        return r

    def get_blocks(self, kind, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.zeros((len(ids), abi.KIND_SIZE[kind]), dtype=np.float64)
        _chk(lib().orc_get_blocks(self.h, C.c_int32(kind), C.c_int32(len(ids)), abi.ptr(ids), abi.ptr(out)), "get_blocks")
        return out

    def get_consensus(self, refs):
        refs = np.ascontiguousarray(refs, dtype=abi.blockref_dtype)
        z = np.zeros((len(refs), 7)); t = np.zeros((len(refs), 6))
        _chk(lib().orc_get_consensus(self.h, C.c_int32(len(refs)), abi.ptr(refs), abi.ptr(z), abi.ptr(t)), "get_consensus")
        return z, t

    def marginalize(self, remove_frame_ids, max_m=1024, max_blk=256):
        rem = np.ascontiguousarray(remove_frame_ids, dtype=np.int64)
        Aflat = np.zeros(max_m * max_m); b = np.zeros(max_m); refs = np.zeros(max_blk, dtype=abi.blockref_dtype); x0 = np.zeros(max_blk * 9)
        m = C.c_int32(); nb = C.c_int32()
        _chk(lib().orc_marginalize_x0(self.h, C.c_int32(len(rem)), abi.ptr(rem), C.byref(m), C.c_int32(max_m), abi.ptr(Aflat), abi.ptr(b),
                                      C.byref(nb), C.c_int32(max_blk), abi.ptr(refs), abi.ptr(x0)), "marginalize")
        mm = m.value
        refs = refs[: nb.value].copy()
        nx = int(sum(abi.KIND_SIZE[int(k)] for k in refs["kind"]))
        return Aflat[: mm * mm].reshape(mm, mm).copy(), b[:mm].copy(), refs, x0[:nx].copy()

    def debug_linearize(self):
        _chk(lib().orc_debug_linearize(self.h), "debug_linearize")

    def debug_get(self, item, dtype=np.float64):
        need = C.c_int64()
        lib().orc_debug_get(self.h, C.c_int32(item), None, C.c_int64(0), C.byref(need))
        out = np.zeros(max(need.value, 1) // np.dtype(dtype).itemsize + 1, dtype=dtype)
        _chk(lib().orc_debug_get(self.h, C.c_int32(item), abi.ptr(out), C.c_int64(out.nbytes), C.byref(need)), "debug_get")
        return out[: need.value // np.dtype(dtype).itemsize]


def admm_solve(agents, fixed_mode=False):
    n = len(agents)
    arr = (C.c_void_p * n)(*[a.h for a in agents])
    reps = (abi.Report * n)()
    _chk(lib().orc_admm_solve(arr, C.c_int32(n), C.c_int32(1 if fixed_mode else 0), reps), "admm_solve")
    return list(reps)


def solve_many(oracles, nthreads, fixed_iters=0):
    n = len(oracles)
    arr = (C.c_void_p * n)(*[a.h for a in oracles])
    reps = (abi.Report * n)()
    _chk(lib().orc_solve_many(arr, C.c_int32(n), C.c_int32(nthreads), C.c_int32(fixed_iters), reps), "solve_many")
    return list(reps)


def admm_many(swarms, nthreads, fixed_mode=True):
    """swarms: list of lists of Oracle (one list per swarm, same length)."""
    ns, na = len(swarms), len(swarms[0])
    flat = [a.h for sw in swarms for a in sw]
    arr = (C.c_void_p * len(flat))(*flat)
    reps = (abi.Report * len(flat))()
    _chk(lib().orc_admm_many(arr, C.c_int32(ns), C.c_int32(na), C.c_int32(nthreads), C.c_int32(1 if fixed_mode else 0), reps), "admm_many")
    return list(reps)
