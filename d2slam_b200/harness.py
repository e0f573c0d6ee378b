"""ctypes front-end of libd2ba_harness.so: the C++ stand-in for D2Estimator's solve sequence
(reset -> add blocks / residuals -> finalize -> solve -> read back), used for end-to-end timing."""
import ctypes as C
import os

import numpy as np

from . import abi
from .solver import lib as d2ba_lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        d2ba_lib()
        p = os.path.join(_HERE, "libd2ba_harness.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} not built: run `python -m d2slam_b200.build`")
        _LIB = C.CDLL(p)
        _LIB.rp_create.restype = C.c_void_p
        _LIB.rp_run.restype = C.c_double
        _LIB.rp_run_pipelined.restype = C.c_double
    return _LIB


class Replay:
    def __init__(self, problems):
        self.n = len(problems)
        self.ctx = C.c_void_p(lib().rp_create(C.c_int(self.n)))
        self._keep = []
        for w, p in enumerate(problems):
            self._set(w, p)

    def _set(self, w, p):
        def a(x, dt):
            v = np.ascontiguousarray(x, dtype=dt); self._keep.append(v); return v
        fi = a(p["frame_ids"], np.int64); po = a(p["poses"], np.float64); pc = a(p["pose_const"], np.uint8)
        ci = a(p["cam_ids"], np.int64); ex = a(p["ext"], np.float64); ec = a(p["ext_const"], np.uint8)
        si = a(p["sb_ids"], np.int64); sb = a(p["sb"], np.float64)
        li = a(p["lm_ids"], np.int64); lm = a(p["inv_dep"], np.float64)
        ob = a(p["obs"], abi.proj_obs_dtype); im = a(p["imu"], abi.imu_dtype)
        if p.get("prior") is not None:
            A, b, refs, x0 = p["prior"]
            A = a(A, np.float64); b = a(b, np.float64); refs = a(refs, abi.blockref_dtype); x0 = a(x0, np.float64)
            pm, pn, xl = len(b), len(refs), x0.size
        else:
            A = b = refs = x0 = None; pm = pn = xl = 0
        if p.get("consensus") is not None:
            cr, sl, ns = p["consensus"]
            cr = a(cr, abi.blockref_dtype); sl = a(sl, np.int32); nc = len(cr)
        else:
            cr = sl = None; nc = 0; ns = 0
        rc = lib().rp_set_window(self.ctx, C.c_int(w), C.c_int(len(fi)), abi.ptr(fi), abi.ptr(po), abi.ptr(pc), C.c_int(len(ci)), abi.ptr(ci),
                                 abi.ptr(ex), abi.ptr(ec), C.c_int(len(si)), abi.ptr(si), abi.ptr(sb), C.c_double(float(p["td"])),
                                 C.c_int(int(p["td_const"])), C.c_int(len(li)), abi.ptr(li), abi.ptr(lm), C.c_int(len(ob)), abi.ptr(ob),
                                 C.c_int(len(im)), abi.ptr(im), C.c_int(pm), abi.ptr(A), abi.ptr(b), C.c_int(pn), abi.ptr(refs), C.c_int(xl),
                                 abi.ptr(x0), C.c_int(nc), abi.ptr(cr), abi.ptr(sl), C.c_int(ns))
        if rc:
            raise RuntimeError(f"rp_set_window rc={rc}")

    def run(self, solver, steps, iters, nthreads):
        reps = (abi.Report * self.n)()
        t = lib().rp_run(self.ctx, solver.h, C.c_int(steps), C.c_int(iters), C.c_int(nthreads), reps)
        if t < 0:
            raise RuntimeError(f"rp_run failed: {t} ({d2ba_lib().d2ba_last_error(solver.h)})")
        bd = np.zeros(8)
        lib().rp_breakdown(self.ctx, abi.ptr(bd))
        self.breakdown = {"feed_s": bd[0], "finalize_s": bd[1], "solve_s": bd[2], "fetch_s": bd[3]}
        # thread-summed seconds inside each group of C-ABI calls of the feed phase (diagnostics)
        self.feed_calls = {"set_blocks_s": bd[4], "add_proj_s": bd[5], "add_imu_s": bd[6], "prior_cons_s": bd[7]}
        return t, list(reps)

    def run_pipelined(self, solvers, steps, iters, nthreads):
        """Same per-step sequence with consecutive steps overlapped across ``len(solvers)`` handles (feed | finalize | solve+fetch)."""
        reps = (abi.Report * self.n)()
        hs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
        t = lib().rp_run_pipelined(self.ctx, hs, C.c_int(len(solvers)), C.c_int(steps), C.c_int(iters), C.c_int(nthreads), reps)
        if t < 0:
            raise RuntimeError(f"rp_run_pipelined failed: {t} " + " | ".join(str(d2ba_lib().d2ba_last_error(s.h)) for s in solvers))
        bd = np.zeros(8)
        lib().rp_breakdown(self.ctx, abi.ptr(bd))
        self.breakdown = {"feed_s": bd[0], "finalize_s": bd[1], "solve_s": bd[2], "fetch_s": bd[3]}   # stage busy times; stages overlap
        return t, list(reps)

    def outputs(self, w, n_pose, n_sb, n_lm):
        po = np.zeros((n_pose, 7)); sb = np.zeros((n_sb, 9)); lm = np.zeros(n_lm)
        lib().rp_get_outputs(self.ctx, C.c_int(w), abi.ptr(po), abi.ptr(sb), abi.ptr(lm))
        return po, sb, lm

    def close(self):
        if self.ctx:
            lib().rp_destroy(self.ctx); self.ctx = C.c_void_p()
