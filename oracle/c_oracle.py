"""ctypes front-end of oracle/hdsm_oracle.c (TEST / BASELINE INFRASTRUCTURE ONLY).

Builds libhdsm_oracle.so on demand with oracle/Makefile.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's CPU-baseline legs - never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS_NAMES = {0: "OPTIMAL", 1: "INFEASIBLE", 2: "MAX_ITER", 3: "NUMERICAL", 4: "NODE_LIMIT"}


class OrcParams(C.Structure):
    _fields_ = [("n_hor", C.c_int32), ("poly_hor", C.c_int32), ("rk4", C.c_int32), ("max_iter", C.c_int32),
                ("max_nodes", C.c_int32), ("prune", C.c_int32), ("width", C.c_int32), ("warm_start", C.c_int32),
                ("dt", C.c_double), ("drag", C.c_double * 3), ("r_u", C.c_double), ("r_x", C.c_double * 6),
                ("r_n", C.c_double * 6), ("max_vel", C.c_double), ("min_acc_xy", C.c_double),
                ("max_acc_xy", C.c_double), ("min_acc_z", C.c_double), ("max_acc_z", C.c_double),
                ("max_jerk", C.c_double), ("drone_radius", C.c_double), ("drone_z_offset", C.c_double),
                ("tilt", C.c_double), ("tol", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("nodes", C.c_int32), ("rows", C.c_int32),
                ("obj", C.c_double), ("kkt", C.c_double)]


RESULT_DTYPE = np.dtype([("status", "i4"), ("iters", "i4"), ("nodes", "i4"), ("rows", "i4"),
                         ("obj", "f8"), ("kkt", "f8")])


def build(force=False):
    so = os.path.join(_HERE, "libhdsm_oracle.so")
    src = os.path.join(_HERE, "hdsm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libhdsm_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_solve_batch.restype = C.c_int
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def make_params(d, max_iter=60, max_nodes=100000, prune=True, tol=1e-8, width=1, warm_start=False):
    p = OrcParams()
    p.n_hor, p.poly_hor, p.rk4 = int(d["n_hor"]), int(d["poly_hor"]), int(bool(d["rk4"]))
    p.max_iter, p.max_nodes, p.prune, p.width, p.warm_start = max_iter, max_nodes, int(prune), int(width), int(bool(warm_start))
    p.dt = d["dt"]
    p.drag[:] = d["drag"]
    p.r_u = d["r_u"]
    p.r_x[:] = d["r_x"][:6]
    p.r_n[:] = d["r_n"][:6]
    for k in ("max_vel", "min_acc_xy", "max_acc_xy", "min_acc_z", "max_acc_z", "max_jerk", "drone_radius",
              "drone_z_offset", "tilt"):
        setattr(p, k, float(d[k]))
    p.tol = tol
    return p


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def solve_batch(batch, assign_in=None, n_threads=0, **kw):
    """batch: multi_agent_pkgs_b200.scenarios.Batch-like object.  Returns dict of outputs."""
    p = make_params(batch.params, **kw)
    n, N, P = batch.x0.shape[0], p.n_hor, p.poly_hor
    c = np.ascontiguousarray
    gid, nb0, nb1 = c(batch.global_id, np.int32), c(batch.nbr_begin, np.int32), c(batch.nbr_end, np.int32)
    x0, ref = c(batch.x0, np.float64), c(batch.ref, np.float64)
    pA, pb, pr = c(batch.poly_A, np.float64), c(batch.poly_b, np.float64), c(batch.poly_rows, np.int32)
    prev, allp, allv = c(batch.prev_self_pos, np.float64), c(batch.all_pos, np.float64), c(batch.all_valid, np.uint8)
    ain = c(assign_in, np.int32) if assign_in is not None else None
    traj = np.zeros((n, N + 1, 9))
    ctrl = np.zeros((n, N, 3))
    used = np.zeros((n, P), np.uint8)
    aout = np.zeros((n, N), np.int32)
    res = np.zeros(n, RESULT_DTYPE)
    rc = lib().orc_solve_batch(C.byref(p), n, _ptr(gid, C.c_int32), _ptr(nb0, C.c_int32), _ptr(nb1, C.c_int32),
                               _ptr(x0, C.c_double), _ptr(ref, C.c_double), _ptr(pA, C.c_double),
                               _ptr(pb, C.c_double), _ptr(pr, C.c_int32), int(pA.shape[2]), _ptr(prev, C.c_double),
                               _ptr(allp, C.c_double), _ptr(allv, C.c_uint8), int(allp.shape[0]),
                               _ptr(ain, C.c_int32), _ptr(traj, C.c_double), _ptr(ctrl, C.c_double),
                               _ptr(used, C.c_uint8), _ptr(aout, C.c_int32),
                               res.ctypes.data_as(C.POINTER(OrcResult)), int(n_threads))
    if rc:
        raise RuntimeError(f"orc_solve_batch failed: {rc}")
    return dict(traj=traj, ctrl=ctrl, poly_used=used, assign=aout, res=res)


def plane(params, pc, po):
    p = make_params(params)
    pc, po = np.asarray(pc, float).copy(), np.asarray(po, float).copy()
    out = np.zeros(4)
    lib().orc_plane(C.byref(p), _ptr(pc, C.c_double), _ptr(po, C.c_double), _ptr(out, C.c_double))
    return out[:3], out[3]


def max_threads():
    return lib().orc_max_threads()
