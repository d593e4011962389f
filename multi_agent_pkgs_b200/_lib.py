"""ctypes binding of the C ABI in include/hdsm.h (the same stub a cgo / JNI / ROS-side caller writes)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build

STATUS_NAMES = {0: "OPTIMAL", 1: "INFEASIBLE", 2: "MAX_ITER", 3: "NUMERICAL", 4: "NODE_LIMIT", 5: "ROW_OVERFLOW"}
OPTIMAL, INFEASIBLE, MAX_ITER, NUMERICAL, NODE_LIMIT, ROW_OVERFLOW = range(6)

EXPORTS = ["hdsm_version", "hdsm_create", "hdsm_destroy", "hdsm_last_error", "hdsm_solve_batch",
           "hdsm_solve_batch_device", "hdsm_launch_count", "hdsm_smem_bytes", "hdsm_comm_unique_id",
           "hdsm_comm_init", "hdsm_allgather_positions", "hdsm_comm_destroy", "hdsm_exchange_plans", "hdsm_advance_device", "hdsm_planes",
           "hdsm_corridor_create", "hdsm_corridor_destroy", "hdsm_corridor_last_error", "hdsm_corridor_launch_count",
           "hdsm_corridor_smem_bytes", "hdsm_corridor_batch", "hdsm_corridor_batch_device",
           "hdsm_reftraj_create", "hdsm_reftraj_destroy", "hdsm_reftraj_last_error", "hdsm_reftraj_launch_count",
           "hdsm_reftraj_batch", "hdsm_reftraj_batch_device",
           "hdsm_map_create", "hdsm_map_destroy", "hdsm_map_last_error", "hdsm_map_launch_count", "hdsm_map_batch",
           "hdsm_map_batch_device", "hdsm_map_distance_table",
           "hdsm_sense_grid_dims", "hdsm_sense_create", "hdsm_sense_destroy", "hdsm_sense_last_error",
           "hdsm_sense_launch_count", "hdsm_sense_batch", "hdsm_sense_batch_device"]


class HdsmParams(C.Structure):
    _fields_ = [("n_hor", C.c_int32), ("poly_hor", C.c_int32), ("max_rows_per_poly", C.c_int32), ("rk4", C.c_int32),
                ("max_iter", C.c_int32), ("max_nodes", C.c_int32), ("prune", C.c_int32), ("search_width", C.c_int32),
                ("dt", C.c_double), ("drag", C.c_double * 3), ("r_u", C.c_double), ("r_x", C.c_double * 6),
                ("r_n", C.c_double * 6), ("max_vel", C.c_double), ("min_acc_xy", C.c_double),
                ("max_acc_xy", C.c_double), ("min_acc_z", C.c_double), ("max_acc_z", C.c_double),
                ("max_jerk", C.c_double), ("drone_radius", C.c_double), ("drone_z_offset", C.c_double),
                ("tilt", C.c_double), ("tol", C.c_double), ("warm_start", C.c_int32), ("reserved", C.c_int32)]


class HdsmResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("nodes", C.c_int32), ("rows", C.c_int32),
                ("obj", C.c_double), ("kkt_res", C.c_double)]


RESULT_DTYPE = np.dtype([("status", "i4"), ("iters", "i4"), ("nodes", "i4"), ("rows", "i4"),
                         ("obj", "f8"), ("kkt_res", "f8")])
assert RESULT_DTYPE.itemsize == C.sizeof(HdsmResult)

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build: bool = True) -> C.CDLL:
    """Load libhdsm.so (building it first when stale).  Raises if it cannot be had - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HDSM_LIB") or _build.LIB   # HDSM_LIB: another build of the same ABI (A/B experiments)
    if build and path == _build.LIB:
        _build.build_lib()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp, ip, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.hdsm_version.restype = C.c_int
    L.hdsm_create.restype = C.c_int
    L.hdsm_create.argtypes = [C.POINTER(HdsmParams), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.hdsm_destroy.restype = None
    L.hdsm_destroy.argtypes = [vp]
    L.hdsm_last_error.restype = C.c_char_p
    L.hdsm_last_error.argtypes = [vp]
    host_args = [vp, C.c_int, ip, ip, ip, dp, dp, dp, dp, ip, dp, dp, u8p, C.c_int, ip, dp, dp, u8p, ip,
                 C.POINTER(HdsmResult)]
    L.hdsm_solve_batch.restype = C.c_int
    L.hdsm_solve_batch.argtypes = host_args
    L.hdsm_solve_batch_device.restype = C.c_int
    L.hdsm_solve_batch_device.argtypes = [vp, C.c_int] + [vp] * 11 + [C.c_int] + [vp] * 8
    L.hdsm_launch_count.restype = C.c_int64
    L.hdsm_launch_count.argtypes = [vp]
    L.hdsm_smem_bytes.restype = C.c_int
    L.hdsm_smem_bytes.argtypes = [vp]
    L.hdsm_comm_unique_id.restype = C.c_int
    L.hdsm_comm_unique_id.argtypes = [u8p]
    L.hdsm_comm_init.restype = C.c_int
    L.hdsm_comm_init.argtypes = [vp, C.c_int, C.c_int, u8p]
    L.hdsm_allgather_positions.restype = C.c_int
    L.hdsm_allgather_positions.argtypes = [vp, vp, vp, C.c_int, vp]
    L.hdsm_exchange_plans.restype = C.c_int
    L.hdsm_exchange_plans.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.hdsm_advance_device.restype = C.c_int
    L.hdsm_advance_device.argtypes = [vp, C.c_int] + [vp] * 9
    L.hdsm_planes.restype = C.c_int
    L.hdsm_planes.argtypes = [vp, C.c_int, dp, dp, dp]
    L.hdsm_comm_destroy.restype = None
    L.hdsm_comm_destroy.argtypes = [vp]
    _lib = L
    return L


def make_params(d, rmax=18, max_iter=60, max_nodes=64, prune=True, tol=1e-8, width=1, warm_start=False) -> HdsmParams:
    p = HdsmParams()
    p.n_hor, p.poly_hor, p.max_rows_per_poly, p.rk4 = int(d["n_hor"]), int(d["poly_hor"]), int(rmax), int(bool(d["rk4"]))
    p.max_iter, p.max_nodes, p.prune, p.search_width = int(max_iter), int(max_nodes), int(bool(prune)), int(width)
    p.dt = float(d["dt"])
    p.drag[:] = [float(x) for x in d["drag"]]
    p.r_u = float(d["r_u"])
    p.r_x[:] = [float(x) for x in d["r_x"][:6]]
    p.r_n[:] = [float(x) for x in d["r_n"][:6]]
    for k in ("max_vel", "min_acc_xy", "max_acc_xy", "min_acc_z", "max_acc_z", "max_jerk", "drone_radius",
              "drone_z_offset", "tilt"):
        setattr(p, k, float(d[k]))
    p.tol = float(tol)
    p.warm_start = int(bool(warm_start))
    return p
