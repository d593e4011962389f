"""ctypes front-ends of the reference-trajectory checkers (TEST / BASELINE INFRASTRUCTURE ONLY).

* `c_raycast`, `c_generate`  - oracle/reftraj_oracle.c (built on demand);
* `ref_raycast`              - oracle/_ref/libref_voxel.so: the reference's OWN voxel_grid_util::Raycast
                               (raycast.cpp + voxel_grid.cpp compiled unmodified), where it was built.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SO = os.path.join(_HERE, "libreftraj_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_voxel.so")
_libs = {}


class RefTrajParams(C.Structure):
    _fields_ = [("n_hor", C.c_int32), ("max_path", C.c_int32), ("n_traj", C.c_int32), ("reserved", C.c_int32),
                ("dt", C.c_double), ("path_vel_min", C.c_double), ("path_vel_max", C.c_double), ("path_vel_dec", C.c_double),
                ("sens_dist", C.c_double), ("sens_pot", C.c_double), ("sens_other_agents", C.c_double), ("voxel", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "reftraj_oracle.c")
    if force or not os.path.exists(_C_SO) or os.path.getmtime(_C_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libreftraj_oracle.so"])
    return _C_SO


def have_ref():
    return os.path.exists(_REF_SO)


def _lib(which):
    if which not in _libs:
        _libs[which] = C.CDLL(build() if which == "c" else _REF_SO)
    return _libs[which]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _raycast(fn, grid, start, end, max_dist, cap=1600):
    grid = np.ascontiguousarray(grid, np.int8)
    dim = np.array([grid.shape[2], grid.shape[1], grid.shape[0]], np.int32)
    start, end = np.ascontiguousarray(start, np.float64), np.ascontiguousarray(end, np.float64)
    vis, col = np.zeros((cap, 3)), np.zeros(3)
    fn.restype = C.c_int
    n = fn(_p(grid), _p(dim), _p(start), _p(end), C.c_double(max_dist), _p(vis), C.c_int(cap), _p(col))
    return vis[:max(n, 0)].copy(), col, n


def ref_raycast(grid, start, end, max_dist):
    return _raycast(_lib("ref").ref_raycast, grid, start, end, max_dist)


def c_raycast(grid, start, end, max_dist):
    return _raycast(_lib("c").rt_raycast, grid, start, end, max_dist)


def max_threads():
    return os.cpu_count() or 1


def make_params(rb):
    return RefTrajParams(rb.n_hor, rb.path.shape[1], rb.traj.shape[1], 0, rb.dt, rb.path_vel_min, rb.path_vel_max, rb.path_vel_dec,
                         rb.sens_dist, rb.sens_pot, rb.sens_other_agents, rb.voxel)


def c_generate(rb, n_threads=None):
    """GenerateReferenceTrajectory for a RefTrajBatch (multi_agent_pkgs_b200.reftraj.RefTrajBatch layout)."""
    L = _lib("c")
    P = make_params(rb)
    n, N1 = rb.n, rb.n_hor + 1
    ref, vel = np.zeros((n, N1, 6)), np.zeros(n)
    G = rb.grids.shape[0]
    grids = np.ascontiguousarray(rb.grids.reshape(G, -1))
    L.rt_generate_batch(C.byref(P), C.c_int(n), _p(grids), C.c_size_t(grids.shape[1]), _p(rb.grid_index), _p(rb.dims), _p(rb.origins),
                        _p(rb.path), _p(rb.n_path), _p(rb.prev_ref), _p(rb.have_prev), _p(rb.increment), _p(rb.traj),
                        _p(rb.global_id), _p(rb.nbr_begin), _p(rb.nbr_end), _p(rb.all_pos), _p(rb.all_valid),
                        C.c_int(rb.all_pos.shape[0]), _p(ref), _p(vel), C.c_int(n_threads or max_threads()))
    return dict(ref=ref, path_vel=vel)
