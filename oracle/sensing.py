"""ctypes front-ends of the local-map acquisition checkers (TEST / BASELINE INFRASTRUCTURE ONLY).

* `c_update`   - oracle/sense_oracle.c: sequential restatement of map_builder.cpp:80-205 (crop, RaycastAndClear, merge);
* `emu_update` - oracle/sense_emu.cpp: the kernel's own per-ray code (csrc/hdsm_sense_core.h) run on the CPU in a
                 scrambled ray order.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SO = os.path.join(_HERE, "libsense_oracle.so")
_EMU_SO = os.path.join(_HERE, "libsense_emu.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_map.so")
_libs = {}


def build(force=False):
    for so, src, deps in ((_C_SO, "sense_oracle.c", ("reftraj_oracle.c",)),
                          (_EMU_SO, "sense_emu.cpp", ("../multi_agent_pkgs_b200/csrc/hdsm_sense_core.h",))):
        newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in (src,) + deps)
        if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", os.path.basename(so)])
    return _C_SO, _EMU_SO


def _lib(which):
    if which not in _libs:
        if which == "ref":
            _libs[which] = C.CDLL(_REF_SO)
        else:
            c_so, emu_so = build()
            _libs[which] = C.CDLL(c_so if which == "c" else emu_so)
    return _libs[which]


def have_ref():
    """oracle/_ref/libref_map.so: the reference's OWN map_builder.cpp (+ path_tools.cpp, raycast.cpp, voxel_grid.cpp) compiled
    unmodified on stand-in ROS / Eigen headers (make -C oracle ref; needs /root/reference)."""
    return os.path.exists(_REF_SO)


def ref_update(env, origin_env, pos, voxel, rng, free_grid=False, rot=None, fov=None, old_grid=None, old_origin=None,
               inflation=0.3, potential=1.5, power=4.0, times_ms=None):
    """One MapBuilder::EnvironmentVoxelGridCallback of ONE agent in the reference itself.  Returns (voxel_grid_curr_ after
    the call [dz][dy][dx], its origin (3,), the published grid [dz][dy][dx]).  times_ms: optional float64 array [3] receiving the node's
    own timers of that call (ray casting, merge, whole callback) in milliseconds."""
    env = np.ascontiguousarray(env, np.int8)
    dim_env = np.array([env.shape[2], env.shape[1], env.shape[0]], np.int32)
    origin_env = np.ascontiguousarray(origin_env, np.float64)
    pos = np.ascontiguousarray(pos, np.float64).reshape(3)
    dx, dy, dz = local_dims(voxel, rng)
    cur, pub, org = np.empty((dz, dy, dx), np.int8), np.empty((dz, dy, dx), np.int8), np.empty(3)
    rot = None if rot is None else np.ascontiguousarray(rot, np.float64).reshape(9)
    if old_grid is not None:
        old_grid = np.ascontiguousarray(old_grid, np.int8).reshape(dz, dy, dx)
        old_origin = np.ascontiguousarray(old_origin, np.float64).reshape(3)
    L = _lib("ref")
    L.ref_map_update.restype = C.c_int
    n = L.ref_map_update((C.c_double * 3)(*rng), C.c_int(int(free_grid)), C.c_double(inflation), C.c_double(potential), C.c_double(power),
                         C.c_int(int(fov is not None)), C.c_double(fov[0] if fov else 1.57), C.c_double(fov[1] if fov else 1.57),
                         _p(env), _p(dim_env), _p(origin_env), C.c_double(voxel), _p(pos), _p(rot), _p(old_grid), _p(old_origin),
                         _p(cur), _p(org), _p(pub), _p(times_ms))
    if n != cur.size:
        raise RuntimeError(f"ref_map_update returned {n}")
    return cur, org, pub


class SenseParams(C.Structure):
    _fields_ = [("voxel", C.c_double), ("range", C.c_double * 3), ("free_grid", C.c_int32), ("limited_fov", C.c_int32),
                ("cos_half_fov_x", C.c_double), ("cos_half_fov_y", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads():
    return os.cpu_count() or 1


def local_dims(voxel, rng):
    """(dx, dy, dz) of the local grid: floor(range / voxel) (map_builder.cpp:103-107)."""
    return tuple(int(math.floor(r / voxel)) for r in rng)


def _prep(env, origin_env, pos, rot, old_grids, old_origin, have_old, voxel, rng):
    env = np.ascontiguousarray(env, np.int8)
    dim_env = np.array([env.shape[2], env.shape[1], env.shape[0]], np.int32)
    pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
    n = pos.shape[0]
    dx, dy, dz = local_dims(voxel, rng)
    out = np.empty((n, dz, dy, dx), np.int8)
    origin_out = np.empty((n, 3), np.float64)
    if rot is not None:
        rot = np.ascontiguousarray(rot, np.float64).reshape(n, 9)
    if old_grids is not None:
        old_grids = np.ascontiguousarray(old_grids, np.int8).reshape(n, dz, dy, dx)
        old_origin = np.ascontiguousarray(old_origin, np.float64).reshape(n, 3)
        have_old = np.ones(n, np.uint8) if have_old is None else np.ascontiguousarray(have_old, np.uint8)
    else:
        old_origin = have_old = None
    return env, dim_env, np.ascontiguousarray(origin_env, np.float64), pos, n, rot, old_grids, old_origin, have_old, out, origin_out


def c_update(env, origin_env, pos, voxel, rng, free_grid=False, rot=None, fov=None, old_grids=None, old_origin=None, have_old=None,
             n_threads=None):
    """env [ez][ey][ex] int8; pos [n][3]; returns (grids [n][dz][dy][dx], origins [n][3])."""
    env, dim_env, origin_env, pos, n, rot, old_grids, old_origin, have_old, out, origin_out = _prep(
        env, origin_env, pos, rot, old_grids, old_origin, have_old, voxel, rng)
    P = SenseParams(voxel, (C.c_double * 3)(*rng), int(free_grid), int(fov is not None),
                    math.cos(fov[0] / 2) if fov else 0.0, math.cos(fov[1] / 2) if fov else 0.0)
    L = _lib("c")
    L.sense_update_batch.restype = C.c_int
    rc = L.sense_update_batch(C.byref(P), C.c_int(n), _p(env), _p(dim_env), _p(origin_env), _p(pos), _p(rot), _p(old_grids), _p(old_origin),
                              _p(have_old), C.c_size_t(out[0].size), _p(out), _p(origin_out), C.c_int(n_threads or max_threads()))
    if rc:
        raise RuntimeError("sense_update_batch: a ray crossed more than 1500 voxels (the reference throws)")
    return out, origin_out


def emu_update(env, origin_env, pos, voxel, rng, free_grid=False, rot=None, fov=None, old_grids=None, old_origin=None, have_old=None, seed=1,
               bits_form=False, conflicts=None):
    """conflicts: optional int64 array [n] receiving, with bits_form, the number of voxels written both free and occupied."""
    env, dim_env, origin_env, pos, n, rot, old_grids, old_origin, have_old, out, origin_out = _prep(
        env, origin_env, pos, rot, old_grids, old_origin, have_old, voxel, rng)
    L = _lib("emu")
    L.sense_emu_batch.restype = C.c_int
    rc = L.sense_emu_batch(C.c_double(voxel), (C.c_double * 3)(*rng), C.c_int(int(free_grid)), C.c_int(int(fov is not None)),
                           C.c_double(math.cos(fov[0] / 2) if fov else 0.0), C.c_double(math.cos(fov[1] / 2) if fov else 0.0), C.c_int(n),
                           _p(env), _p(dim_env), _p(origin_env), _p(pos), _p(rot), _p(old_grids), _p(old_origin), _p(have_old),
                           C.c_size_t(out[0].size), _p(out), _p(origin_out), C.c_uint(seed), C.c_int(int(bits_form)), _p(conflicts))
    if rc:
        raise RuntimeError("sense_emu_batch failed")
    return out, origin_out
