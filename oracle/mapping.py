"""ctypes front-ends of the local-map checkers (TEST / BASELINE INFRASTRUCTURE ONLY).

* `c_mask`, `c_inflate_potential`, `c_process`  - oracle/map_oracle.c (built on demand);
* `ref_mask`, `ref_inflate_potential`           - oracle/_ref/libref_voxel.so: the reference's OWN VoxelGrid
                                                   (voxel_grid.cpp compiled unmodified), where it was built.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SO = os.path.join(_HERE, "libmap_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_voxel.so")
_libs = {}


def build(force=False):
    src = os.path.join(_HERE, "map_oracle.c")
    if force or not os.path.exists(_C_SO) or os.path.getmtime(_C_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmap_oracle.so"])
    return _C_SO


def have_ref():
    if not os.path.exists(_REF_SO):
        return False
    return hasattr(_lib("ref"), "ref_inflate_and_potential")


def _lib(which):
    if which not in _libs:
        _libs[which] = C.CDLL(build() if which == "c" else _REF_SO)
    return _libs[which]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _mask(fn, vox, dist, power):
    off, val = np.zeros((4096, 3), np.int32), np.zeros(4096, np.int8)
    fn.restype = C.c_int
    n = fn(C.c_double(vox), C.c_double(dist), C.c_double(power), _p(off), _p(val), C.c_int(4096))
    return off[:n].copy(), val[:n].copy()


def ref_mask(vox, dist, power):
    return _mask(_lib("ref").ref_create_mask, vox, dist, power)


def c_mask(vox, dist, power):
    return _mask(_lib("c").map_create_mask, vox, dist, power)


def _dim(grid):
    return np.array([grid.shape[2], grid.shape[1], grid.shape[0]], np.int32)


def ref_inflate_potential(grid, vox, inflation, potential, power):
    out = np.ascontiguousarray(grid, np.int8).copy()
    _lib("ref").ref_inflate_and_potential(_p(out), _p(_dim(out)), C.c_double(vox), C.c_double(inflation), C.c_double(potential), C.c_int(power))
    return out


def c_inflate_potential(grid, vox, inflation, potential, power):
    out = np.ascontiguousarray(grid, np.int8).copy()
    L = _lib("c")
    if inflation > 0:
        L.map_inflate(_p(out), _p(_dim(out)), C.c_double(vox), C.c_double(inflation))
    if potential > 0:
        L.map_potential(_p(out), _p(_dim(out)), C.c_double(vox), C.c_double(potential), C.c_int(power))
    return out


def max_threads():
    return os.cpu_count() or 1


def c_process(grids, vox, inflation, potential, power, n_threads=None):
    """The whole per-step sequence (SetUncertainToUnknown, InflateObstacles, CreatePotentialField) on [n][dz][dy][dx]."""
    grids = np.ascontiguousarray(grids, np.int8)
    out = np.empty_like(grids)
    n = grids.shape[0]
    dims = np.tile(_dim(grids[0]), (n, 1)).astype(np.int32)
    _lib("c").map_process_batch(C.c_int(n), _p(grids), _p(out), C.c_size_t(grids[0].size), _p(dims), C.c_double(vox),
                                C.c_double(inflation), C.c_double(potential), C.c_int(power), C.c_int(n_threads or max_threads()))
    return out
