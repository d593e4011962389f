"""ctypes front-ends of the corridor checkers (TEST / BASELINE INFRASTRUCTURE ONLY).

* `c_poly`, `c_safe_corridor`  - oracle/corridor_oracle.c, the C restatement (built on demand);
* `ref_poly`                   - oracle/_ref/libref_corridor.so, the reference's OWN GetPolyOcta3D /
                                 GetPolyOcta3DNew compiled unmodified (exists only where /root/reference
                                 was present at build time, or where the built .so travelled).
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("HDSM_REFERENCE", "/root/reference")
_REF_SO = os.path.join(_HERE, "_ref", "libref_corridor.so")
_C_SO = os.path.join(_HERE, "libcorridor_oracle.so")
_libs = {}

FLAG_SQUEEZED, FLAG_ROWS, FLAG_SEED_OUT, FLAG_WINDOW = 1, 2, 4, 8
MAX_PLANES = 18


class CorParams(C.Structure):
    _fields_ = [("poly_hor", C.c_int32), ("n_it", C.c_int32), ("rmax", C.c_int32), ("n_traj", C.c_int32),
                ("max_path", C.c_int32), ("reserved", C.c_int32), ("voxel", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "corridor_oracle.c")
    if force or not os.path.exists(_C_SO) or os.path.getmtime(_C_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libcorridor_oracle.so"])
    return _C_SO


def build_ref(force=False):
    """Compile the parts of the reference that build here (convex_decomp.cpp; voxel_grid.cpp + raycast.cpp; map_builder.cpp +
    path_tools.cpp and agent_class.cpp on stand-in ROS / Gurobi headers) into oracle/_ref/ when the checkout is present."""
    src = os.path.join(REFERENCE, "convex_decomp_util", "src", "convex_decomp.cpp")
    if not os.path.exists(src):
        return _REF_SO if os.path.exists(_REF_SO) else None
    voxel_so = os.path.join(_HERE, "_ref", "libref_voxel.so")
    map_so = os.path.join(_HERE, "_ref", "libref_map.so")
    agent_so = os.path.join(_HERE, "_ref", "libref_agent.so")
    if force or not os.path.exists(_REF_SO) or not os.path.exists(voxel_so) or not os.path.exists(map_so) or not os.path.exists(agent_so) or \
            os.path.getmtime(agent_so) < os.path.getmtime(os.path.join(_HERE, "ref_wrap_agent.cpp")) or \
            os.path.getmtime(map_so) < os.path.getmtime(os.path.join(_HERE, "ref_wrap_map.cpp")) or \
            os.path.getmtime(_REF_SO) < os.path.getmtime(os.path.join(_HERE, "ref_wrap.cpp")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "ref", f"REFERENCE={REFERENCE}"])
    return _REF_SO


def have_ref():
    return os.path.exists(_REF_SO)


def _lib(which):
    if which not in _libs:
        _libs[which] = C.CDLL(build() if which == "c" else _REF_SO)
    return _libs[which]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _poly(fn, grid, seed, n_it, res, conv, origin, extra=()):
    grid = np.ascontiguousarray(grid, np.int8)
    dz, dy, dx = grid.shape
    dim = np.array([dx, dy, dz], np.int32)
    seed = np.ascontiguousarray(seed, np.int32)
    origin = np.ascontiguousarray(origin, np.float64)
    pts, nrm = np.zeros((32, 3)), np.zeros((32, 3))
    return grid, dim, seed, origin, pts, nrm


def ref_poly(grid, seed, n_it, res, conv, origin, use_new=False):
    """(points, normals, marked grid) from the reference's GetPolyOcta3D (GetPolyOcta3DNew if use_new)."""
    L = _lib("ref")
    L.ref_get_poly_octa_3d.restype = C.c_int
    grid, dim, seed, origin, pts, nrm = _poly(None, grid, seed, n_it, res, conv, origin)
    out = np.empty_like(grid)
    n = L.ref_get_poly_octa_3d(_p(seed), _p(grid), _p(dim), C.c_int(n_it), C.c_double(res), C.c_int(conv), _p(origin),
                               C.c_int(int(use_new)), _p(pts), _p(nrm), _p(out))
    return pts[:n].copy(), nrm[:n].copy(), out


def c_poly(grid, seed, n_it, res, conv, origin, use_new=False):
    """Same from the C restatement (cor_poly_octa, or cor_poly_octa_new if use_new)."""
    L = _lib("c")
    fn = L.cor_poly_octa_new if use_new else L.cor_poly_octa
    fn.restype = C.c_int
    grid, dim, seed, origin, pts, nrm = _poly(None, grid, seed, n_it, res, conv, origin)
    out = grid.copy()
    n = fn(_p(seed), _p(out), _p(dim), C.c_int(n_it), C.c_double(res), C.c_int(conv), _p(origin), _p(pts), _p(nrm))
    return pts[:n].copy(), nrm[:n].copy(), out


def max_threads():
    return os.cpu_count() or 1


def c_safe_corridor(cb, n_threads=None):
    """GenerateSafeCorridor for a CorridorBatch (multi_agent_pkgs_b200.corridor.CorridorBatch layout)."""
    L = _lib("c")
    P = CorParams(cb.poly_hor, cb.n_it, cb.rmax, cb.prev_traj.shape[1], cb.path.shape[1], int(getattr(cb, "use_cvx_new", False)), cb.voxel)
    n, PH, R = cb.n, cb.poly_hor, cb.rmax
    out = dict(poly_A=np.zeros((n, PH, R, 3)), poly_b=np.zeros((n, PH, R)), poly_rows=np.zeros((n, PH), np.int32),
               seeds=np.zeros((n, PH, 3)), flags=np.zeros(n, np.int32))
    grids = np.ascontiguousarray(cb.grids, np.int8)
    L.cor_safe_corridor_batch(C.byref(P), C.c_int(n), _p(grids), C.c_size_t(grids[0].size), _p(cb.dims), _p(cb.origins),
                              _p(cb.pos), _p(cb.path), _p(cb.n_path), _p(cb.prev_n), _p(cb.prev_A), _p(cb.prev_b),
                              _p(cb.prev_rows), _p(cb.prev_seeds), _p(cb.prev_used), _p(cb.prev_traj),
                              _p(out["poly_A"]), _p(out["poly_b"]), _p(out["poly_rows"]), _p(out["seeds"]),
                              _p(out["flags"]), C.c_int(n_threads or max_threads()))
    return out
