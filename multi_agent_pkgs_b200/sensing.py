"""Host-side mirror of the reference's per-agent local-map acquisition, on top of the C ABI (include/hdsm.h).

Reference: mapping_util/src/map_builder.cpp - MapBuilder::EnvironmentVoxelGridCallback (:80-205: frame, crop of the
environment grid, RaycastAndClear :280-329 / ClearLine :367-432, MergeVoxelGrids :242-278, ClearVoxelsCenter :434-447);
SURVEY.md 8(f) row 4.  `LocalMapBuilder.update` is `hdsm_sense_batch`; like the node it keeps `voxel_grid_curr_`
(grids and origins of the previous update) between calls.  Its output goes to `mapping.MapProcessor.process`
(SetUncertainToUnknown, InflateObstacles, CreatePotentialField) and from there to the corridor / reference generators.

There is no CPU fallback: without the CUDA library / a GPU `LocalMapBuilder` raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class HdsmSenseParams(C.Structure):
    _fields_ = [("voxel_size", C.c_double), ("range", C.c_double * 3), ("free_grid", C.c_int32), ("limited_fov", C.c_int32),
                ("fov_x", C.c_double), ("fov_y", C.c_double)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grid_dims(voxel_size, grid_range):
    """(dx, dy, dz) of the local grid, floor(range / voxel) (map_builder.cpp:103-107), from hdsm_sense_grid_dims."""
    L = _lib.load()
    L.hdsm_sense_grid_dims.restype = C.c_int
    prm = HdsmSenseParams(voxel_size, (C.c_double * 3)(*grid_range), 0, 0, 0.0, 0.0)
    dim = (C.c_int32 * 3)()
    if L.hdsm_sense_grid_dims(C.byref(prm), dim) != 0:
        raise ValueError("voxel_size and grid_range must be positive and give at least one voxel per axis")
    return tuple(dim)


class LocalMapBuilder:
    """hdsm_sense_create / hdsm_sense_batch / hdsm_sense_destroy (mapping_util/config defaults: range 20 x 20 x 6 m,
    fov 1.57 / 1.57 rad; `free_grid=False` is the unknown-environment mode of multi_agent_planner_long)."""

    def __init__(self, voxel_size, max_agents, grid_range=(20.0, 20.0, 6.0), free_grid=False, limited_fov=False, fov_x=1.57,
                 fov_y=1.57, device=0):
        self.L = _lib.load()
        L = self.L
        for f in (L.hdsm_sense_create, L.hdsm_sense_batch, L.hdsm_sense_batch_device, L.hdsm_sense_grid_dims):
            f.restype = C.c_int
        L.hdsm_sense_last_error.restype = C.c_char_p
        L.hdsm_sense_launch_count.restype = C.c_int64
        self.prm = HdsmSenseParams(voxel_size, (C.c_double * 3)(*grid_range), int(free_grid), int(limited_fov), fov_x, fov_y)
        self.dims = grid_dims(voxel_size, grid_range)
        self.grid_stride = int(np.prod(self.dims))
        self.max_agents = int(max_agents)
        self.h = C.c_void_p()
        rc = L.hdsm_sense_create(C.byref(self.prm), C.c_int(max_agents), C.c_size_t(self.grid_stride), C.c_int(device), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError(f"hdsm_sense_create failed ({rc}): needs a CUDA device; grid sides may not sum to more than 1400 voxels")
        self.grids = None    # voxel_grid_curr_ of every agent, [n][dz][dy][dx]
        self.origins = None  # their origins, [n][3]

    @property
    def launch_count(self):
        return int(self.L.hdsm_sense_launch_count(self.h))

    def reset(self):
        """Forget the kept grids: the next update is every agent's first (map_builder.cpp:160-167)."""
        self.grids = self.origins = None

    def update(self, env, origin_env, pos, rot=None):
        """One map update of all agents.  env [ez][ey][ex] int8 (the environment grid they share), origin_env (3,),
        pos [n][3]; rot [n][3][3] = rot_mat_cam_ with limited_fov.  Returns (grids [n][dz][dy][dx], origins [n][3])
        and keeps them for the next call."""
        env = np.ascontiguousarray(env, np.int8)
        dim_env = np.array([env.shape[2], env.shape[1], env.shape[0]], np.int32)
        origin_env = np.ascontiguousarray(origin_env, np.float64)
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        n = pos.shape[0]
        if rot is not None:
            rot = np.ascontiguousarray(rot, np.float64).reshape(n, 9)
        dx, dy, dz = self.dims
        out = np.empty((n, dz, dy, dx), np.int8)
        origins = np.empty((n, 3), np.float64)
        old, old_org, have = self.grids, self.origins, None
        if old is not None:
            if old.shape[0] != n:
                raise ValueError("the number of agents changed between updates: call reset() first")
            have = np.ones(n, np.uint8)
        rc = self.L.hdsm_sense_batch(self.h, C.c_int(n), _p(env), _p(dim_env), _p(origin_env), _p(pos), _p(rot), _p(old), _p(old_org),
                                     _p(have), _p(out), _p(origins))
        if rc != 0:
            raise RuntimeError(f"hdsm_sense_batch failed ({rc}): {self.L.hdsm_sense_last_error(self.h).decode()}")
        self.grids, self.origins = out, origins
        return out, origins

    def update_device(self, t_env, dim_env, origin_env, t_pos, t_rot, t_old, t_old_origin, t_have, t_out, t_origin_out, stream_ptr=0):
        """Device twin on torch tensors (int8 / float64 / uint8, contiguous, on the handle's device); dim_env and
        origin_env are host sequences.  t_rot / t_old / t_old_origin / t_have may be None."""
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = self.L.hdsm_sense_batch_device(self.h, C.c_int(t_pos.shape[0]), ptr(t_env), (C.c_int32 * 3)(*[int(v) for v in dim_env]),
                                            (C.c_double * 3)(*[float(v) for v in origin_env]), ptr(t_pos), ptr(t_rot), ptr(t_old),
                                            ptr(t_old_origin), ptr(t_have), ptr(t_out), ptr(t_origin_out), C.c_void_p(stream_ptr))
        if rc != 0:
            raise RuntimeError(f"hdsm_sense_batch_device failed ({rc}): {self.L.hdsm_sense_last_error(self.h).decode()}")

    def close(self):
        if self.h is not None:
            self.L.hdsm_sense_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def environment_grid(world, voxel=0.3, col_half=0.05, z_range=(0.0, 6.0), margin=3.0):
    """The environment grid env_builder would publish for a `scenarios.Forest` (env_builder/src/environment_builder.cpp:
    189-231: columns voxelised without inflation): (env [ez][ey][ex] int8 with 0 free / 100 occupied, origin (3,))."""
    cols = np.asarray(world.cols, np.float64).reshape(-1, 2)
    lo = np.array([-margin, -margin, z_range[0]])
    hi = np.array([margin, margin, z_range[1]])
    if len(cols):
        lo[:2], hi[:2] = cols.min(0) - margin, cols.max(0) + margin
    origin = np.floor(lo / voxel) * voxel
    dim = np.ceil((hi - origin) / voxel).astype(int)
    env = np.zeros((dim[2], dim[1], dim[0]), np.int8)
    for cx, cy in cols:
        x0, x1 = int(np.floor((cx - col_half - origin[0]) / voxel)), int(np.floor((cx + col_half - origin[0]) / voxel))
        y0, y1 = int(np.floor((cy - col_half - origin[1]) / voxel)), int(np.floor((cy + col_half - origin[1]) / voxel))
        env[:, max(y0, 0):y1 + 1, max(x0, 0):x1 + 1] = 100
    return env, origin
