"""Host-side mirror of the reference's reference-trajectory step, on top of the C ABI (include/hdsm.h).

Reference: Agent::GenerateReferenceTrajectory (multi_agent_planner/src/agent_class.cpp:1449-1553) with
SamplePath (:1591-1663), KeepOnlyFreeReference (:1665-1693), ComputePathVelocity (:1695-1803) and the ray
casts of voxel_grid_util::Raycast (voxel_grid_util/src/raycast.cpp:21-186); SURVEY.md 8(f) row 2.
`ReferenceTrajectoryGenerator.generate` is `hdsm_reftraj_batch`; its `ref` output is `traj_ref_curr_`, whose
first N rows are `hdsm_solve_batch`'s `ref` input.

There is no CPU fallback: without the CUDA library / a GPU the generator raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .corridor import OCC, UNKNOWN, clipped_path, local_grid


class HdsmRefTrajParams(C.Structure):
    _fields_ = [("n_hor", C.c_int32), ("max_path", C.c_int32), ("n_traj", C.c_int32), ("reserved", C.c_int32),
                ("dt", C.c_double), ("path_vel_min", C.c_double), ("path_vel_max", C.c_double), ("path_vel_dec", C.c_double),
                ("sens_dist", C.c_double), ("sens_pot", C.c_double), ("sens_other_agents", C.c_double), ("voxel_size", C.c_double)]


@dataclass
class RefTrajBatch:
    """Inputs of one reference-trajectory update for n agents, arrays as hdsm_reftraj_batch takes them."""
    n_hor: int
    dt: float
    voxel: float
    grids: np.ndarray          # [G][dz][dy][dx] int8 (potential-field values 1..99 slow the agent down)
    grid_index: Optional[np.ndarray]
    dims: np.ndarray           # [n][3] int32
    origins: np.ndarray        # [n][3]
    path: np.ndarray           # [n][max_path][3] path_curr_
    n_path: np.ndarray         # [n] int32 (>= 1)
    prev_ref: np.ndarray       # [n][N+1][3]
    have_prev: np.ndarray      # [n] uint8
    increment: np.ndarray      # [n] uint8
    traj: np.ndarray           # [n][n_traj][3]
    global_id: np.ndarray      # [n] int32
    nbr_begin: Optional[np.ndarray]
    nbr_end: Optional[np.ndarray]
    all_pos: np.ndarray        # [n_rob][n_traj][3]
    all_valid: np.ndarray      # [n_rob] uint8
    path_vel_min: float = 4.5  # agent_agile_config.yaml:16-20
    path_vel_max: float = 9.0
    path_vel_dec: float = 0.0
    sens_dist: float = 0.05
    sens_pot: float = 0.18
    sens_other_agents: float = 1.0  # agent_class.cpp:2213

    @property
    def n(self):
        return self.path.shape[0]


def _p(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


class ReferenceTrajectoryGenerator:
    """hdsm_reftraj_create / hdsm_reftraj_batch / hdsm_reftraj_destroy."""

    def __init__(self, rb: RefTrajBatch, max_agents=None, max_grids=None, device=0):
        self.L = _lib.load()
        L = self.L
        for f in (L.hdsm_reftraj_create, L.hdsm_reftraj_batch, L.hdsm_reftraj_batch_device):
            f.restype = C.c_int
        L.hdsm_reftraj_last_error.restype = C.c_char_p
        L.hdsm_reftraj_launch_count.restype = C.c_int64
        self.prm = HdsmRefTrajParams(rb.n_hor, rb.path.shape[1], rb.traj.shape[1], 0, rb.dt, rb.path_vel_min, rb.path_vel_max,
                                     rb.path_vel_dec, rb.sens_dist, rb.sens_pot, rb.sens_other_agents, rb.voxel)
        self.grid_stride = int(rb.grids[0].size)
        self.h = C.c_void_p()
        rc = L.hdsm_reftraj_create(C.byref(self.prm), C.c_int(max_agents or rb.n), C.c_int(max_grids or rb.grids.shape[0]),
                                   C.c_size_t(self.grid_stride), C.c_int(device), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError(f"hdsm_reftraj_create failed ({rc}): needs a CUDA device, max_path <= 32, n_hor <= 12")

    @property
    def launch_count(self):
        return int(self.L.hdsm_reftraj_launch_count(self.h))

    def generate(self, rb: RefTrajBatch):
        n, N1 = rb.n, self.prm.n_hor + 1
        G = rb.grids.shape[0]
        grids = np.ascontiguousarray(rb.grids.reshape(G, -1))
        assert grids.shape[1] == self.grid_stride
        ref, vel = np.zeros((n, N1, 6)), np.zeros(n)
        f8, i4, u1 = np.float64, np.int32, np.uint8
        rc = self.L.hdsm_reftraj_batch(
            self.h, C.c_int(n), C.c_int(G), _p(grids, np.int8), _p(rb.grid_index, i4), _p(rb.dims, i4), _p(rb.origins, f8),
            _p(rb.path, f8), _p(rb.n_path, i4), _p(rb.prev_ref, f8), _p(rb.have_prev, u1), _p(rb.increment, u1), _p(rb.traj, f8),
            _p(rb.global_id, i4), _p(rb.nbr_begin, i4), _p(rb.nbr_end, i4), _p(rb.all_pos, f8), _p(rb.all_valid, u1),
            C.c_int(rb.all_pos.shape[0]), _p(ref, f8), _p(vel, f8))
        if rc != 0:
            raise RuntimeError(f"hdsm_reftraj_batch failed ({rc}): {self.L.hdsm_reftraj_last_error(self.h).decode()}")
        return dict(ref=ref, path_vel=vel)

    def generate_device(self, t, n, n_rob, stream_ptr=0):
        """hdsm_reftraj_batch_device on a dict of torch tensors (keys as RefTrajBatch fields plus the outputs
        ref, path_vel and optionally ref_solver); stream ordered, no synchronisation."""
        def dp(k):
            v = t.get(k)
            return None if v is None else C.c_void_p(v.data_ptr())
        rc = self.L.hdsm_reftraj_batch_device(
            self.h, C.c_int(n), dp("grids"), dp("grid_index"), dp("dims"), dp("origins"), dp("path"), dp("n_path"), dp("prev_ref"),
            dp("have_prev"), dp("increment"), dp("traj"), dp("global_id"), dp("nbr_begin"), dp("nbr_end"), dp("all_pos"),
            dp("all_valid"), C.c_int(n_rob), dp("ref"), dp("ref_solver"), dp("path_vel"), C.c_void_p(stream_ptr))
        if rc != 0:
            raise RuntimeError(f"hdsm_reftraj_batch_device failed ({rc}): {self.L.hdsm_reftraj_last_error(self.h).decode()}")

    def close(self):
        if self.h is not None:
            self.L.hdsm_reftraj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------
# synthetic inputs (stand-ins for mapping_util's potential field and the path planner)
# --------------------------------------------------------------------------------------
def add_potential_field(grid, voxel=0.3, potential_dist=1.5, power=4.0):
    """Free voxels within potential_dist of an occupied / unknown one get (1 - d / potential_dist)^power * 100
    (mapping_util/config: potential_dist 1.5, potential_pow 4; voxel_grid.cpp:192-226), occupied stay 100."""
    from scipy import ndimage
    blocked = (grid == OCC) | (grid == UNKNOWN)
    d = ndimage.distance_transform_edt(~blocked) * voxel
    h = (100.0 * np.clip(1.0 - d / potential_dist, 0.0, 1.0) ** power).astype(np.int8)
    out = grid.copy()
    free = grid == 0
    out[free] = h[free]
    return out


def reftraj_batch(sw, prev_ref=None, max_path=16, voxel=0.3, ids=None):
    """RefTrajBatch for the agents of a scenarios.Swarm; prev_ref [n][N+1][3] = last step's reference positions."""
    ids = np.arange(sw.n) if ids is None else np.asarray(ids)
    n, N = len(ids), sw.params["n_hor"]
    grids, dims, origins = [], np.zeros((n, 3), np.int32), np.zeros((n, 3))
    path, n_path = np.zeros((n, max_path, 3)), np.zeros(n, np.int32)
    for r, i in enumerate(ids):
        pos = sw.state[i, :3]
        g, o = local_grid(sw.world, pos, voxel)
        grids.append(add_potential_field(g, voxel))
        dims[r], origins[r] = (g.shape[2], g.shape[1], g.shape[0]), o
        start = pos if prev_ref is None else prev_ref[r, 1]
        pts = np.vstack([start[None, :], clipped_path(start, sw.goal[i], sw.world)])[:max_path]
        path[r, :len(pts)], n_path[r] = pts, len(pts)
    traj = np.repeat(sw.state[:, None, :3], N + 1, 1) if sw.traj is None else np.ascontiguousarray(sw.traj[:, :, :3])
    have = np.zeros(sw.n, np.uint8) if sw.have_plan is None else sw.have_plan.astype(np.uint8)
    return RefTrajBatch(N, sw.params["dt"], voxel, np.stack(grids), None, dims, origins, path, n_path,
                        np.zeros((n, N + 1, 3)) if prev_ref is None else np.ascontiguousarray(prev_ref),
                        np.full(n, 0 if prev_ref is None else 1, np.uint8), np.ones(n, np.uint8),
                        np.ascontiguousarray(traj[ids]), ids.astype(np.int32), sw.group_begin[ids].astype(np.int32),
                        sw.group_end[ids].astype(np.int32), np.ascontiguousarray(traj), have)
