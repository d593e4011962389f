"""Host-side mirror of the reference's safe-corridor step, on top of the C ABI (include/hdsm.h).

Reference: Agent::GenerateSafeCorridor (multi_agent_planner/src/agent_class.cpp:1236-1447) calling
convex_decomp_lib::GetPolyOcta3D / GetPolyOcta3DNew (convex_decomp_util/src/convex_decomp.cpp:5-376, :590-1162),
SURVEY.md 8(f) row 1.
`SafeCorridorGenerator.generate` is `hdsm_corridor_batch`; its outputs (`poly_A`, `poly_b`, `poly_rows`) are
`hdsm_solve_batch`'s polytope inputs, `seeds` are the reference's `poly_seeds_`.

The voxel grids come from the ROS mapping nodes in the reference (mapping_util, out of scope); the
helpers at the bottom fabricate local grids of the same shape for the synthetic scenarios
(voxel 0.3 m, range 20 x 20 x 6 m around the agent, obstacles inflated by 0.3 m:
mapping_util/config/*.yaml, map_builder.cpp:90-122, :207-212).

There is no CPU fallback: without the CUDA library / a GPU `SafeCorridorGenerator` raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .scenarios import Forest, plan_path

FLAG_SQUEEZED, FLAG_ROW_OVERFLOW, FLAG_SEED_OUTSIDE, FLAG_LIST_OVERFLOW = 1, 2, 4, 8  # SQUEEZED is informational
OCC, FREE, UNKNOWN = 100, 0, -1


class HdsmCorridorParams(C.Structure):
    _fields_ = [("poly_hor", C.c_int32), ("n_it_decomp", C.c_int32), ("max_rows_per_poly", C.c_int32),
                ("n_traj", C.c_int32), ("max_path", C.c_int32), ("use_cvx_new", C.c_int32), ("voxel_size", C.c_double)]


@dataclass
class CorridorBatch:
    """Inputs of one corridor update for n agents, arrays as hdsm_corridor_batch takes them."""
    poly_hor: int
    n_it: int                  # n_it_decomp
    rmax: int
    voxel: float
    grids: np.ndarray          # [G][dz][dy][dx] int8
    grid_index: Optional[np.ndarray]  # [n] int32 or None (agent i -> grid i)
    dims: np.ndarray           # [n][3] int32 (dx, dy, dz)
    origins: np.ndarray        # [n][3]
    pos: np.ndarray            # [n][3]
    path: np.ndarray           # [n][max_path][3]
    n_path: np.ndarray         # [n] int32
    prev_traj: np.ndarray      # [n][n_traj][3] (n_traj may be 0)
    prev_n: Optional[np.ndarray] = None      # [n] int32; None on the first step
    prev_A: Optional[np.ndarray] = None      # [n][P][rmax][3]
    prev_b: Optional[np.ndarray] = None
    prev_rows: Optional[np.ndarray] = None
    prev_seeds: Optional[np.ndarray] = None  # [n][P][3]
    prev_used: Optional[np.ndarray] = None   # [n][P] uint8
    use_cvx_new: bool = False                # use_cvx_new_ (agent_class.cpp:1383)

    @property
    def n(self):
        return self.pos.shape[0]

    def input_bytes(self):
        arrs = [self.grids, self.grid_index, self.dims, self.origins, self.pos, self.path, self.n_path, self.prev_traj,
                self.prev_n, self.prev_A, self.prev_b, self.prev_rows, self.prev_seeds, self.prev_used]
        return int(sum(a.nbytes for a in arrs if a is not None))

    def with_previous(self, out, poly_used, traj_pos):
        """The next step's batch skeleton fields from this step's outputs (poly_const_vec_, poly_seeds_,
        poly_used_idx_, traj_curr_)."""
        self.prev_n = (out["poly_rows"] > 0).sum(1).astype(np.int32)
        self.prev_A, self.prev_b, self.prev_rows = out["poly_A"].copy(), out["poly_b"].copy(), out["poly_rows"].copy()
        self.prev_seeds = out["seeds"].copy()
        self.prev_used = np.ascontiguousarray(poly_used, np.uint8)
        self.prev_traj = np.ascontiguousarray(traj_pos, np.float64)
        return self


def _p(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


class SafeCorridorGenerator:
    """hdsm_corridor_create / hdsm_corridor_batch / hdsm_corridor_destroy."""

    def __init__(self, poly_hor, n_it_decomp, voxel_size, max_agents, max_grids, grid_stride, n_traj, max_path,
                 rmax=18, device=0, use_cvx_new=False):
        self.L = _lib.load()
        L = self.L
        L.hdsm_corridor_create.restype = C.c_int
        L.hdsm_corridor_batch.restype = C.c_int
        L.hdsm_corridor_batch_device.restype = C.c_int
        L.hdsm_corridor_last_error.restype = C.c_char_p
        L.hdsm_corridor_launch_count.restype = C.c_int64
        L.hdsm_corridor_smem_bytes.restype = C.c_int
        self.prm = HdsmCorridorParams(poly_hor, n_it_decomp, rmax, n_traj, max_path, int(use_cvx_new), voxel_size)
        self.h = C.c_void_p()
        rc = L.hdsm_corridor_create(C.byref(self.prm), C.c_int(max_agents), C.c_int(max_grids), C.c_size_t(grid_stride),
                                    C.c_int(device), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError(f"hdsm_corridor_create failed ({rc}): the corridor generator needs a CUDA device "
                               f"and n_it_decomp <= 90, 18 <= rmax <= 32")
        self.grid_stride = grid_stride

    @property
    def launch_count(self):
        return int(self.L.hdsm_corridor_launch_count(self.h))

    @property
    def smem_bytes(self):
        return int(self.L.hdsm_corridor_smem_bytes(self.h))

    def generate(self, cb: CorridorBatch):
        n, PH, R = cb.n, self.prm.poly_hor, self.prm.max_rows_per_poly
        assert cb.poly_hor == PH and cb.rmax == R and cb.n_it == self.prm.n_it_decomp
        assert cb.path.shape[1] == self.prm.max_path and cb.prev_traj.shape[1] == self.prm.n_traj
        G = cb.grids.shape[0]
        grids = np.ascontiguousarray(cb.grids.reshape(G, -1))
        assert grids.shape[1] == self.grid_stride, (grids.shape, self.grid_stride)
        out = dict(poly_A=np.zeros((n, PH, R, 3)), poly_b=np.zeros((n, PH, R)), poly_rows=np.zeros((n, PH), np.int32),
                   seeds=np.zeros((n, PH, 3)), flags=np.zeros(n, np.int32))
        f8, i4 = np.float64, np.int32
        rc = self.L.hdsm_corridor_batch(
            self.h, C.c_int(n), C.c_int(G), _p(grids, np.int8), _p(cb.grid_index, i4), _p(cb.dims, i4), _p(cb.origins, f8),
            _p(cb.pos, f8), _p(cb.path, f8), _p(cb.n_path, i4), _p(cb.prev_n, i4), _p(cb.prev_A, f8), _p(cb.prev_b, f8),
            _p(cb.prev_rows, i4), _p(cb.prev_seeds, f8), _p(cb.prev_used, np.uint8), _p(cb.prev_traj, f8),
            _p(out["poly_A"], f8), _p(out["poly_b"], f8), _p(out["poly_rows"], i4), _p(out["seeds"], f8), _p(out["flags"], i4))
        if rc != 0:
            raise RuntimeError(f"hdsm_corridor_batch failed ({rc}): {self.L.hdsm_corridor_last_error(self.h).decode()}")
        return out

    def generate_device(self, t, n, stream_ptr=0):
        """hdsm_corridor_batch_device on a DeviceCorridorBatch's tensors `t`; stream ordered, no sync."""
        def dp(k):
            v = t.get(k)
            return None if v is None else C.c_void_p(v.data_ptr())
        rc = self.L.hdsm_corridor_batch_device(
            self.h, C.c_int(n), dp("grids"), dp("grid_index"), dp("dims"), dp("origins"), dp("pos"), dp("path"), dp("n_path"),
            dp("prev_n"), dp("prev_A"), dp("prev_b"), dp("prev_rows"), dp("prev_seeds"), dp("prev_used"), dp("prev_traj"),
            dp("poly_A"), dp("poly_b"), dp("poly_rows"), dp("seeds"), dp("flags"), C.c_void_p(stream_ptr))
        if rc != 0:
            raise RuntimeError(f"hdsm_corridor_batch_device failed ({rc}): {self.L.hdsm_corridor_last_error(self.h).decode()}")

    def close(self):
        if self.h is not None:
            self.L.hdsm_corridor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceCorridorBatch:
    """A CorridorBatch as contiguous torch tensors on `device`, plus the output tensors."""

    IN_KEYS = ("grids", "grid_index", "dims", "origins", "pos", "path", "n_path", "prev_n", "prev_A", "prev_b",
               "prev_rows", "prev_seeds", "prev_used", "prev_traj")

    def __init__(self, cb: CorridorBatch, device):
        import torch
        self.n = cb.n
        self.t = {}
        for k in self.IN_KEYS:
            a = getattr(cb, k)
            if a is None or (k == "prev_traj" and cb.prev_n is None):
                continue
            a = np.ascontiguousarray(a.reshape(a.shape[0], -1) if k == "grids" else a)
            self.t[k] = torch.from_numpy(a).to(device)
        n, PH, R = cb.n, cb.poly_hor, cb.rmax
        f64 = torch.float64
        self.t["poly_A"] = torch.zeros((n, PH, R, 3), dtype=f64, device=device)
        self.t["poly_b"] = torch.zeros((n, PH, R), dtype=f64, device=device)
        self.t["poly_rows"] = torch.zeros((n, PH), dtype=torch.int32, device=device)
        self.t["seeds"] = torch.zeros((n, PH, 3), dtype=f64, device=device)
        self.t["flags"] = torch.zeros(n, dtype=torch.int32, device=device)


def corridor_algorithmic_bytes(cb: CorridorBatch, poly_rows) -> float:
    """Compulsory HBM traffic of one corridor update, bytes per agent: the occupancy rows a polytope can
    look at ((2g + 3)^2 rows of 32 voxels around its seed, g = ceil(n_it / 6)), the path, the kept
    polytopes read once, and every output written once."""
    g = (cb.n_it + 5) // 6
    new_polys = (np.asarray(poly_rows) > 0).sum(1) - (0 if cb.prev_n is None else 0)
    window = (2 * g + 3) ** 2 * 32
    out_bytes = cb.poly_hor * (cb.rmax * 32 + 4 + 24) + 4
    return float(np.mean(new_polys * window + cb.n_path * 24 + 24 + 36 + out_bytes))


# --------------------------------------------------------------------------------------
# synthetic local voxel grids and paths (stand-ins for mapping_util / path_finding_util)
# --------------------------------------------------------------------------------------
def local_grid(world: Forest, centre, voxel=0.3, grid_range=(20.0, 20.0, 6.0), inflation=0.3, col_half=0.05,
               unknown_outside=True):
    """int8 grid [dz][dy][dx] around `centre` and its origin: columns of the forest voxelised and inflated
    by ceil(inflation / voxel) voxels in x and y (map_builder.cpp:336), space below z = 0 unknown."""
    centre = np.asarray(centre, float)
    rng3 = np.asarray(grid_range, float)
    origin = np.round((centre - rng3 / 2) / voxel) * voxel           # map_builder.cpp:95-104 with a zero map origin
    dim = np.floor(rng3 / voxel).astype(int)                          # :108-110
    g = np.zeros((dim[2], dim[1], dim[0]), np.int8)
    k = int(np.ceil(inflation / voxel))
    cols = world.near(centre, rng3[0] / 2 + 1.0) if len(world.cols) else world.cols
    for c in cols:
        x0 = int(np.floor((c[0] - col_half - origin[0]) / voxel)) - k
        x1 = int(np.floor((c[0] + col_half - origin[0]) / voxel)) + k
        y0 = int(np.floor((c[1] - col_half - origin[1]) / voxel)) - k
        y1 = int(np.floor((c[1] + col_half - origin[1]) / voxel)) + k
        if x1 < 0 or y1 < 0 or x0 >= dim[0] or y0 >= dim[1]:
            continue
        g[:, max(y0, 0):y1 + 1, max(x0, 0):x1 + 1] = OCC
    if unknown_outside:
        zk = int(np.ceil((0.0 - origin[2]) / voxel - 1e-9))           # voxels entirely below the ground plane
        if zk > 0:
            g[:min(zk, dim[2])] = UNKNOWN
    return g, origin


def clipped_path(pos, goal, world: Forest, reach=8.5):
    """path_curr_ stand-in: the synthetic planner's way-points (current position excluded), cut where the
    path leaves a ball of `reach` metres so that every seed stays inside the local grid."""
    pts = plan_path(pos, goal, world)
    out = []
    pos = np.asarray(pos, float)
    for a, b in zip(pts[:-1], pts[1:]):
        if np.linalg.norm(b - pos) <= reach:
            out.append(b.copy())
            continue
        d = b - a
        # largest t in [0, 1] with |a + t d - pos| = reach
        aa, bb, cc = d @ d, 2 * d @ (a - pos), (a - pos) @ (a - pos) - reach * reach
        disc = max(bb * bb - 4 * aa * cc, 0.0)
        t = min(1.0, max(0.0, (-bb + np.sqrt(disc)) / (2 * aa))) if aa > 0 else 0.0
        out.append(a + t * d)
        break
    return np.array(out) if out else pos[None, :].copy()


def corridor_batch(sw, poly_hor=None, n_it=42, rmax=18, voxel=0.3, max_path=16, ids=None, shared_grids=False):
    """CorridorBatch for the agents of a scenarios.Swarm (first step: no previous polytopes)."""
    P = poly_hor or sw.params["poly_hor"]
    ids = np.arange(sw.n) if ids is None else np.asarray(ids)
    n = len(ids)
    grids, dims, origins = [], np.zeros((n, 3), np.int32), np.zeros((n, 3))
    path = np.zeros((n, max_path, 3))
    n_path = np.zeros(n, np.int32)
    for r, i in enumerate(ids):
        pos = sw.state[i, :3]
        g, o = local_grid(sw.world, pos, voxel)
        grids.append(g)
        dims[r] = (g.shape[2], g.shape[1], g.shape[0])
        origins[r] = o
        pts = clipped_path(pos, sw.goal[i], sw.world)[:max_path]
        path[r, :len(pts)] = pts
        n_path[r] = len(pts)
    N = sw.params["n_hor"]
    prev_traj = np.zeros((n, N + 1, 3))
    return CorridorBatch(P, n_it, rmax, voxel, np.stack(grids), None, dims, origins,
                         np.ascontiguousarray(sw.state[ids, :3]), path, n_path, prev_traj)


class CorridorLoop:
    """Closed loop of the two replaced calls for a scenarios.Swarm (agent_class.cpp:165-174): every step the
    corridor generator gets the previous step's polytopes, seeds, `poly_used_idx_` and plan, and its rows replace
    the synthetic polytopes of `Swarm.make_batch`.  `generate(cb)` is SafeCorridorGenerator.generate or a checker
    with the same signature."""

    def __init__(self, sw, n_it=42, max_path=16):
        self.sw, self.n_it, self.max_path = sw, n_it, max_path
        self.prev_out = None
        self.prev_used = None

    def corridor_inputs(self) -> CorridorBatch:
        cb = corridor_batch(self.sw, n_it=self.n_it, max_path=self.max_path)
        if self.prev_out is not None:
            traj = self.sw.traj[:, :, :3] if self.sw.traj is not None else cb.prev_traj
            cb.with_previous(self.prev_out, self.prev_used, np.ascontiguousarray(traj))
        return cb

    def solver_inputs(self, out):
        b = self.sw.make_batch()
        b.poly_A, b.poly_b, b.poly_rows = out["poly_A"].copy(), out["poly_b"].copy(), out["poly_rows"].copy()
        return b

    def advance(self, out, solved, ok):
        """`solved`: dict(traj, ctrl, poly_used) of this step; failed agents keep poly_used_idx_ (:997-1019)."""
        used = np.ascontiguousarray(solved["poly_used"], np.uint8).copy()
        if self.prev_used is not None:
            used[~ok] = self.prev_used[~ok]
        self.prev_out, self.prev_used = out, used
        self.sw.advance(solved["traj"], solved["ctrl"], ok)
