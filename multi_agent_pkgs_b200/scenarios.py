"""Synthetic swarm scenarios: the inputs of the per-agent trajectory optimisation.

The reference produces these inputs with ROS2 nodes (voxel map -> safe corridor ->
reference trajectory); none of that is in scope (SURVEY.md section 8(f) lists it as "next").
This module fabricates inputs of the same *shape and statistics* on the CPU with NumPy so
that the hot path can be exercised, tested and benchmarked stand-alone:

* polytopes: <= 12 chamfer planes with small-integer normals followed by 6 axis faces,
  b = p . n  (convex_decomp_util/src/convex_decomp.cpp:335-375, agent_class.cpp:1428-1437);
* reference trajectory: N+1 samples along the path at path_vel*dt spacing with the
  reference's (backwards-pointing) velocity convention (agent_class.cpp:1591-1663, :1520-1546);
* scenario geometry: circle swap (multi_agent_planner_circle.launch.py:25-44) over a forest
  of 0.1 m columns on whole-metre offsets (env_default_config.yaml:3-13,
  environment_builder.cpp:205-215), and the larger synthetic swarms of BASELINE.json.

Array layouts are the C-ABI layouts of include/hdsm.h (row-major, contiguous doubles).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

VOXEL = 0.3
GROW = 7 * VOXEL            # n_it_decomp=42 grows 7 voxels per face in free space
KEEP_OUT = 0.45             # column half-width 0.05 + 0.3 inflation, snapped up to the voxel grid


def agile_params(n_hor: int = 10) -> Dict:
    """multi_agent_planner/config/agent_agile_config.yaml with n_hor overridden (BASELINE.json)."""
    return dict(n_hor=n_hor, poly_hor=4, dt=0.1, rk4=False, drag=(0.0, 0.0, 0.0), r_u=0.01,
                r_x=(100.0, 100.0, 100.0, 1.0, 1.0, 1.0), r_n=(100.0, 100.0, 100.0, 1.0, 1.0, 1.0),
                max_vel=20.0, min_acc_xy=-15.0, max_acc_xy=15.0, min_acc_z=-15.0, max_acc_z=15.0,
                max_jerk=60.0, drone_radius=0.25, drone_z_offset=0.25, tilt=0.1)


def default_params(n_hor: int = 9) -> Dict:
    """agent_default_config.yaml (circle-empty and window scenarios)."""
    d = agile_params(n_hor)
    d.update(poly_hor=3, drone_radius=0.125, drone_z_offset=0.125, max_vel=9.5, min_acc_xy=-20.0,
             max_acc_xy=20.0, min_acc_z=-20.0, max_acc_z=20.0, max_jerk=30.0)
    return d


def crazyflie_params(n_hor: int = 9) -> Dict:
    """agent_crazyflie_config.yaml (hardware)."""
    d = agile_params(n_hor)
    d.update(poly_hor=3, drone_radius=0.3, drone_z_offset=0.6, max_vel=1.5, min_acc_xy=-3.0,
             max_acc_xy=3.0, min_acc_z=-3.0, max_acc_z=3.0, max_jerk=4.0)
    return d


# --------------------------------------------------------------------------------------
# world
# --------------------------------------------------------------------------------------
@dataclass
class Forest:
    cols: np.ndarray  # (n, 2) column centres; empty array = empty map

    @staticmethod
    def empty():
        return Forest(np.zeros((0, 2)))

    @staticmethod
    def reference_grid(rng, n_obst=180, origin=(6.5, 6.5), extent=(30.0, 30.0)):
        """Whole-metre offsets like the reference's integer division (environment_builder.cpp:210-215)."""
        xy = np.stack([rng.integers(0, int(extent[0]) + 1, n_obst), rng.integers(0, int(extent[1]) + 1, n_obst)], 1)
        return Forest(xy.astype(float) + np.asarray(origin))

    @staticmethod
    def density(rng, lo, hi, per_m2=0.2):
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        n = int(per_m2 * float(np.prod(hi - lo)))
        return Forest(lo + rng.random((n, 2)) * (hi - lo))

    _CELL = 4.0  # bucket size of the lazily built index (metres)

    def _index(self):
        """Uniform-grid buckets over the columns (built on first use; large forests only): `near` then looks at a
        few buckets instead of every column.  Candidates are visited in ascending column order, so results are
        identical to the plain scan."""
        ix = self.__dict__.get("_ix")
        if ix is None:
            key = np.floor(self.cols / self._CELL).astype(np.int64)
            ix = {}
            for i, (a, b) in enumerate(key):
                ix.setdefault((int(a), int(b)), []).append(i)
            ix = {k: np.asarray(v) for k, v in ix.items()}
            self.__dict__["_ix"] = ix
        return ix

    def _candidates(self, xy, rad):
        if len(self.cols) < 512:
            return self.cols
        ix = self._index()
        x0, x1 = int(math.floor((xy[0] - rad) / self._CELL)), int(math.floor((xy[0] + rad) / self._CELL))
        y0, y1 = int(math.floor((xy[1] - rad) / self._CELL)), int(math.floor((xy[1] + rad) / self._CELL))
        got = [ix[(a, b)] for a in range(x0, x1 + 1) for b in range(y0, y1 + 1) if (a, b) in ix]
        if not got:
            return self.cols[:0]
        return self.cols[np.sort(np.concatenate(got))]

    def near(self, xy, rad):
        if len(self.cols) == 0:
            return self.cols
        c = self._candidates(xy, rad)
        d = np.abs(c - np.asarray(xy)[None, :2])
        return c[(d[:, 0] < rad) & (d[:, 1] < rad)]

    def is_free(self, xy, margin=0.0):
        if len(self.cols) == 0:
            return True
        c = self._candidates(xy, KEEP_OUT + margin)
        d = np.abs(c - np.asarray(xy)[None, :2])
        return not np.any((d[:, 0] < KEEP_OUT + margin) & (d[:, 1] < KEEP_OUT + margin))

    def push_free(self, xy, margin=0.1):
        """Move a point out of any keep-out square (shortest axis push)."""
        xy = np.array(xy[:2], float)
        for _ in range(8):
            if self.is_free(xy, margin):
                break
            for c in self.near(xy, KEEP_OUT + margin):
                d = xy - c
                if abs(d[0]) < KEEP_OUT + margin and abs(d[1]) < KEEP_OUT + margin:
                    ax = 0 if abs(d[0]) >= abs(d[1]) else 1
                    sgn = 1.0 if d[ax] >= 0 else -1.0
                    xy[ax] = c[ax] + sgn * (KEEP_OUT + margin + 1e-3)
        return xy


# --------------------------------------------------------------------------------------
# polytopes
# --------------------------------------------------------------------------------------
def make_polytope(seed, world: Forest, rng, max_rows=18, z_floor=0.0, extra_chamfers=True):
    """Obstacle-free convex cell around `seed`: rows [chamfers..., 6 faces], A x <= b."""
    s = np.asarray(seed, float)
    lo = np.array([s[0] - GROW, s[1] - GROW, max(z_floor, s[2] - GROW)])
    hi = np.array([s[0] + GROW, s[1] + GROW, s[2] + GROW])
    cham: List[Tuple[np.ndarray, float]] = []
    cols = world.near(s, GROW + KEEP_OUT)
    if len(cols):
        order = np.argsort(np.max(np.abs(cols - s[None, :2]), axis=1))
        for c in cols[order]:
            if not (lo[0] - KEEP_OUT < c[0] < hi[0] + KEEP_OUT and lo[1] - KEEP_OUT < c[1] < hi[1] + KEEP_OUT):
                continue
            if any(n[:2] @ (c - np.sign(n[:2]) * KEEP_OUT) >= b - 1e-9 for n, b in cham):
                continue  # already cut away by a chamfer
            d = c - s[:2]
            sg = np.where(d >= 0, 1.0, -1.0)
            ad = np.abs(d)
            slope = 1.0
            if ad[0] > 2 * ad[1]:
                nd = np.array([2.0 * sg[0], sg[1], 0.0])   # slope-2 chamfer, like (s,1,0) in the reference
            elif ad[1] > 2 * ad[0]:
                nd = np.array([sg[0], 2.0 * sg[1], 0.0])
            else:
                nd = np.array([sg[0], sg[1], 0.0])
            corner = np.array([c[0] - sg[0] * KEEP_OUT, c[1] - sg[1] * KEEP_OUT, 0.0])
            dist_diag = (nd @ corner - nd @ np.array([s[0], s[1], 0.0])) / np.linalg.norm(nd)
            dist_ax = ad - KEEP_OUT
            best_ax = int(np.argmax(dist_ax))
            if dist_diag > dist_ax[best_ax] and dist_diag > 0.15 and len(cham) < 8:
                cham.append((nd, float(nd @ corner)))
            else:
                if sg[best_ax] > 0:
                    hi[best_ax] = min(hi[best_ax], c[best_ax] - KEEP_OUT)
                else:
                    lo[best_ax] = max(lo[best_ax], c[best_ax] + KEEP_OUT)
    if extra_chamfers:
        # random x-z / y-z corner chamfers (the reference has 12 possible edges, 4 per axis pair)
        for ax in (0, 1):
            for sx in (-1.0, 1.0):
                for sz in (-1.0, 1.0):
                    if rng.random() < 0.35 and len(cham) < max_rows - 6:
                        n = np.zeros(3)
                        n[ax] = sx
                        n[2] = sz
                        cx = hi[ax] if sx > 0 else lo[ax]
                        cz = hi[2] if sz > 0 else lo[2]
                        cut = VOXEL * int(rng.integers(1, 4))
                        b = sx * cx + sz * cz - cut
                        if n @ s < b - 0.3:
                            cham.append((n, float(b)))
    A = [n for n, _ in cham]
    b = [bb for _, bb in cham]
    for ax in range(3):  # faces: +x, -x, +y, -y, +z, -z
        e = np.zeros(3)
        e[ax] = 1.0
        A.append(e.copy())
        b.append(hi[ax])
        A.append(-e)
        b.append(-lo[ax])
    return np.array(A), np.array(b)


def inside(poly, pt, tol=0.0):
    A, b = poly
    return bool(np.all(A @ np.asarray(pt, float) - b <= tol))


def plan_path(pos, goal, world: Forest, lookahead=14.0):
    """Piecewise-linear path from pos toward goal with side-steps around columns."""
    pos = np.asarray(pos, float)
    goal = np.asarray(goal, float)
    pts = [pos.copy()]
    cur = pos.copy()
    for _ in range(12):
        d = goal - cur
        L = float(np.linalg.norm(d))
        if L < 1e-6 or np.linalg.norm(cur - pos) > lookahead:
            break
        t = d / L
        hit = None
        best = min(L, lookahead)
        for c in world.near(cur + t * best / 2, best / 2 + 1.0):
            rel = c - cur[:2]
            along = rel @ t[:2]
            if along <= 0.05 or along > best:
                continue
            perp = rel - along * t[:2]
            if np.max(np.abs(perp)) < KEEP_OUT + 0.2 and along < best:
                best, hit = along, (c, perp)
        if hit is None:
            break
        c, perp = hit
        nrm = np.array([-t[1], t[0]])
        side = -1.0 if perp @ nrm > 0 else 1.0
        wp = np.array([*(c + side * nrm * (KEEP_OUT * math.sqrt(2) + 0.45)), cur[2] + (goal[2] - cur[2]) * best / L])
        wp[:2] = world.push_free(wp[:2], 0.25)
        if np.linalg.norm(wp - cur) < 0.05:
            break
        pts.append(wp)
        cur = wp
    pts.append(goal.copy())
    return np.array(pts)


def sample_path(path, path_vel, n_hor, dt):
    """SamplePath (agent_class.cpp:1591-1663, path_vel_dec = 0) + the velocity reference (:1520-1546).
    Returns (n_hor+1, 6); the optimisation reads rows 0..n_hor-1."""
    samp = path_vel * dt
    pts = [path[0].copy()]
    cur = path[0].copy()
    idx = 1
    limit = samp
    while len(pts) < n_hor + 1:
        nxt = path[idx]
        diff = nxt - cur
        dist = float(np.linalg.norm(diff))
        if dist > limit:
            cur = cur + limit * diff / dist
            pts.append(cur.copy())
            limit = samp
        else:
            cur = nxt.copy()
            idx += 1
            if idx == len(path):
                while len(pts) < n_hor + 1:
                    pts.append(path[-1].copy())
                break
            limit -= dist
    pts = np.array(pts)
    ref = np.zeros((n_hor + 1, 6))
    ref[:, :3] = pts
    v = np.zeros(3)
    for i in range(n_hor):
        dist = float(np.linalg.norm(pts[i] - pts[i + 1]))
        v = path_vel * (pts[i] - pts[i + 1]) / dist if dist > 1e-2 else np.zeros(3)
        ref[i, 3:] = v
    ref[n_hor, 3:] = v
    return ref


def corridor(pos, path, world: Forest, rng, poly_hor, max_rows=18, extra_chamfers=True):
    """Up to poly_hor overlapping polytopes seeded along the path (agent_class.cpp:1236-1447 in spirit)."""
    polys = [make_polytope(pos, world, rng, max_rows, extra_chamfers=extra_chamfers)]
    step = VOXEL / 3
    last_inside = np.asarray(pos, float).copy()
    for a, b in zip(path[:-1], path[1:]):
        L = float(np.linalg.norm(b - a))
        n = max(1, int(L / step))
        for i in range(1, n + 1):
            pt = a + (b - a) * (i / n)
            if inside(polys[-1], pt, -0.05):
                last_inside = pt
            elif len(polys) < poly_hor:
                seed = last_inside.copy()
                seed[:2] = world.push_free(seed[:2], 0.12)
                polys.append(make_polytope(seed, world, rng, max_rows, extra_chamfers=extra_chamfers))
                if inside(polys[-1], pt, -0.05):
                    last_inside = pt
            else:
                return polys
    return polys


# --------------------------------------------------------------------------------------
# batch container in C-ABI layout
# --------------------------------------------------------------------------------------
@dataclass
class Batch:
    """One replanning step's inputs for n agents, arrays as hdsm_solve_batch takes them."""
    params: Dict
    global_id: np.ndarray      # [n] int32   index into all_pos
    nbr_begin: np.ndarray      # [n] int32   neighbour candidates = all_pos[nbr_begin:nbr_end]
    nbr_end: np.ndarray        # [n] int32
    x0: np.ndarray             # [n][9]
    ref: np.ndarray            # [n][N][6]
    poly_A: np.ndarray         # [n][P][Rmax][3]
    poly_b: np.ndarray         # [n][P][Rmax]
    poly_rows: np.ndarray      # [n][P] int32, 0 = absent
    prev_self_pos: np.ndarray  # [n][N+1][3]
    all_pos: np.ndarray        # [n_rob][N+1][3]
    all_valid: np.ndarray      # [n_rob] uint8
    rmax: int = 18

    @property
    def n(self):
        return self.x0.shape[0]

    def polys_of(self, i):
        return [(self.poly_A[i, p, :r].copy(), self.poly_b[i, p, :r].copy())
                for p, r in enumerate(self.poly_rows[i]) if r > 0]

    def take(self, idx):
        idx = np.asarray(idx)
        return Batch(self.params, self.global_id[idx], self.nbr_begin[idx], self.nbr_end[idx], self.x0[idx],
                     self.ref[idx], self.poly_A[idx], self.poly_b[idx], self.poly_rows[idx],
                     self.prev_self_pos[idx], self.all_pos, self.all_valid, self.rmax)

    def save(self, path):
        np.savez_compressed(path, params=np.array(repr(self.params)), global_id=self.global_id,
                            nbr_begin=self.nbr_begin, nbr_end=self.nbr_end, x0=self.x0, ref=self.ref,
                            poly_A=self.poly_A, poly_b=self.poly_b, poly_rows=self.poly_rows,
                            prev_self_pos=self.prev_self_pos, all_pos=self.all_pos, all_valid=self.all_valid,
                            rmax=self.rmax)

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        params = eval(str(z["params"]), {"__builtins__": {}})  # repr of a dict of numbers/tuples/bools
        return Batch(params, z["global_id"], z["nbr_begin"], z["nbr_end"], z["x0"], z["ref"], z["poly_A"],
                     z["poly_b"], z["poly_rows"], z["prev_self_pos"], z["all_pos"], z["all_valid"], int(z["rmax"]))


def pack_polys(polys_per_agent, P, rmax):
    n = len(polys_per_agent)
    A = np.zeros((n, P, rmax, 3))
    b = np.zeros((n, P, rmax))
    rows = np.zeros((n, P), np.int32)
    for i, polys in enumerate(polys_per_agent):
        for p, (Ap, bp) in enumerate(polys[:P]):
            r = len(bp)
            assert r <= rmax
            A[i, p, :r] = Ap
            b[i, p, :r] = bp
            rows[i, p] = r
    return A, b, rows


# --------------------------------------------------------------------------------------
# swarms
# --------------------------------------------------------------------------------------
@dataclass
class Swarm:
    """State of a (set of) simulated swarm(s) between replanning steps (closed loop, perfect
    tracking: state_curr_ := traj_curr_[step_plan], agent_class.cpp:233-238)."""
    params: Dict
    world: Forest
    state: np.ndarray          # [n][9]
    goal: np.ndarray           # [n][3]
    path_vel: np.ndarray       # [n]
    group_begin: np.ndarray    # [n] first agent of this agent's swarm instance
    group_end: np.ndarray      # [n]
    traj: Optional[np.ndarray] = None    # [n][N+1][9] last plans (None before the first solve)
    ctrl: Optional[np.ndarray] = None
    have_plan: Optional[np.ndarray] = None
    rng: Optional[np.random.Generator] = None
    rmax: int = 18
    extra_chamfers: bool = True
    seed: int = 0
    step_count: int = 0

    @property
    def n(self):
        return self.state.shape[0]

    def make_batch(self, ids=None) -> Batch:
        """Inputs of the next replanning step for agents `ids` (default all)."""
        P, N = self.params["poly_hor"], self.params["n_hor"]
        n = self.n
        ids = np.arange(n) if ids is None else np.asarray(ids)
        if self.have_plan is None:
            self.have_plan = np.zeros(n, np.uint8)
        all_pos = np.zeros((n, N + 1, 3))
        if self.traj is not None:
            all_pos[:] = self.traj[:, :, :3]
        x0 = self.state[ids].copy()
        prev = np.zeros((len(ids), N + 1, 3))
        for r, i in enumerate(ids):
            if self.have_plan[i]:
                prev[r] = self.traj[i, :, :3]
            else:
                prev[r] = self.state[i, None, :3]  # state_ini_ before the first solve (agent_class.cpp:1108-1110)
        refs, A, b, rows = self.make_inputs(ids)
        return Batch(self.params, ids.astype(np.int32), self.group_begin[ids].astype(np.int32),
                     self.group_end[ids].astype(np.int32), x0, refs, A, b, rows, prev, all_pos,
                     self.have_plan.copy(), self.rmax)

    def make_batch_pooled(self, pool) -> "Batch":
        """make_batch() of the whole swarm with the input producers spread over an InputPool."""
        b = self.make_batch(np.zeros(0, np.int64))
        ids = np.arange(self.n)
        ref, A, bb, rows = self.make_inputs(ids, pool=pool)
        N = self.params["n_hor"]
        prev = np.repeat(self.state[:, None, :3], N + 1, axis=1)  # state_ini_ before the first solve (:1108-1110)
        if self.traj is not None:
            have = self.have_plan != 0
            prev[have] = self.traj[have][:, :, :3]
        return Batch(self.params, ids.astype(np.int32), self.group_begin.astype(np.int32), self.group_end.astype(np.int32),
                     self.state.copy(), ref, A, bb, rows, np.ascontiguousarray(prev), b.all_pos, b.all_valid, self.rmax)

    def agent_inputs(self, i, pos, step):
        """Reference trajectory and corridor cells of agent i standing at `pos` in replanning step `step` - the
        stand-ins for GenerateReferenceTrajectory / GenerateSafeCorridor.  Random choices come from a per-agent,
        per-step stream, so a shard of the swarm generates the same inputs as the whole."""
        P, N = self.params["poly_hor"], self.params["n_hor"]
        path = plan_path(pos, self.goal[i], self.world)
        ref = sample_path(path, self.path_vel[i], N, self.params["dt"])[:N]
        rng_i = np.random.default_rng([self.seed, int(step), int(i)])
        return ref, corridor(pos, path, self.world, rng_i, P, self.rmax, self.extra_chamfers)

    def make_inputs(self, ids=None, pos=None, step=None, pool=None):
        """(ref [n][N][6], poly_A, poly_b, poly_rows) for agents `ids` standing at `pos` (default: their current
        states) in step `step` (default: the current one).  `pool`: an InputPool over this swarm."""
        P, N = self.params["poly_hor"], self.params["n_hor"]
        ids = np.arange(self.n) if ids is None else np.asarray(ids)
        pos = self.state[ids, :3] if pos is None else np.asarray(pos, float)
        step = self.step_count if step is None else int(step)
        if pool is not None:
            res = pool.map(step, ids, pos)
        else:
            res = [self.agent_inputs(int(i), pos[r], step) for r, i in enumerate(ids)]
        refs = np.zeros((len(ids), N, 6))
        for r, (ref, _) in enumerate(res):
            refs[r] = ref
        A, b, rows = pack_polys([p for _, p in res], P, self.rmax)
        return refs, A, b, rows

    def advance(self, traj, ctrl, ok, ids=None):
        """Apply one step's results: failed agents shift their previous plan (agent_class.cpp:1000-1019)."""
        n = self.n
        N = self.params["n_hor"]
        ids = np.arange(n) if ids is None else np.asarray(ids)
        self.step_count += 1
        if self.traj is None:
            self.traj = np.zeros((n, N + 1, 9))
            self.traj[:] = self.state[:, None, :]
            self.ctrl = np.zeros((n, N, 3))
        for r, i in enumerate(ids):
            if ok[r]:
                self.traj[i] = traj[r]
                self.ctrl[i] = ctrl[r]
                self.have_plan[i] = 1
            elif self.have_plan[i]:
                self.traj[i] = np.concatenate([self.traj[i, 1:], self.traj[i, -1:]])
                self.ctrl[i] = np.concatenate([self.ctrl[i, 1:], self.ctrl[i, -1:]])
            if self.have_plan[i]:
                self.state[i] = self.traj[i, 1]


def _mk_swarm(params, world, starts, goals, groups, rng, extra_chamfers=True, seed=0):
    n = len(starts)
    state = np.zeros((n, 9))
    state[:, :3] = starts
    for i in range(n):  # a start inside an (inflated) column has no valid corridor
        state[i, :2] = world.push_free(state[i, :2], 0.2)
    vel = rng.uniform(4.5, 9.0, n)  # path_vel_min/max, agent_agile_config.yaml:16-17
    gb = np.zeros(n, np.int32)
    ge = np.zeros(n, np.int32)
    for (a, b) in groups:
        gb[a:b] = a
        ge[a:b] = b
    return Swarm(params, world, state, np.asarray(goals, float), vel, gb, ge, rng=rng, extra_chamfers=extra_chamfers,
                 seed=seed)


def config1_single_agent(seed=1, n_hor=10):
    """1 agent, empty known map (BASELINE.json configs[0]; agent_agile_config.yaml:42-43)."""
    rng = np.random.default_rng(seed)
    params = agile_params(n_hor)
    return _mk_swarm(params, Forest.empty(), np.array([[0.0, 0.0, 1.5]]), np.array([[42.15, 42.15, 1.5]]),
                     [(0, 1)], rng, extra_chamfers=False, seed=seed)


def config2_circle(seed=2, n_swarms=1, n_rob=10, n_hor=10, radius=22.0, centre=(18.0, 15.0)):
    """n_swarms independent copies of the 10-agent circle swap over the forest (configs[1]).
    Each copy gets its own forest realisation (shifted) and start-angle jitter."""
    rng = np.random.default_rng(seed)
    params = agile_params(n_hor)
    starts, goals, groups, cols = [], [], [], []
    for s in range(n_swarms):
        off = np.array([200.0 * s, 0.0])  # swarms live in disjoint strips of one world
        f = Forest.reference_grid(rng)
        cols.append(f.cols + off)
        phase = rng.uniform(0, 2 * math.pi) if s > 0 else 0.0
        pts = []
        for i in range(n_rob):
            ang = phase + 2 * math.pi * i / n_rob
            pts.append([centre[0] + off[0] + radius * math.cos(ang), centre[1] + off[1] + radius * math.sin(ang), 1.5])
        pts = np.array(pts)
        starts.append(pts)
        goals.append(pts[(np.arange(n_rob) + n_rob // 2) % n_rob])
        groups.append((s * n_rob, (s + 1) * n_rob))
    world = Forest(np.concatenate(cols))
    return _mk_swarm(params, world, np.concatenate(starts), np.concatenate(goals), groups, rng, seed=seed)


def config3_line(seed=3, n_rob=10, n_hor=10):
    """10 agents in a line crossing forest + wall-with-gaps + forest (configs[2];
    multi_agent_planner_long.launch.py:24-42, generate_random_grid.py:74-115)."""
    rng = np.random.default_rng(seed)
    params = agile_params(n_hor)
    f1 = rng.uniform([5, 0], [40, 30], (90, 2))
    wall_y = np.arange(0.0, 30.0, 0.3)
    gaps = np.linspace(1.0, 29.0, 15)
    wall_y = wall_y[np.min(np.abs(wall_y[:, None] - gaps[None, :]), axis=1) > 0.75]
    wall = np.stack([np.full_like(wall_y, 48.0), wall_y], 1)
    f2 = rng.uniform([55, 0], [90, 30], (180, 2))
    world = Forest(np.concatenate([f1, wall, f2]))
    starts = np.array([[0.0, 5 + 2.01 * i, 1.0] for i in range(n_rob)])
    goals = starts + np.array([96.01, 0.0, 0.0])
    return _mk_swarm(params, world, starts, goals, [(0, n_rob)], rng, seed=seed)


def config4_circle256(seed=4, n_rob=256, n_hor=10, radius=60.0):
    """256 agents on a 60 m circle, antipodal goals, forest density 0.2 /m^2 (configs[3])."""
    rng = np.random.default_rng(seed)
    params = agile_params(n_hor)
    world = Forest.density(rng, (-45.0, -45.0), (45.0, 45.0), 0.2)
    ang = 2 * math.pi * np.arange(n_rob) / n_rob
    starts = np.stack([radius * np.cos(ang), radius * np.sin(ang), np.full(n_rob, 1.5)], 1)
    goals = starts[(np.arange(n_rob) + n_rob // 2) % n_rob]
    return _mk_swarm(params, world, starts, goals, [(0, n_rob)], rng, seed=seed)


def config5_random(seed=5, n_rob=4096, n_hor=10, side=200.0):
    """n_rob agents, dart-throwing starts (>= 1 m apart) in side x side x [1,3] m, goals >= 50 m
    away, forest density 0.2 /m^2 (configs[4])."""
    rng = np.random.default_rng(seed)
    params = agile_params(n_hor)
    world = Forest.density(rng, (0.0, 0.0), (side, side), 0.2)
    cell = {}
    starts = []
    while len(starts) < n_rob:
        pt = np.array([*rng.uniform(0, side, 2), rng.uniform(1.0, 3.0)])
        key = (int(pt[0]), int(pt[1]))
        ok = world.is_free(pt, 0.2)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for q in cell.get((key[0] + dx, key[1] + dy), []):
                    if np.linalg.norm(q[:2] - pt[:2]) < 1.0:
                        ok = False
        if ok:
            cell.setdefault(key, []).append(pt)
            starts.append(pt)
    starts = np.array(starts)
    goals = np.zeros_like(starts)
    for i in range(n_rob):
        while True:
            g = np.array([*rng.uniform(0, side, 2), rng.uniform(1.0, 3.0)])
            if np.linalg.norm(g[:2] - starts[i, :2]) >= min(50.0, side / 2):
                goals[i] = g
                break
    return _mk_swarm(params, world, starts, goals, [(0, n_rob)], rng, seed=seed)


# --------------------------------------------------------------------------------------
# process pool for the host-side input producers (pure Python: ~5 ms per agent and step)
# --------------------------------------------------------------------------------------
_POOL_SWARM = None


def _pool_chunk(args):
    step, ids, pos = args
    return [_POOL_SWARM.agent_inputs(int(i), pos[r], step) for r, i in enumerate(ids)]


class InputPool:
    """Fork pool over one Swarm's static data (world, goals, speeds, seed).  Create it BEFORE the process
    initialises CUDA: the workers are forked copies and must never touch the GPU."""

    def __init__(self, swarm: Swarm, procs: Optional[int] = None):
        import multiprocessing as mp
        import os
        global _POOL_SWARM
        self.procs = max(1, int(procs or (os.cpu_count() or 1)))
        _POOL_SWARM = swarm
        self.pool = mp.get_context("fork").Pool(self.procs) if self.procs > 1 else None
        self.swarm = swarm

    def map(self, step, ids, pos):
        ids, pos = np.asarray(ids), np.asarray(pos, float)
        if self.pool is None or len(ids) < 4 * self.procs:
            return [self.swarm.agent_inputs(int(i), pos[r], step) for r, i in enumerate(ids)]
        cuts = np.linspace(0, len(ids), 4 * self.procs + 1).astype(int)
        parts = self.pool.map(_pool_chunk, [(step, ids[a:b], pos[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a])
        return [x for part in parts for x in part]

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool.join()
            self.pool = None
