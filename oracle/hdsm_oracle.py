"""CPU oracle for the HDSM per-agent trajectory optimisation (TEST INFRASTRUCTURE ONLY).

This file is a float64 NumPy restatement of the one hot path of
lis-epfl/multi_agent_pkgs that the CUDA library replaces.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import it.  Nothing under ``multi_agent_pkgs_b200/`` imports or calls it.

Reference functions restated here (paths relative to the reference checkout):

* bounds                       multi_agent_planner/src/agent_class.cpp:2169-2188
* variables, terminal, dynamics  agent_class.cpp:2071-2153, ModelODE :2155-2167
* objective, x0, corridor rows  agent_class.cpp:858-941, :1071-1084
* inter-agent planes            agent_class.cpp:1086-1215, AddHyperplane :1217-1234
* row semantics  A x <= b       decomp_ros/decomp_util/include/decomp_geometry/polyhedron.h:98-147
* failure fallback              agent_class.cpp:997-1019

PARITY PIN STATUS.  The arithmetic of the reference lives in Gurobi 10.0.x
(closed source, licence-gated; linked at multi_agent_planner/CMakeLists.txt:46) which
is not available, and the reference ships no tests or golden vectors for this
path, so **parity against Gurobi's SOLVE is unpinned**.
The problem DATA is pinned: the reference's own agent_class.cpp, compiled
unmodified on stand-in ROS / Eigen headers and a recording stand-in for the
Gurobi C++ API (oracle/_ref/libref_agent.so, oracle/ref_wrap_agent.cpp), builds
its model on the same inputs, and ``time_aware_planes`` / ``build_qp_full`` /
``Params`` bounds reproduce its planes, objective, bounds, dynamics rows, one-hot
rows and indicator rows (tests/test_ref_agent.py, fixture
tests/golden/agent_model_ref.npz).  What pins the solutions:

1. every fixed-assignment QP is strictly convex, so a KKT certificate is
   self-validating; ``kkt_residual`` checks it in the *full* (x, u) variable
   space exactly as the reference poses the problem;
2. an independent third-party solver - HiGHS (the build vendored in
   scipy 1.18.1, ``scipy.optimize._highspy``) - solves the same full-space QP
   (``solve_qp_highs``) and must agree with the interior-point solver here;
3. the mixed-integer optimum is pinned by brute-force enumeration of all
   polytope assignments on small instances (``solve_miqp_enumerate``).

The formulation is deliberately the *un-condensed* one (all 9(N+1)+3N
variables, dynamics as equality rows) so that it shares no algebra with the
condensed null-space formulation used by the CUDA kernels.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

try:  # tiny dense systems: multi-threaded BLAS spins 8 threads on 240x240 LUs and is ~50x slower
    from threadpoolctl import threadpool_limits as _tpl
    _tpl(limits=1)
except Exception:  # pragma: no cover
    pass
import scipy.linalg as sla

OPTIMAL = 0
INFEASIBLE = 1
MAX_ITER = 2
NUMERICAL = 3
NODE_LIMIT = 4  # search stopped early; an incumbent (if any) is returned, like Gurobi's TimeLimit (:952)

FEAS_TOL = 1e-6  # Gurobi FeasibilityTol default; used for constant-row checks


# --------------------------------------------------------------------------------------
# parameters (mirror of hdsm_params in include/hdsm.h; values = agent_agile_config.yaml)
# --------------------------------------------------------------------------------------
@dataclass
class Params:
    n_hor: int = 10
    poly_hor: int = 4
    dt: float = 0.1
    rk4: bool = False
    drag: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    r_u: float = 0.01
    r_x: Tuple[float, ...] = (100.0, 100.0, 100.0, 1.0, 1.0, 1.0)
    r_n: Tuple[float, ...] = (100.0, 100.0, 100.0, 1.0, 1.0, 1.0)
    max_vel: float = 20.0
    min_acc_xy: float = -15.0
    max_acc_xy: float = 15.0
    min_acc_z: float = -15.0
    max_acc_z: float = 15.0
    max_jerk: float = 60.0
    drone_radius: float = 0.25
    drone_z_offset: float = 0.25
    tilt: float = 0.1  # var_tmp, agent_class.cpp:1180

    # agent_class.cpp:2179-2186 (n_x == 9 branch)
    def x_lb(self):
        return np.array([-np.inf] * 3 + [-self.max_vel] * 3 + [self.min_acc_xy, self.min_acc_xy, self.min_acc_z])

    def x_ub(self):
        return np.array([np.inf] * 3 + [self.max_vel] * 3 + [self.max_acc_xy, self.max_acc_xy, self.max_acc_z])

    def u_lb(self):
        return np.array([-self.max_jerk] * 3)

    def u_ub(self):
        return np.array([self.max_jerk] * 3)


# --------------------------------------------------------------------------------------
# dynamics  (agent_class.cpp:2115-2167)
# --------------------------------------------------------------------------------------
def model_ode(x, u, drag):
    """f(x,u) of ModelODE, agent_class.cpp:2155-2167 (hard-wired to 9 states)."""
    return np.array([
        x[3], x[4], x[5],
        x[6] - drag[0] * x[3], x[7] - drag[1] * x[4], x[8] - drag[2] * x[5],
        u[0], u[1], u[2],
    ])


def step_map(x, u, p: Params):
    """x_{k+1} = x_k + dt * mod_final, agent_class.cpp:2117-2151 (Euler or RK4)."""
    k1 = model_ode(x, u, p.drag)
    if p.rk4:
        k2 = model_ode(x + (p.dt / 2) * k1, u, p.drag)
        k3 = model_ode(x + (p.dt / 2) * k2, u, p.drag)
        k4 = model_ode(x + p.dt * k3, u, p.drag)
        mod = (k1 + 2 * k2 + 2 * k3 + k4) / 6
    else:
        mod = k1
    return x + p.dt * mod


def discrete_dynamics(p: Params):
    """(A, B) with x+ = A x + B u; obtained by pushing unit vectors through step_map
    (the map is linear and homogeneous, so this is exact)."""
    A = np.zeros((9, 9))
    B = np.zeros((9, 3))
    for j in range(9):
        e = np.zeros(9)
        e[j] = 1.0
        A[:, j] = step_map(e, np.zeros(3), p)
    for j in range(3):
        e = np.zeros(3)
        e[j] = 1.0
        B[:, j] = step_map(np.zeros(9), e, p)
    return A, B


def rollout(p: Params, x0, u):
    """States (N+1, 9) from x0 and inputs (N, 3)."""
    A, B = discrete_dynamics(p)
    xs = [np.asarray(x0, float)]
    for k in range(p.n_hor):
        xs.append(A @ xs[-1] + B @ u[k])
    return np.array(xs)


# --------------------------------------------------------------------------------------
# inter-agent separating planes  (agent_class.cpp:1086-1215)
# --------------------------------------------------------------------------------------
def interagent_plane(p: Params, pos_curr, pos_other):
    """One plane (normal n_f, offset n_f . pt) for own point pos_curr and neighbour point
    pos_other; line-by-line agent_class.cpp:1152-1205 with pert = 0 (:1197)."""
    pos_curr = np.asarray(pos_curr, float)
    pos_other = np.asarray(pos_other, float)
    plane_normal = pos_other - pos_curr
    nrm = math.sqrt(float(plane_normal @ plane_normal))
    # Eigen's normalized() returns the vector unchanged when its norm is 0 -> NaN planes
    # in the reference.  The oracle flags it with NaN as well.
    nn = plane_normal / nrm if nrm > 0 else plane_normal * np.nan
    pos_mid = (pos_curr + pos_other) / 2
    angle_x_axis = math.pi / 2 - abs(math.acos(max(-1.0, min(1.0, nn[2])))) if nrm > 0 else float("nan")
    t_val = math.atan(p.drone_radius / p.drone_z_offset * math.tan(angle_x_axis))
    x_val = p.drone_radius * math.cos(t_val)
    y_val = p.drone_z_offset * math.sin(t_val)
    safety_dist = math.hypot(x_val, y_val)
    plane_point = pos_mid - min(2 * safety_dist, nrm) / 2 * nn
    up = np.array([0.0, 0.0, 1.0])
    up_2 = np.array([0.0, 1.0, 0.0])
    right = np.cross(nn, up) + np.cross(nn, up_2)
    up_final = np.cross(nn, up_2)
    normal_final = (p.tilt + 0.0) * right + p.tilt * up_final + nn
    offset = float(normal_final @ plane_point)  # AddHyperplane :1227-1228
    return normal_final, offset


def time_aware_planes(p: Params, prev_self_pos, all_pos, all_valid, self_id):
    """GenerateTimeAwareSafeCorridor, agent_class.cpp:1096-1211.

    prev_self_pos : (N+1, 3) own previous plan positions (or state_ini repeated, :1103-1110)
    all_pos       : (n_rob, N+1, 3) last received plans of every agent
    all_valid     : (n_rob,) plan received?  (self is skipped like the empty own slot, :1132-1134)
    returns list over k of (normals (n_nb, 3), offsets (n_nb,)) in neighbour-id order.
    """
    N = p.n_hor
    out = []
    for k in range(N):
        ns, bs = [], []
        for j in range(all_pos.shape[0]):
            if j == self_id or not all_valid[j]:
                continue
            n_f, b = interagent_plane(p, prev_self_pos[k + 1], all_pos[j, k + 1])
            ns.append(n_f)
            bs.append(b)
        out.append((np.array(ns).reshape(-1, 3), np.array(bs).reshape(-1)))
    return out


# --------------------------------------------------------------------------------------
# full-space QP builder  (agent_class.cpp:858-941 + :2071-2153)
# --------------------------------------------------------------------------------------
@dataclass
class FullQP:
    """min z'diag(P)z + q'z + c0   s.t.  Aeq z = beq,  C z <= d   (no 1/2: agent_class.cpp:870-883)."""
    N: int
    Pdiag: np.ndarray
    q: np.ndarray
    c0: float
    Aeq: np.ndarray
    beq: np.ndarray
    C: np.ndarray
    d: np.ndarray
    infeasible_constant_row: bool = False
    n_dropped_constant: int = 0

    @property
    def nz(self):
        return self.Pdiag.size

    def objective(self, z):
        return float(z @ (self.Pdiag * z) + self.q @ z + self.c0)

    def split(self, z):
        N = self.N
        return z[: 9 * (N + 1)].reshape(N + 1, 9), z[9 * (N + 1):].reshape(N, 3)


def xi(k, j):
    return 9 * k + j


def ui(N, k, j):
    return 9 * (N + 1) + 3 * k + j


def _sensitivity(p: Params):
    """G[k] (9, 3N): d x_k / d u, used only to classify rows as constant given x0."""
    A, B = discrete_dynamics(p)
    N = p.n_hor
    G = np.zeros((N + 1, 9, 3 * N))
    for k in range(N):
        G[k + 1] = A @ G[k]
        G[k + 1][:, 3 * k:3 * k + 3] += B
    return G


def build_qp_full(p: Params, x0, ref, polys: Sequence[Tuple[np.ndarray, np.ndarray]],
                  planes: Sequence[Tuple[np.ndarray, np.ndarray]], sigma: Sequence[int],
                  drop_constant_rows: bool = True, hull=None) -> FullQP:
    """The continuous problem for a fixed per-step polytope assignment sigma[k].

    polys  : list of (A (R,3), b (R,)) static corridor polytopes (poly_const_vec_)
    planes : output of time_aware_planes (identical for every polytope at step k, :1205)
    sigma  : N polytope indices, or -1 to leave a step's static rows out (B&B relaxation;
             neighbour planes stay - they belong to every polytope).
    hull   : optional (A, b) rows valid for the union of all polytopes (union_hull_rows);
             imposed on the undecided steps to tighten the relaxation.
    """
    N = p.n_hor
    nx = 9 * (N + 1)
    nz = nx + 3 * N
    A, B = discrete_dynamics(p)
    x0 = np.asarray(x0, float)
    ref = np.asarray(ref, float)

    Pd = np.zeros(nz)
    q = np.zeros(nz)
    c0 = 0.0
    for k in range(N):  # :2098
        for j in range(3):
            Pd[ui(N, k, j)] = p.r_u
    for i in range(1, N + 1):  # :871-883  x_i tracks ref[i-1], pos+vel only
        w = p.r_n if i == N else p.r_x
        for j in range(6):
            Pd[xi(i, j)] += w[j]
            q[xi(i, j)] += -2.0 * w[j] * ref[i - 1][j]
            c0 += w[j] * ref[i - 1][j] ** 2

    rows, rhs = [], []
    for j in range(9):  # x0 fixed, :886-889
        r = np.zeros(nz)
        r[xi(0, j)] = 1
        rows.append(r)
        rhs.append(x0[j])
    for k in range(N):  # dynamics, :2146-2151
        for j in range(9):
            r = np.zeros(nz)
            r[xi(k + 1, j)] = 1
            r[9 * k:9 * k + 9] -= A[j]
            r[nx + 3 * k:nx + 3 * k + 3] -= B[j]
            rows.append(r)
            rhs.append(0.0)
    for j in range(3, 9):  # terminal vel = acc = 0, :2078-2081
        r = np.zeros(nz)
        r[xi(N, j)] = 1
        rows.append(r)
        rhs.append(0.0)
    Aeq = np.array(rows)
    beq = np.array(rhs)

    G = _sensitivity(p)
    Phi = [np.eye(9)]
    for k in range(N):
        Phi.append(A @ Phi[-1])
    crow, cd = [], []
    infeas = False
    ndrop = 0

    def add_state_row(k, coef9, bound):
        nonlocal infeas, ndrop
        if drop_constant_rows and not np.any(coef9 @ G[k]):
            val = coef9 @ (Phi[k] @ x0)
            if val - bound > FEAS_TOL:
                infeas = True
            ndrop += 1
            return
        r = np.zeros(nz)
        r[9 * k:9 * k + 9] = coef9
        crow.append(r)
        cd.append(bound)

    xl, xu = p.x_lb(), p.x_ub()
    for k in range(1, N):  # state boxes; k = 0 is overwritten by x0, k = N is fixed/free
        for j in range(3, 9):
            e = np.zeros(9)
            e[j] = 1
            add_state_row(k, e, xu[j])
            add_state_row(k, -e, -xl[j])
    ul, uu = p.u_lb(), p.u_ub()
    for k in range(N):
        for j in range(3):
            r = np.zeros(nz)
            r[ui(N, k, j)] = 1
            crow.append(r)
            cd.append(uu[j])
            crow.append(-r)
            cd.append(-ul[j])
    for k in range(N):  # corridor rows on x_k and x_{k+1}, :909-937 / :1071-1084
        blocks = []
        if sigma[k] >= 0:
            blocks.append(polys[sigma[k]])
        elif hull is not None:
            blocks.append(hull)
        blocks.append(planes[k])
        for (Ak, bk) in blocks:
            for i in range(len(bk)):
                e = np.zeros(9)
                e[:3] = Ak[i]
                add_state_row(k, e, bk[i])
                add_state_row(k + 1, e, bk[i])
    C = np.array(crow).reshape(-1, nz)
    d = np.array(cd)
    return FullQP(N, Pd, q, c0, Aeq, beq, C, d, infeas, ndrop)


# --------------------------------------------------------------------------------------
# KKT certificate
# --------------------------------------------------------------------------------------
def kkt_residual(qp: FullQP, z, y, lam):
    """max of scaled stationarity, equality, inequality and complementarity residuals
    (SURVEY 8(d)); rows are scaled to unit 2-norm for the last two."""
    g = 2.0 * qp.Pdiag * z + qp.q
    stat = g + qp.Aeq.T @ y + qp.C.T @ lam
    rs = np.linalg.norm(qp.C, axis=1)
    rs[rs == 0] = 1.0
    slack = (qp.d - qp.C @ z) / rs
    lam_s = lam * rs
    r_stat = np.max(np.abs(stat)) / (1.0 + np.max(np.abs(qp.q)))
    r_eq = np.max(np.abs(qp.Aeq @ z - qp.beq)) if qp.beq.size else 0.0
    r_in = max(0.0, float(np.max(-slack))) if slack.size else 0.0
    r_dual = max(0.0, float(np.max(-lam))) if lam.size else 0.0
    r_comp = float(np.max(np.abs(lam_s * slack))) / (1.0 + abs(qp.objective(z))) if slack.size else 0.0
    return max(r_stat, r_eq, r_in, r_dual, r_comp)


# --------------------------------------------------------------------------------------
# solver 1: dense Mehrotra predictor-corrector on the full-space problem
# --------------------------------------------------------------------------------------
@dataclass
class QPResult:
    status: int
    z: Optional[np.ndarray] = None
    y: Optional[np.ndarray] = None
    lam: Optional[np.ndarray] = None
    obj: float = float("inf")
    iters: int = 0
    kkt: float = float("inf")


def solve_qp_pdip(qp: FullQP, max_iter: int = 80, tol: float = 1e-9) -> QPResult:
    if qp.infeasible_constant_row:
        return QPResult(INFEASIBLE)
    H = np.diag(2.0 * qp.Pdiag)
    g = qp.q
    A, b, C, d = qp.Aeq, qp.beq, qp.C, qp.d
    n, me, mi = qp.nz, b.size, d.size
    z = np.linalg.lstsq(A, b, rcond=None)[0]
    s = np.maximum(d - C @ z, 1.0)
    lam = np.ones(mi)
    y = np.zeros(me)
    gscale = 1.0 + np.max(np.abs(g))
    for it in range(max_iter):
        rd = H @ z + g + A.T @ y + C.T @ lam
        rp = A @ z - b
        rc = C @ z + s - d
        mu = float(s @ lam) / max(mi, 1)
        if (np.max(np.abs(rd)) <= tol * gscale and np.max(np.abs(rp)) <= tol
                and (mi == 0 or np.max(np.abs(rc)) <= tol) and mu <= tol):
            res = QPResult(OPTIMAL, z, y, lam, qp.objective(z), it)
            res.kkt = kkt_residual(qp, z, y, lam)
            return res
        # Farkas-type infeasibility certificate: lam >= 0, A'y + C'lam ~ 0, b'y + d'lam < 0
        nl = np.max(np.abs(lam)) + np.max(np.abs(y)) if mi else 0.0
        if nl > 1e6:
            cert = np.max(np.abs(A.T @ y + C.T @ lam)) / nl
            if cert < 1e-6 and (b @ y + d @ lam) / nl < -1e-7:
                return QPResult(INFEASIBLE, iters=it)
        D = lam / s
        K = np.zeros((n + me, n + me))
        K[:n, :n] = H + C.T @ (D[:, None] * C)
        K[:n, n:] = A.T
        K[n:, :n] = A
        K[n:, n:] = -1e-13 * np.eye(me)
        try:
            lu = sla.lu_factor(K)
        except Exception:
            return QPResult(NUMERICAL, iters=it)

        def solve(rs):
            rhs = np.concatenate([-rd + C.T @ (rs / s - D * rc), -rp])
            sol = sla.lu_solve(lu, rhs)
            dz, dy = sol[:n], sol[n:]
            ds = -rc - C @ dz
            dl = -(rs + lam * ds) / s
            return dz, dy, ds, dl

        def steplen(ds, dl):
            a = 1.0
            neg = ds < 0
            if np.any(neg):
                a = min(a, float(np.min(-s[neg] / ds[neg])))
            neg = dl < 0
            if np.any(neg):
                a = min(a, float(np.min(-lam[neg] / dl[neg])))
            return a

        dz, dy, ds, dl = solve(s * lam)
        a_aff = steplen(ds, dl)
        mu_aff = float((s + a_aff * ds) @ (lam + a_aff * dl)) / max(mi, 1)
        sig = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dz, dy, ds, dl = solve(s * lam + ds * dl - sig * mu)
        a = min(1.0, 0.995 * steplen(ds, dl))
        z = z + a * dz
        y = y + a * dy
        s = s + a * ds
        lam = lam + a * dl
        if not np.all(np.isfinite(z)):
            return QPResult(NUMERICAL, iters=it)
    return QPResult(MAX_ITER, z, y, lam, qp.objective(z), max_iter)


# --------------------------------------------------------------------------------------
# solver 2: HiGHS (independent third-party QP solver vendored in scipy)
# --------------------------------------------------------------------------------------
def solve_qp_highs(qp: FullQP) -> QPResult:
    from scipy.optimize._highspy import _core as hc
    from scipy.sparse import csr_matrix

    if qp.infeasible_constant_row:
        return QPResult(INFEASIBLE)
    n = qp.nz
    Aall = csr_matrix(np.vstack([qp.Aeq, qp.C]))
    lo = np.concatenate([qp.beq, np.full(qp.d.size, -hc.kHighsInf)])
    hi = np.concatenate([qp.beq, qp.d])
    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    lp = hc.HighsLp()
    lp.num_col_ = n
    lp.num_row_ = Aall.shape[0]
    lp.col_cost_ = qp.q.astype(float)
    lp.offset_ = float(qp.c0)
    lp.col_lower_ = np.full(n, -hc.kHighsInf)
    lp.col_upper_ = np.full(n, hc.kHighsInf)
    lp.row_lower_ = lo
    lp.row_upper_ = hi
    lp.a_matrix_.format_ = hc.MatrixFormat.kRowwise
    lp.a_matrix_.start_ = Aall.indptr.astype(np.int32)
    lp.a_matrix_.index_ = Aall.indices.astype(np.int32)
    lp.a_matrix_.value_ = Aall.data.astype(float)
    h.passModel(lp)
    hs = hc.HighsHessian()
    hs.dim_ = n
    hs.format_ = hc.HessianFormat.kTriangular
    # HiGHS minimises 1/2 z'Qz: pass 2*diag(P), structural zeros left out (default tolerances;
    # its active-set QP solver reports "solve error" when they are tightened to 1e-9).
    nzd = qp.Pdiag != 0
    hs.start_ = np.concatenate([[0], np.cumsum(nzd)]).astype(np.int32)
    hs.index_ = np.nonzero(nzd)[0].astype(np.int32)
    hs.value_ = 2.0 * qp.Pdiag[nzd]
    h.passHessian(hs)
    h.run()
    st = h.getModelStatus()
    if st == hc.HighsModelStatus.kInfeasible:
        return QPResult(INFEASIBLE)
    if st != hc.HighsModelStatus.kOptimal:
        return QPResult(NUMERICAL)
    sol = h.getSolution()
    z = np.array(sol.col_value)
    rd = np.array(sol.row_dual)
    me = qp.beq.size
    res = QPResult(OPTIMAL, z, -rd[:me], -rd[me:], qp.objective(z))
    res.kkt = kkt_residual(qp, z, res.y, np.maximum(res.lam, 0.0))
    return res


# --------------------------------------------------------------------------------------
# mixed-integer layer  (binaries b[k][p], indicator rows, sum_p b[k][p] == 1: :909-941)
# --------------------------------------------------------------------------------------
def p_eff(p: Params, polys):
    return min(p.poly_hor, len(polys))  # :913


def union_hull_rows(polys):
    """Rows valid for the union of the polytopes: a normal that appears (bit-identical) in
    every polytope gives the row (a, max_p min_{rows of p with normal a} b).  The reference's
    cells always share the six axis faces (convex_decomp.cpp:359-373), so this is at least the
    bounding box of the union.  Row order follows the first polytope."""
    A0, b0 = polys[0]
    rows, rhs = [], []
    for i in range(len(b0)):
        if any(np.array_equal(A0[i], r) for r in rows):
            continue
        bmax = -np.inf
        ok = True
        for (A, b) in polys:
            m = np.nonzero(np.all(A == A0[i][None, :], axis=1))[0]
            if len(m) == 0:
                ok = False
                break
            bmax = max(bmax, float(np.min(b[m])))
        if ok:
            rows.append(A0[i])
            rhs.append(bmax)
    return np.array(rows).reshape(-1, 3), np.array(rhs)


def segment_violation(poly, pa, pb):
    A, b = poly
    return max(float(np.max(A @ pa - b)), float(np.max(A @ pb - b)))


@dataclass
class MIQPResult:
    status: int
    sigma: Optional[List[int]] = None
    obj: float = float("inf")
    traj: Optional[np.ndarray] = None
    ctrl: Optional[np.ndarray] = None
    nodes: int = 0
    poly_used: Optional[np.ndarray] = None
    qp: Optional[QPResult] = None


def _finish(p, polys, res: MIQPResult, qp: FullQP):
    xs, us = qp.split(res.qp.z)
    res.traj, res.ctrl = xs, us
    used = np.zeros(p.poly_hor, bool)
    for s in res.sigma:
        used[s] = True  # :981-985 restricted to sigma
    res.poly_used = used
    return res


def solve_miqp_enumerate(p: Params, x0, ref, polys, planes, solver=solve_qp_pdip) -> MIQPResult:
    """Brute force over every assignment (small instances only): pins the MIQP optimum."""
    P = p_eff(p, polys)
    if P == 0:  # sum over an empty set == 1 cannot hold (:939-940)
        return MIQPResult(INFEASIBLE)
    best = MIQPResult(INFEASIBLE)
    n = 0
    for sigma in itertools.product(range(P), repeat=p.n_hor):
        qp = build_qp_full(p, x0, ref, polys, planes, sigma)
        r = solver(qp)
        n += 1
        if r.status == OPTIMAL and r.obj < best.obj - 1e-9 * max(1.0, abs(r.obj)):
            best = MIQPResult(OPTIMAL, list(sigma), r.obj, qp=r)
            bestqp = qp
    best.nodes = n
    if best.status == OPTIMAL:
        _finish(p, polys, best, bestqp)
    return best


def solve_miqp_bnb(p: Params, x0, ref, polys, planes, solver=solve_qp_pdip,
                   contain_tol: float = 1e-7, prune_rel: float = 1e-7,
                   sigma_fixed: Optional[Sequence[int]] = None, max_nodes: int = 100000) -> MIQPResult:
    """Exact depth-first branch and bound over candidate *sets* (refined).

    A node keeps, for every step k, a set S_k of still-allowed polytopes.  Its bound is the QP
    whose rows at step k are union_hull_rows(S_k) (all rows of the polytope when |S_k| = 1) -
    a valid relaxation of "the segment lies in one member of S_k".  If every segment
    (p_k, p_{k+1}) of the node optimum already lies inside one member of S_k the node optimum is
    MIQP-feasible and the node is fathomed; otherwise the set of the most violated uncovered step is
    split in two halves ordered by violation, least-violated half explored first.  Any split rule is exact;
    this one collapses near-identical overlapping cells into a single subtree.
    """
    N = p.n_hor
    P = p_eff(p, polys)
    if P == 0:
        return MIQPResult(INFEASIBLE)
    if sigma_fixed is None:
        root = [tuple(range(P))] * N
    else:
        root = [tuple(range(P)) if s < 0 else (int(s),) for s in sigma_fixed]
    stack = [root]
    best = MIQPResult(INFEASIBLE)
    bestqp = None
    nodes = 0
    hull_cache = {}

    def rows_of(S):
        if S not in hull_cache:
            hull_cache[S] = polys[S[0]] if len(S) == 1 else union_hull_rows([polys[j] for j in S])
        return hull_cache[S]

    while stack and nodes < max_nodes:
        sets = stack.pop()
        qp = build_qp_full(p, x0, ref, [rows_of(S) for S in sets], planes, list(range(N)))
        r = solver(qp)
        nodes += 1
        if r.status != OPTIMAL:
            continue
        if r.obj >= best.obj - prune_rel * max(1.0, abs(best.obj)):
            continue
        xs, _ = qp.split(r.z)
        full = [-1] * N
        branch_k, branch_v = -1, 0.0
        for k in range(N):
            viol = {j: segment_violation(polys[j], xs[k, :3], xs[k + 1, :3]) for j in sets[k]}
            ok = [j for j in sets[k] if viol[j] <= contain_tol]
            if ok:
                full[k] = ok[0]
            elif branch_k < 0 or min(viol.values()) > branch_v:  # farthest from all its candidates
                branch_k, branch_v = k, min(viol.values())
                order = sorted(sets[k], key=lambda j: (viol[j], j))
        if branch_k < 0:
            best = MIQPResult(OPTIMAL, full, r.obj, qp=r)
            bestqp = qp
            continue
        if len(order) == 1:
            continue  # singleton set that does not contain its own optimum: cannot happen, guard
        h = (len(order) + 1) // 2
        for part in (order[h:], order[:h]):  # pushed worse half first -> better half popped first
            child = list(sets)
            child[branch_k] = tuple(sorted(part))
            stack.append(child)
    best.nodes = nodes
    if best.status == OPTIMAL:
        if stack:
            best.status = NODE_LIMIT
        _finish(p, polys, best, bestqp)
    elif stack:
        best.status = NODE_LIMIT
    return best


def fallback_shift(traj_prev, ctrl_prev):
    """agent_class.cpp:1004-1018: drop the first element, duplicate the last."""
    traj = np.concatenate([traj_prev[1:], traj_prev[-1:]], axis=0)
    ctrl = np.concatenate([ctrl_prev[1:], ctrl_prev[-1:]], axis=0)
    return traj, ctrl
