"""ctypes front-end of oracle/_ref/libref_agent.so (TEST INFRASTRUCTURE ONLY): the reference's OWN planner node
(multi_agent_planner/src/agent_class.cpp compiled unmodified on stand-in ROS / Eigen headers and a RECORDING stand-in for the
Gurobi C++ API).  `RefAgent.step` runs its GenerateTimeAwareSafeCorridor + SolveOptimizationProblem and returns the inter-agent
planes and the optimisation model the reference hands to Gurobi.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_agent.so")
_lib = None


def have_ref():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.ref_agent_create.restype = C.c_void_p
        _lib.ref_agent_num_vars.restype = C.c_int
        _lib.ref_agent_num_vars.argtypes = [C.c_void_p]
        _lib.ref_agent_step.restype = C.c_int
        _lib.ref_agent_reftraj.restype = C.c_int
        _lib.ref_agent_corridor.restype = C.c_int
        _lib.ref_agent_loop_step.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(v):
    return (C.c_double * len(v))(*[float(x) for x in v])


class RefAgent:
    """One node of the reference (never destroyed: the reference does not join its worker threads)."""

    def __init__(self, p, n_rob, agent_id, state_ini):
        """p: oracle.hdsm_oracle.Params."""
        L = lib()
        r_x, r_n = list(p.r_x) + [0.0] * (9 - len(p.r_x)), list(p.r_n) + [0.0] * (9 - len(p.r_n))
        si = list(state_ini) + [0.0] * (9 - len(state_ini))
        self.h = C.c_void_p(L.ref_agent_create(
            C.c_int(n_rob), C.c_int(agent_id), C.c_int(p.n_hor), C.c_int(p.poly_hor), C.c_double(p.dt), C.c_int(int(p.rk4)), C.c_double(p.r_u),
            _d(r_x), _d(r_n), C.c_double(p.max_vel), C.c_double(p.min_acc_xy), C.c_double(p.max_acc_xy), C.c_double(p.min_acc_z),
            C.c_double(p.max_acc_z), C.c_double(p.max_jerk), C.c_double(p.drone_radius), C.c_double(p.drone_z_offset), _d(p.drag), _d(si)))
        self.p, self.n_rob, self.id = p, n_rob, agent_id
        self.nz = L.ref_agent_num_vars(self.h)

    def step(self, x0, ref, polys, prev_traj, all_pos, all_valid):
        """x0 (9,), ref (N+1, 6), polys [(A (R,3), b (R,))], prev_traj (N+1, 9) or None, all_pos (n_rob, N+1, 3), all_valid (n_rob,).
        Returns dict(final=[[ (A, b) per polytope ] per step], obj_diag, obj_lin, obj_const, obj_offdiag, lb, ub, vtype, lin (rows, const,
        sense), ind (rows, const, bin), failed)."""
        N, P, nz = self.p.n_hor, len(polys), self.nz
        rmax = max(len(b) for _, b in polys)
        fmax = rmax + self.n_rob
        rows = np.array([len(b) for _, b in polys], np.int32)
        A, b = np.zeros((P, rmax, 3)), np.zeros((P, rmax))
        for i, (Ai, bi) in enumerate(polys):
            A[i, :len(bi)], b[i, :len(bi)] = Ai, bi
        x0, ref = np.ascontiguousarray(x0, np.float64), np.ascontiguousarray(ref, np.float64).reshape(N + 1, 6)
        prev = None if prev_traj is None else np.ascontiguousarray(prev_traj, np.float64).reshape(N + 1, 9)
        all_pos, all_valid = np.ascontiguousarray(all_pos, np.float64), np.ascontiguousarray(all_valid, np.uint8)
        final_rows, final_A, final_b = np.zeros((N, P), np.int32), np.zeros((N, P, fmax, 3)), np.zeros((N, P, fmax))
        obj_diag, obj_lin, lb, ub, vtype = np.zeros(nz), np.zeros(nz), np.zeros(nz), np.zeros(nz), np.zeros(nz, np.int8)
        obj_const, obj_off = C.c_double(), C.c_double()
        cap_lin, cap_ind = 9 * N + N + 8, 2 * N * P * fmax
        n_lin, n_ind, failed = C.c_int32(), C.c_int32(), C.c_int32()
        lin, lin_c, lin_s = np.zeros((cap_lin, nz)), np.zeros(cap_lin), np.zeros(cap_lin, np.int8)
        ind, ind_c, ind_b = np.zeros((cap_ind, nz)), np.zeros(cap_ind), np.zeros(cap_ind, np.int32)
        rc = lib().ref_agent_step(self.h, _p(x0), _p(ref), C.c_int(P), _p(rows), C.c_int(rmax), _p(A), _p(b), C.c_int(int(prev is not None)), _p(prev),
                                  _p(all_pos), _p(all_valid), C.c_int(fmax), _p(final_rows), _p(final_A), _p(final_b), _p(obj_diag), _p(obj_lin),
                                  C.byref(obj_const), C.byref(obj_off), _p(lb), _p(ub), _p(vtype), C.c_int(cap_lin), C.byref(n_lin), _p(lin), _p(lin_c),
                                  _p(lin_s), C.c_int(cap_ind), C.byref(n_ind), _p(ind), _p(ind_c), _p(ind_b), C.byref(failed))
        if rc != 0:
            raise RuntimeError("ref_agent_step: capacity too small")
        final = [[(final_A[k, q, :final_rows[k, q]].copy(), final_b[k, q, :final_rows[k, q]].copy()) for q in range(P)] for k in range(N)]
        nl, ni = n_lin.value, n_ind.value
        return dict(final=final, obj_diag=obj_diag, obj_lin=obj_lin, obj_const=obj_const.value, obj_offdiag=obj_off.value, lb=lb, ub=ub,
                    vtype=vtype, lin=(lin[:nl], lin_c[:nl], lin_s[:nl]), ind=(ind[:ni], ind_c[:ni], ind_b[:ni]), failed=bool(failed.value))

    def loop_step(self, ref, polys, all_pos, all_valid, solver, rmax=18):
        """One closed-loop replanning iteration of the node on an external solver (see ref_agent_loop_step in ref_wrap_agent.cpp).
        ref (N+1, 6); polys [(A, b)]; all_pos (n_rob, N+1, 3) / all_valid (n_rob,): the other agents' last published plans.
        solver: ("hdsm", TrajectoryPlanner)  - hdsm_solve_batch of the CUDA library through its C ABI, or
                ("port", OrcParams)          - orc_solve_batch of the C port of the oracle (CPU-only runs).
        Returns dict(x0, traj, ctrl, poly_used, have_traj, failed, status) read from the node's members afterwards."""
        N, P = self.p.n_hor, self.p.poly_hor
        rows = np.array([len(b) for _, b in polys], np.int32)
        A, b = np.zeros((len(polys), rmax, 3)), np.zeros((len(polys), rmax))
        for i, (Ai, bi) in enumerate(polys):
            A[i, :len(bi)], b[i, :len(bi)] = Ai, bi
        ref = np.ascontiguousarray(ref, np.float64).reshape(N + 1, 6)
        all_pos, all_valid = np.ascontiguousarray(all_pos, np.float64), np.ascontiguousarray(all_valid, np.uint8)
        kind, obj = solver
        if kind == "hdsm":
            fn, ctx, k = C.cast(obj.lib.hdsm_solve_batch, C.c_void_p), obj._h, 0
        else:
            from . import c_oracle
            fn, ctx, k = C.cast(c_oracle.lib().orc_solve_batch, C.c_void_p), C.cast(C.pointer(obj), C.c_void_p), 1
        x0, traj, ctrl, used = np.zeros(9), np.zeros((N + 1, 9)), np.zeros((N, 3)), np.zeros(P, np.uint8)
        have, failed, status = C.c_int32(), C.c_int32(), C.c_int32()
        rc = lib().ref_agent_loop_step(self.h, _p(ref), C.c_int(len(polys)), _p(rows), C.c_int(rmax), _p(A), _p(b), _p(all_pos), _p(all_valid),
                                       C.c_int(k), fn, ctx, _p(x0), _p(traj), _p(ctrl), _p(used), C.byref(have), C.byref(failed), C.byref(status))
        if rc != 0:
            raise RuntimeError(f"ref_agent_loop_step failed ({rc})")
        return dict(x0=x0, traj=traj, ctrl=ctrl, poly_used=used, have_traj=bool(have.value), failed=bool(failed.value), status=status.value)

    def reference_trajectory(self, grid, origin, voxel, path, prev_ref, increment, traj, all_pos, all_valid, path_vel_min, path_vel_max,
                             path_vel_dec, sens_dist, sens_pot, sens_other_agents):
        """The reference's own GenerateReferenceTrajectory.  grid [dz][dy][dx] int8, path (n_path, 3), prev_ref (N+1, 3) or None,
        traj (N+1, 3) or None.  Returns (traj_ref_curr_ (N+1, 6), path_vel_)."""
        N = self.p.n_hor
        grid = np.ascontiguousarray(grid, np.int8)
        dim = np.array([grid.shape[2], grid.shape[1], grid.shape[0]], np.int32)
        path = np.ascontiguousarray(path, np.float64).reshape(-1, 3)
        prev = None if prev_ref is None else np.ascontiguousarray(prev_ref, np.float64).reshape(N + 1, 3)
        tr = None if traj is None else np.ascontiguousarray(traj, np.float64).reshape(N + 1, 3)
        all_pos, all_valid = np.ascontiguousarray(all_pos, np.float64), np.ascontiguousarray(all_valid, np.uint8)
        out, vel = np.zeros((N + 1, 6)), C.c_double()
        rows = lib().ref_agent_reftraj(self.h, _p(grid), _p(dim), _d(origin), C.c_double(voxel), _p(path), C.c_int(path.shape[0]),
                                       C.c_int(int(prev is not None)), _p(prev), C.c_int(int(increment)), C.c_int(int(tr is not None)), _p(tr),
                                       _p(all_pos), _p(all_valid), _d([path_vel_min, path_vel_max, path_vel_dec, sens_dist, sens_pot, sens_other_agents]),
                                       _p(out), C.byref(vel))
        if rows != N + 1:
            raise RuntimeError(f"GenerateReferenceTrajectory gave {rows} rows")
        return out, vel.value

    def safe_corridor(self, grid, origin, voxel, pos, path, n_it, use_cvx_new, prev, prev_traj, rmax=18):
        """The reference's own GenerateSafeCorridor.  prev: None or dict(rows (n,), A (n, rmax, 3), b (n, rmax), seeds (n, 3), used (n,));
        prev_traj (n_traj, 3).  Returns (rows (P,), A (P, rmax, 3), b (P, rmax), seeds (P, 3)) padded with zeros to poly_hor entries."""
        P = self.p.poly_hor
        grid = np.ascontiguousarray(grid, np.int8)
        dim = np.array([grid.shape[2], grid.shape[1], grid.shape[0]], np.int32)
        path = np.ascontiguousarray(path, np.float64).reshape(-1, 3)
        ptraj = np.ascontiguousarray(prev_traj, np.float64).reshape(-1, 3)
        n_prev = 0 if prev is None else len(prev["rows"])
        pr = pa = pb = ps = pu = None
        if n_prev:
            pr, pa, pb = np.ascontiguousarray(prev["rows"], np.int32), np.ascontiguousarray(prev["A"], np.float64), np.ascontiguousarray(prev["b"], np.float64)
            ps, pu = np.ascontiguousarray(prev["seeds"], np.float64), np.ascontiguousarray(prev["used"], np.uint8)
        cap = P + 4
        rows, A, b, seeds = np.zeros(cap, np.int32), np.zeros((cap, rmax, 3)), np.zeros((cap, rmax)), np.zeros((cap, 3))
        n = lib().ref_agent_corridor(self.h, _p(grid), _p(dim), _d(origin), C.c_double(voxel), _d(pos), _p(path), C.c_int(path.shape[0]), C.c_int(n_it),
                                     C.c_int(int(use_cvx_new)), C.c_int(n_prev), _p(pr), C.c_int(rmax), _p(pa), _p(pb), _p(ps), _p(pu),
                                     C.c_int(ptraj.shape[0]), _p(ptraj), C.c_int(cap), _p(rows), _p(A), _p(b), _p(seeds))
        if n < 0:
            raise RuntimeError(f"GenerateSafeCorridor failed ({n})")
        return n, rows[:P], A[:P], b[:P], seeds[:P]
