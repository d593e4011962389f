"""Host-side mirror of the reference interface for the replaced path.

The reference exposes the optimisation as three member functions of ``class Agent``
(multi_agent_planner/include/agent_class.hpp:63-70, 100-104):

* ``CreateGurobiModel()``               agent_class.cpp:2071-2153  -> :class:`TrajectoryPlanner` constructor
* ``GenerateTimeAwareSafeCorridor()``   agent_class.cpp:1086-1215  \\_ :meth:`TrajectoryPlanner.solve_batch`
* ``SolveOptimizationProblem()``        agent_class.cpp:858-1023   /   (planes are built on the device)

:class:`TrajectoryPlanner` is the batched form (any number of agents per call, what the swarm
harness uses); :class:`AgentSolver` keeps the reference's per-agent member names and failure
semantics on top of it.  Everything goes through the C ABI of ``include/hdsm.h``; there is no CPU
path here - without the CUDA library the constructors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import HdsmResult, RESULT_DTYPE, STATUS_NAMES  # noqa: F401


class HdsmError(RuntimeError):
    pass


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class TrajectoryPlanner:
    """One library handle = one ``GRBModel`` worth of set-up, valid for a fixed parameter set."""

    def __init__(self, params: Dict, max_agents: int, max_neighbours: int, device: int = 0, rmax: int = 18,
                 max_iter: int = 60, max_nodes: int = 64, prune: bool = True, tol: float = 1e-8, width: int = 1,
                 warm_start: bool = False):
        self.lib = _lib.load()
        self.params = dict(params)
        self.N, self.P, self.rmax = int(params["n_hor"]), int(params["poly_hor"]), int(rmax)
        self.max_agents = int(max_agents)
        self._cparams = _lib.make_params(params, rmax, max_iter, max_nodes, prune, tol, width, warm_start)
        h = C.c_void_p()
        rc = self.lib.hdsm_create(C.byref(self._cparams), int(max_agents), int(max_neighbours), int(device), C.byref(h))
        if rc != 0:
            raise HdsmError(f"hdsm_create failed with code {rc} (no CUDA device, or unsupported parameters)")
        self._h = h
        self.device = device

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.hdsm_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HdsmError(f"libhdsm error {rc}: {self.lib.hdsm_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self.lib.hdsm_launch_count(self._h))

    @property
    def smem_bytes(self) -> int:
        return int(self.lib.hdsm_smem_bytes(self._h))

    # -- host-pointer path (what the ROS node calls with n_local = 1) -----------------------------
    def solve_batch(self, batch, assign_in: Optional[np.ndarray] = None, out: Optional[Dict] = None) -> Dict[str, np.ndarray]:
        """``batch`` has the attributes of :class:`multi_agent_pkgs_b200.scenarios.Batch`.  ``out``: optional dict of
        preallocated result arrays (traj, ctrl, poly_used, assign, res) - page-locked ones are written by the copy
        engine directly, like page-locked inputs are read directly (no staging copy inside the library)."""
        n, N, P, R = int(batch.x0.shape[0]), self.N, self.P, self.rmax
        gid = _c(batch.global_id, np.int32)
        nb0 = _c(batch.nbr_begin, np.int32) if batch.nbr_begin is not None else None
        nb1 = _c(batch.nbr_end, np.int32) if batch.nbr_end is not None else None
        x0, ref = _c(batch.x0, np.float64), _c(batch.ref, np.float64)
        pA, pb, pr = _c(batch.poly_A, np.float64), _c(batch.poly_b, np.float64), _c(batch.poly_rows, np.int32)
        prev, allp = _c(batch.prev_self_pos, np.float64), _c(batch.all_pos, np.float64)
        allv = _c(batch.all_valid, np.uint8)
        if x0.shape != (n, 9) or ref.shape != (n, N, 6) or pA.shape != (n, P, R, 3) or pb.shape != (n, P, R) \
                or pr.shape != (n, P) or prev.shape != (n, N + 1, 3) or allp.shape[1:] != (N + 1, 3) \
                or allv.shape != (allp.shape[0],):
            raise ValueError("array shapes do not match the handle's n_hor / poly_hor / max_rows_per_poly")
        ain = _c(assign_in, np.int32) if assign_in is not None else None
        if out is not None:
            traj, ctrl, used, aout, res = out["traj"], out["ctrl"], out["poly_used"], out["assign"], out["res"]
            if traj.shape != (n, N + 1, 9) or ctrl.shape != (n, N, 3) or used.shape != (n, P) or aout.shape != (n, N) \
                    or res.shape != (n,) or res.dtype != RESULT_DTYPE:
                raise ValueError("preallocated outputs do not match the batch")
        else:
            traj = np.zeros((n, N + 1, 9))
            ctrl = np.zeros((n, N, 3))
            used = np.zeros((n, P), np.uint8)
            aout = np.zeros((n, N), np.int32)
            res = np.zeros(n, RESULT_DTYPE)
        i32, f64, u8 = C.c_int32, C.c_double, C.c_uint8
        self._check(self.lib.hdsm_solve_batch(
            self._h, n, _p(gid, i32), _p(nb0, i32), _p(nb1, i32), _p(x0, f64), _p(ref, f64), _p(pA, f64), _p(pb, f64),
            _p(pr, i32), _p(prev, f64), _p(allp, f64), _p(allv, u8), int(allp.shape[0]), _p(ain, i32), _p(traj, f64),
            _p(ctrl, f64), _p(used, u8), _p(aout, i32), res.ctypes.data_as(C.POINTER(HdsmResult))))
        return dict(traj=traj, ctrl=ctrl, poly_used=used, assign=aout, res=res)

    # -- device-pointer path (swarm harness; tensors are torch CUDA tensors) ----------------------
    def solve_batch_device(self, t: Dict, n_rob: int, stream_ptr: int = 0):
        """``t``: dict of contiguous CUDA tensors with the C-ABI layouts (see DeviceBatch in swarm.py).
        Enqueues one kernel on ``stream_ptr`` (0 = the handle's own stream); no synchronisation."""
        def dp(name):
            x = t.get(name)
            return C.c_void_p(x.data_ptr()) if x is not None else None
        n = int(t["x0"].shape[0])
        self._check(self.lib.hdsm_solve_batch_device(
            self._h, n, dp("global_id"), dp("nbr_begin"), dp("nbr_end"), dp("x0"), dp("ref"), dp("poly_A"),
            dp("poly_b"), dp("poly_rows"), dp("prev_self_pos"), dp("all_pos"), dp("all_valid"), int(n_rob),
            dp("assign_in"), dp("traj"), dp("ctrl"), dp("poly_used"), dp("assign_out"), dp("res"), dp("pos_out"),
            C.c_void_p(stream_ptr) if stream_ptr else None))

    # -- K1 alone -----------------------------------------------------------------------------------
    def planes(self, own_pos: np.ndarray, other_pos: np.ndarray) -> np.ndarray:
        """Inter-agent planes (n_f, b) of n point pairs as the solver builds them (agent_class.cpp:1152-1205)."""
        a, b = _c(own_pos, np.float64).reshape(-1, 3), _c(other_pos, np.float64).reshape(-1, 3)
        out = np.zeros((a.shape[0], 4))
        self._check(self.lib.hdsm_planes(self._h, a.shape[0], _p(a, C.c_double), _p(b, C.c_double), _p(out, C.c_double)))
        return out

    # -- NCCL exchange ----------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.lib.hdsm_comm_unique_id(buf)
        if rc != 0:
            raise HdsmError(f"hdsm_comm_unique_id failed: {rc}")
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self.lib.hdsm_comm_init(self._h, n_ranks, rank, buf))

    def allgather_positions(self, send, recv, n_local: int, stream_ptr: int = 0):
        self._check(self.lib.hdsm_allgather_positions(self._h, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()),
                                                      int(n_local), C.c_void_p(stream_ptr) if stream_ptr else None))


    def exchange_plans(self, send_pos, recv_pos, send_valid, recv_valid, n_local: int, stream_ptr: int = 0):
        """Plan positions and "plan received" flags of all ranks in one NCCL group (hdsm_exchange_plans)."""
        self._check(self.lib.hdsm_exchange_plans(self._h, C.c_void_p(send_pos.data_ptr()), C.c_void_p(recv_pos.data_ptr()),
                                                 C.c_void_p(send_valid.data_ptr()), C.c_void_p(recv_valid.data_ptr()),
                                                 int(n_local), C.c_void_p(stream_ptr) if stream_ptr else None))

    # -- read-back / fallback / state advance on the device (agent_class.cpp:962-1019, :233-238) ---------
    def advance_device(self, t: Dict, stream_ptr: int = 0):
        """``t`` holds this step's ``traj``/``ctrl``/``res`` and the persistent ``traj_curr``, ``ctrl_curr``,
        ``have_plan``, ``x0`` and (optionally) ``prev_self_pos`` tensors; all but the first three are updated in place."""
        def dp(name):
            x = t.get(name)
            return C.c_void_p(x.data_ptr()) if x is not None else None
        n = int(t["x0"].shape[0])
        self._check(self.lib.hdsm_advance_device(self._h, n, dp("traj"), dp("ctrl"), dp("res"), dp("traj_curr"), dp("ctrl_curr"),
                                                 dp("have_plan"), dp("x0"), dp("prev_self_pos"),
                                                 C.c_void_p(stream_ptr) if stream_ptr else None))


class _OneAgentBatch:
    pass


class AgentSolver:
    """Per-agent view with the reference's member names (agent_class.hpp:283-432).

    ``traj_other_agents_`` is the snapshot of the other agents' last plans that the reference
    copies under its mutexes at agent_class.cpp:1115-1118; here it is an (n_rob, N+1, 3) array plus
    a validity mask, i.e. exactly the ``all_pos`` / ``all_valid`` arguments of the C ABI.
    """

    def __init__(self, params: Dict, agent_id: int, n_rob: int, device: int = 0, **kw):
        self.id_, self.n_rob_ = int(agent_id), int(n_rob)
        self.n_hor_, self.poly_hor_ = int(params["n_hor"]), int(params["poly_hor"])
        self.planner = TrajectoryPlanner(params, 1, n_rob, device, **kw)  # CreateGurobiModel (:32)
        self.rmax = self.planner.rmax
        self.state_curr_ = np.zeros(9)
        self.state_ini_ = np.zeros(9)
        self.traj_ref_curr_ = np.zeros((self.n_hor_ + 1, 6))
        self.poly_const_vec_ = []  # list of (A (R,3), b (R,))
        self.traj_curr_ = np.zeros((0, 9))
        self.control_curr_ = np.zeros((0, 3))
        self.traj_other_agents_ = np.zeros((n_rob, self.n_hor_ + 1, 3))
        self.traj_other_valid_ = np.zeros(n_rob, np.uint8)
        self.poly_used_idx_ = np.zeros(self.poly_hor_, bool)
        self.optimization_failed_ = False
        self.last_result = None
        self._snapshot = None

    def GenerateTimeAwareSafeCorridor(self):
        """Takes the snapshot of neighbour plans and of the own previous plan (:1096-1118); the planes
        themselves are assembled on the device inside SolveOptimizationProblem."""
        N = self.n_hor_
        prev = self.traj_curr_[:, :3].copy() if len(self.traj_curr_) else np.repeat(self.state_ini_[None, :3], N + 1, 0)
        valid = self.traj_other_valid_.copy()
        valid[self.id_] = 0  # own slot is empty in the reference (:1132-1134)
        self._snapshot = (prev, self.traj_other_agents_.copy(), valid)

    def SolveOptimizationProblem(self):
        from .scenarios import pack_polys
        if self._snapshot is None:
            self.GenerateTimeAwareSafeCorridor()
        prev, allp, valid = self._snapshot
        self._snapshot = None
        N, P = self.n_hor_, self.poly_hor_
        b = _OneAgentBatch()
        b.global_id, b.nbr_begin, b.nbr_end = np.array([self.id_]), None, None
        b.x0 = self.state_curr_[None, :].copy()
        b.ref = np.asarray(self.traj_ref_curr_, float)[None, :N, :6].copy()
        b.poly_A, b.poly_b, b.poly_rows = pack_polys([self.poly_const_vec_], P, self.rmax)
        b.prev_self_pos, b.all_pos, b.all_valid = prev[None], allp, valid
        out = self.planner.solve_batch(b)
        r = out["res"][0]
        self.last_result = r
        ok = r["status"] == _lib.OPTIMAL or (r["status"] == _lib.NODE_LIMIT and np.isfinite(r["obj"]))
        self.optimization_failed_ = not ok
        if ok:  # :962-987
            self.traj_curr_, self.control_curr_ = out["traj"][0], out["ctrl"][0]
            self.poly_used_idx_ = out["poly_used"][0].astype(bool)
        elif len(self.traj_curr_):  # :1004-1018
            self.traj_curr_ = np.concatenate([self.traj_curr_[1:], self.traj_curr_[-1:]])
            self.control_curr_ = np.concatenate([self.control_curr_[1:], self.control_curr_[-1:]])
        return ok
