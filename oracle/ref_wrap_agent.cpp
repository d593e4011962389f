// TEST INFRASTRUCTURE - C entry points around the reference's own planner node.
//
// Linked with /root/reference/multi_agent_planner/src/agent_class.cpp (and mapping_util/src/map_builder.cpp,
// path_finding_util/src/path_tools.cpp, voxel_grid_util/src/{voxel_grid,raycast}.cpp, convex_decomp_util/src/
// convex_decomp.cpp), all compiled UNMODIFIED from where they lie, against the stand-ins in oracle/ref_shim/: Eigen
// vectors, rclcpp / tf2 / message classes, jps3d planner classes that find no path (the path thread is out of scope), and
// the RECORDING stand-in for the Gurobi C++ API (oracle/ref_shim/gurobi/gurobi_c++.h).  Result: oracle/_ref/libref_agent.so.
//
// ref_agent_step runs the reference's own Agent::GenerateTimeAwareSafeCorridor (agent_class.cpp:1086-1215) and
// Agent::SolveOptimizationProblem (:858-1023) on given inputs and returns (a) the inter-agent planes it appended to every
// polytope and (b) the optimisation MODEL it handed to Gurobi: objective, variable bounds, dynamics rows, indicator rows,
// one-hot rows.  That pins the optimisation's DATA of oracle/hdsm_oracle.py against the reference; what Gurobi then does
// with the model is closed source and stays unpinned.
//
// The node's members are private; this file - not the reference - opens them with the macro below after every standard
// header it needs has been included.  agent_class.cpp itself is a separate, untouched translation unit.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#include "agent_class.hpp"
#undef private

using multi_agent_planner::Agent;

extern "C" {

// Parameters as the node's ROS parameters (agent_class.cpp:2190-2248).  The node object is never destroyed (its two worker
// threads are plain std::thread members that the reference never joins); they are let through their start-up waits and,
// with rclcpp::ok() false in the stand-in, return at once.
void* ref_agent_create(int n_rob, int id, int n_hor, int poly_hor, double dt, int rk4, double r_u, const double* r_x, const double* r_n,
                       double max_vel, double min_acc_xy, double max_acc_xy, double min_acc_z, double max_acc_z, double max_jerk,
                       double drone_radius, double drone_z_offset, const double* drag, const double* state_ini) {
  auto& ov = rclcpp::parameter_overrides();
  ov.clear();
  ov["n_rob"] = rclcpp::Parameter(n_rob), ov["id"] = rclcpp::Parameter(id), ov["n_hor"] = rclcpp::Parameter(n_hor);
  ov["poly_hor"] = rclcpp::Parameter(poly_hor), ov["dt"] = rclcpp::Parameter(dt), ov["rk4"] = rclcpp::Parameter(rk4 != 0);
  ov["r_u"] = rclcpp::Parameter(r_u);
  ov["r_x"] = rclcpp::Parameter(std::vector<double>(r_x, r_x + 9)), ov["r_n"] = rclcpp::Parameter(std::vector<double>(r_n, r_n + 9));
  ov["max_vel"] = rclcpp::Parameter(max_vel), ov["min_acc_xy"] = rclcpp::Parameter(min_acc_xy), ov["max_acc_xy"] = rclcpp::Parameter(max_acc_xy);
  ov["min_acc_z"] = rclcpp::Parameter(min_acc_z), ov["max_acc_z"] = rclcpp::Parameter(max_acc_z), ov["max_jerk"] = rclcpp::Parameter(max_jerk);
  ov["drone_radius"] = rclcpp::Parameter(drone_radius), ov["drone_z_offset"] = rclcpp::Parameter(drone_z_offset);
  ov["drag_coeff"] = rclcpp::Parameter(std::vector<double>(drag, drag + 3));
  ov["state_ini"] = rclcpp::Parameter(std::vector<double>(state_ini, state_ini + 9));
  ov["path_planning_period"] = rclcpp::Parameter(0.001);
  ov["save_stats"] = rclcpp::Parameter(false), ov["planner_verbose"] = rclcpp::Parameter(false), ov["gurobi_verbose"] = rclcpp::Parameter(false);
  std::streambuf *keep_out = std::cout.rdbuf(), *keep_err = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf()), std::cerr.rdbuf(sink.rdbuf());
  Agent* a = new Agent();
  a->voxel_grid_ready_ = true;
  a->path_ready_ = true;
  if (a->path_planning_thread_.joinable()) a->path_planning_thread_.join();
  if (a->traj_planning_thread_.joinable()) a->traj_planning_thread_.join();
  std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
  return a;
}

int ref_agent_num_vars(void* h) { return (int)static_cast<Agent*>(h)->model_.vars.size(); }

// One GenerateTimeAwareSafeCorridor + SolveOptimizationProblem of the reference.
//   x0 [9] state_curr_;  ref [N+1][6] traj_ref_curr_;  polytopes: n_poly, rows[n_poly], A [n_poly][rmax][3], b [n_poly][rmax]
//   have_prev / prev_traj [N+1][9]: traj_curr_ (empty before the first solve);  all_pos [n_rob][N+1][3], all_valid [n_rob]
// Outputs (dense, nz = number of model variables: x (N+1) x 9, u N x 3, binaries N x P):
//   final_rows [N][n_poly]: rows of poly_const_final_vec_[k][p];  final_A [N][n_poly][fmax][3], final_b [N][n_poly][fmax]
//   obj_diag [nz], obj_lin [nz], obj_const, obj_offdiag (sum of |off-diagonal quadratic coefficients|)
//   var_lb / var_ub [nz], var_type [nz]
//   lin_*: the ACTIVE linear constraints (expr sense 0): n_lin, dense [cap_lin][nz], constant [cap_lin], sense [cap_lin]
//   ind_*: the ACTIVE indicator constraints: n_ind, dense [cap_ind][nz], constant [cap_ind], bin variable [cap_ind]
// Returns 0, or -1 when a capacity is too small.
int ref_agent_step(void* h, const double* x0, const double* ref, int n_poly, const int32_t* rows, int rmax, const double* A, const double* b,
                   int have_prev, const double* prev_traj, const double* all_pos, const uint8_t* all_valid, int fmax, int32_t* final_rows,
                   double* final_A, double* final_b, double* obj_diag, double* obj_lin, double* obj_const, double* obj_offdiag, double* var_lb,
                   double* var_ub, int8_t* var_type, int cap_lin, int32_t* n_lin, double* lin_dense, double* lin_const, int8_t* lin_sense,
                   int cap_ind, int32_t* n_ind, double* ind_dense, double* ind_const, int32_t* ind_bin, int32_t* failed) {
  Agent* a = static_cast<Agent*>(h);
  const int N = a->n_hor_, n_rob = a->n_rob_;
  std::streambuf *keep_out = std::cout.rdbuf(), *keep_err = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf()), std::cerr.rdbuf(sink.rdbuf());
  a->state_curr_.assign(x0, x0 + 9);
  a->traj_ref_curr_.assign(N + 1, std::vector<double>(6));
  for (int i = 0; i <= N; ++i)
    for (int j = 0; j < 6; ++j) a->traj_ref_curr_[i][j] = ref[6 * i + j];
  a->poly_const_vec_.clear();
  for (int p = 0; p < n_poly; ++p) {
    MatDNf<3> Am(rows[p], 3);
    VecDf bv(rows[p]);
    for (int r = 0; r < rows[p]; ++r) {
      for (int c = 0; c < 3; ++c) Am(r, c) = A[((size_t)p * rmax + r) * 3 + c];
      bv(r) = b[(size_t)p * rmax + r];
    }
    a->poly_const_vec_.push_back(LinearConstraint3D(Am, bv));
  }
  a->traj_curr_.clear();
  a->control_curr_.clear();
  if (have_prev) {
    a->traj_curr_.assign(N + 1, std::vector<double>(9));
    for (int i = 0; i <= N; ++i)
      for (int j = 0; j < 9; ++j) a->traj_curr_[i][j] = prev_traj[9 * i + j];
    a->control_curr_.assign(N, std::vector<double>(3, 0.0));
  }
  for (int j = 0; j < n_rob; ++j) {
    multi_agent_planner_msgs::msg::Trajectory t;
    if (all_valid[j] && j != a->id_) {
      t.states.resize(N + 1);
      for (int k = 0; k <= N; ++k) {
        const double* q = all_pos + ((size_t)j * (N + 1) + k) * 3;
        t.states[k].position = {q[0], q[1], q[2]};
        t.states[k].velocity = {0.0, 0.0, 0.0};
        t.states[k].acceleration = {0.0, 0.0, 0.0};
      }
    }
    a->traj_other_agents_[j] = t;
  }
  a->GenerateTimeAwareSafeCorridor();
  a->SolveOptimizationProblem();
  std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
  *failed = a->optimization_failed_ ? 1 : 0;

  int rc = 0;
  for (int k = 0; k < N; ++k)
    for (int p = 0; p < n_poly; ++p) {
      const LinearConstraint3D& lc = a->poly_const_final_vec_[k][p];
      const int r_kp = lc.A_.rows();
      final_rows[k * n_poly + p] = r_kp;
      if (r_kp > fmax) { rc = -1; continue; }
      for (int r = 0; r < r_kp; ++r) {
        for (int c = 0; c < 3; ++c) final_A[(((size_t)k * n_poly + p) * fmax + r) * 3 + c] = lc.A_(r, c);
        final_b[((size_t)k * n_poly + p) * fmax + r] = lc.b_(r);
      }
    }
  const GRBModel& m = a->model_;
  const int nz = (int)m.vars.size();
  for (int i = 0; i < nz; ++i) obj_diag[i] = 0, obj_lin[i] = m.objective.lin.coeff_of(i), var_lb[i] = m.vars[i].lb, var_ub[i] = m.vars[i].ub, var_type[i] = m.vars[i].type;
  *obj_const = m.objective.lin.constant;
  *obj_offdiag = 0;
  for (const auto& kv : m.objective.quad) {
    if (kv.first.first == kv.first.second) obj_diag[kv.first.first] = kv.second;
    else *obj_offdiag += std::fabs(kv.second);
  }
  int nl = 0;
  for (const auto& c : m.lin) {
    if (c.removed) continue;
    if (nl < cap_lin) {
      for (int i = 0; i < nz; ++i) lin_dense[(size_t)nl * nz + i] = c.expr.coeff_of(i);
      lin_const[nl] = c.expr.constant, lin_sense[nl] = c.sense;
    } else rc = -1;
    ++nl;
  }
  *n_lin = nl;
  int ni = 0;
  for (const auto& c : m.ind) {
    if (c.removed) continue;
    if (ni < cap_ind) {
      for (int i = 0; i < nz; ++i) ind_dense[(size_t)ni * nz + i] = c.expr.coeff_of(i);
      ind_const[ni] = c.expr.constant - c.rhs, ind_bin[ni] = c.bin_var;
    } else rc = -1;
    ++ni;
  }
  *n_ind = ni;
  return rc;
}
// ---- the reference's node running CLOSED LOOP on an external solver ----------------------------------------------------
// ref_agent_loop_step is one replanning iteration of the node as TrajPlanningIteration runs it (agent_class.cpp:157-258),
// minus the producers that are given as inputs: state advance state_curr_ := traj_curr_[step_plan_] (:233-238, step_plan_ = 1),
// GenerateTimeAwareSafeCorridor() (:168) and SolveOptimizationProblem() (:174) - the reference's own, unmodified code.
// Inside the latter, model_.optimize() (:959) reaches GRBModel::solver_hook of the stand-in; the hook packs the NODE'S OWN
// MEMBERS (state_curr_, traj_ref_curr_, poly_const_vec_, traj_curr_, traj_other_agents_ - exactly what the patch in
// INTEGRATION.md reads) into the C-ABI arrays, calls the solver through a function pointer with the signature of
// hdsm_solve_batch (kind 0: libhdsm, the CUDA library) or orc_solve_batch (kind 1: the C port of the oracle, for the
// CPU-only run of this test) and writes the solution into the model's variables; the reference's own read-back (:962-987)
// and, when the solver reports no usable solution (the hook throws like Gurobi does), its own fallback (:997-1019)
// consume it.  Outputs are the node's members after the iteration.
typedef int (*hdsm_solve_fn)(void*, int, const int32_t*, const int32_t*, const int32_t*, const double*, const double*, const double*,
                             const double*, const int32_t*, const double*, const double*, const uint8_t*, int, const int32_t*, double*,
                             double*, uint8_t*, int32_t*, void*);
typedef int (*orc_solve_fn)(const void*, int, const int32_t*, const int32_t*, const int32_t*, const double*, const double*, const double*,
                            const double*, const int32_t*, int, const double*, const double*, const uint8_t*, int, const int32_t*, double*,
                            double*, uint8_t*, int32_t*, void*, int);
struct LoopResult { int32_t status, iters, nodes, rows; double obj, kkt; };  // layout of hdsm_result / orc_result

int ref_agent_loop_step(void* h, const double* ref, int n_poly, const int32_t* rows, int rmax, const double* A, const double* b,
                        const double* all_pos, const uint8_t* all_valid, int kind, void* fn, void* ctx, double* x0_out, double* traj_out,
                        double* ctrl_out, uint8_t* used_out, int32_t* have_traj, int32_t* failed, int32_t* solver_status) {
  Agent* a = static_cast<Agent*>(h);
  const int N = a->n_hor_, n_rob = a->n_rob_, P = a->poly_hor_;
  std::streambuf *keep_out = std::cout.rdbuf(), *keep_err = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf()), std::cerr.rdbuf(sink.rdbuf());
  if (!a->traj_curr_.empty()) a->state_curr_ = a->traj_curr_[1];  // :233-238 with step_plan_ = 1
  for (int j = 0; j < 9; ++j) x0_out[j] = a->state_curr_[j];
  a->traj_ref_curr_.assign(N + 1, std::vector<double>(6));
  for (int i = 0; i <= N; ++i)
    for (int j = 0; j < 6; ++j) a->traj_ref_curr_[i][j] = ref[6 * i + j];
  a->poly_const_vec_.clear();
  for (int p = 0; p < n_poly; ++p) {
    MatDNf<3> Am(rows[p], 3);
    VecDf bv(rows[p]);
    for (int r = 0; r < rows[p]; ++r) {
      for (int c = 0; c < 3; ++c) Am(r, c) = A[((size_t)p * rmax + r) * 3 + c];
      bv(r) = b[(size_t)p * rmax + r];
    }
    a->poly_const_vec_.push_back(LinearConstraint3D(Am, bv));
  }
  for (int j = 0; j < n_rob; ++j) {  // what TrajectoryOtherAgentsCallback stored (:629-643); the own slot stays empty
    multi_agent_planner_msgs::msg::Trajectory t;
    if (all_valid[j] && j != a->id_) {
      t.states.resize(N + 1);
      for (int k = 0; k <= N; ++k) {
        const double* q = all_pos + ((size_t)j * (N + 1) + k) * 3;
        t.states[k].position = {q[0], q[1], q[2]};
        t.states[k].velocity = {0.0, 0.0, 0.0};
        t.states[k].acceleration = {0.0, 0.0, 0.0};
      }
    }
    a->traj_other_agents_[j] = t;
  }
  *solver_status = -1;
  GRBModel::solver_hook() = [&](GRBModel& m) {
    // pack the node's members into the C-ABI arrays (include/hdsm.h)
    std::vector<double> x0(a->state_curr_.begin(), a->state_curr_.begin() + 9), rf((size_t)N * 6), pA((size_t)P * rmax * 3, 0.0),
        pb((size_t)P * rmax, 0.0), prev((size_t)(N + 1) * 3), ap((size_t)n_rob * (N + 1) * 3, 0.0), traj((size_t)(N + 1) * 9), ctrl((size_t)N * 3);
    std::vector<int32_t> prow(P, 0), sig(N, -1);
    std::vector<uint8_t> av(n_rob, 0), used(P, 0);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 6; ++j) rf[(size_t)i * 6 + j] = a->traj_ref_curr_[i][j];
    const int np = std::min<int>(P, (int)a->poly_const_vec_.size());
    for (int p = 0; p < np; ++p) {
      const LinearConstraint3D& lc = a->poly_const_vec_[p];
      prow[p] = (int32_t)lc.A_.rows();
      for (int r = 0; r < lc.A_.rows(); ++r) {
        for (int c = 0; c < 3; ++c) pA[((size_t)p * rmax + r) * 3 + c] = lc.A_(r, c);
        pb[(size_t)p * rmax + r] = lc.b_(r);
      }
    }
    for (int k = 0; k <= N; ++k)
      for (int c = 0; c < 3; ++c) prev[(size_t)k * 3 + c] = a->traj_curr_.empty() ? a->state_ini_[c] : a->traj_curr_[k][c];  // :1103-1110
    for (int j = 0; j < n_rob; ++j) {
      const auto& st = a->traj_other_agents_[j].states;
      av[j] = !st.empty();
      for (int k = 0; k <= N && k < (int)st.size(); ++k)
        for (int c = 0; c < 3; ++c) ap[((size_t)j * (N + 1) + k) * 3 + c] = st[k].position[c];
    }
    const int32_t gid = a->id_;
    LoopResult res{};
    int rc;
    if (kind == 0)
      rc = reinterpret_cast<hdsm_solve_fn>(fn)(ctx, 1, &gid, nullptr, nullptr, x0.data(), rf.data(), pA.data(), pb.data(), prow.data(), prev.data(),
                                               ap.data(), av.data(), n_rob, nullptr, traj.data(), ctrl.data(), used.data(), sig.data(), &res);
    else
      rc = reinterpret_cast<orc_solve_fn>(fn)(ctx, 1, &gid, nullptr, nullptr, x0.data(), rf.data(), pA.data(), pb.data(), prow.data(), rmax,
                                              prev.data(), ap.data(), av.data(), n_rob, nullptr, traj.data(), ctrl.data(), used.data(), sig.data(),
                                              &res, 1);
    *solver_status = rc != 0 ? -2 : res.status;
    const bool usable = rc == 0 && (res.status == 0 || (res.status == 4 && std::isfinite(res.obj)));
    if (!usable) throw GRBException("no solution from the external solver", 10005);  // what reading X without an incumbent raises
    for (int i = 0; i <= N; ++i)
      for (int j = 0; j < 9; ++j) m.vars[a->x_grb_[i][j].index].x = traj[(size_t)i * 9 + j];
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < 3; ++j) m.vars[a->u_grb_[i][j].index].x = ctrl[(size_t)i * 3 + j];
      for (int j = 0; j < P; ++j) m.vars[a->b_grb_[i][j].index].x = sig[i] == j ? 1.0 : 0.0;
    }
  };
  try {
    a->GenerateTimeAwareSafeCorridor();
    a->SolveOptimizationProblem();
  } catch (...) {
    GRBModel::solver_hook() = nullptr;
    std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
    return -2;
  }
  GRBModel::solver_hook() = nullptr;
  std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
  *failed = a->optimization_failed_ ? 1 : 0;
  *have_traj = a->traj_curr_.empty() ? 0 : 1;
  if (!a->traj_curr_.empty()) {
    for (int i = 0; i <= N; ++i)
      for (int j = 0; j < 9; ++j) traj_out[(size_t)i * 9 + j] = a->traj_curr_[i][j];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < 3; ++j) ctrl_out[(size_t)i * 3 + j] = a->control_curr_[i][j];
  }
  for (int j = 0; j < P; ++j) used_out[j] = j < (int)a->poly_used_idx_.size() && a->poly_used_idx_[j] ? 1 : 0;
  return 0;
}

// The reference's own Agent::GenerateReferenceTrajectory (agent_class.cpp:1449-1553, with SamplePath, KeepOnlyFreeReference,
// ComputePathVelocity, GetVelocityLimit) on one agent.
//   grid / dim / origin / voxel : voxel_grid_;  path [n_path][3] : path_curr_;  prev_ref [N+1][3] (have_prev) : positions of
//   traj_ref_curr_;  increment : increment_traj_ref_;  traj [N+1][3] (have_traj) : positions of traj_curr_;  all_pos / all_valid
//   sens / vel : path_vel_min, path_vel_max, path_vel_dec, sens_dist, sens_pot, sens_other_agents (the ROS parameters)
// Outputs: ref_out [N+1][6] = traj_ref_curr_, *path_vel = path_vel_.  Returns the number of rows of traj_ref_curr_.
int ref_agent_reftraj(void* h, const int8_t* grid, const int32_t* dim, const double* origin, double voxel, const double* path, int n_path,
                      int have_prev, const double* prev_ref, int increment, int have_traj, const double* traj, const double* all_pos,
                      const uint8_t* all_valid, const double* sens_vel, double* ref_out, double* path_vel) {
  Agent* a = static_cast<Agent*>(h);
  const int N = a->n_hor_, n_rob = a->n_rob_;
  std::streambuf *keep_out = std::cout.rdbuf(), *keep_err = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf()), std::cerr.rdbuf(sink.rdbuf());
  a->path_vel_min_ = sens_vel[0], a->path_vel_max_ = sens_vel[1], a->path_vel_dec_ = sens_vel[2];
  a->sens_dist_ = sens_vel[3], a->sens_pot_ = sens_vel[4], a->sens_other_agents_ = sens_vel[5];
  Eigen::Vector3d org(origin[0], origin[1], origin[2]);
  Eigen::Vector3i d(dim[0], dim[1], dim[2]);
  std::vector<voxel_grid_util::voxel_data_type> data(grid, grid + (size_t)dim[0] * dim[1] * dim[2]);
  a->voxel_grid_ = voxel_grid_util::VoxelGrid(org, d, voxel, data);
  a->path_curr_.assign(n_path, std::vector<double>(3));
  for (int i = 0; i < n_path; ++i)
    for (int c = 0; c < 3; ++c) a->path_curr_[i][c] = path[3 * i + c];
  a->traj_ref_curr_.clear();
  if (have_prev) {
    a->traj_ref_curr_.assign(N + 1, std::vector<double>(6, 0.0));
    for (int i = 0; i <= N; ++i)
      for (int c = 0; c < 3; ++c) a->traj_ref_curr_[i][c] = prev_ref[3 * i + c];
  }
  a->reset_path_ = false;
  a->increment_traj_ref_ = increment != 0;
  a->traj_curr_.clear();
  if (have_traj) {
    a->traj_curr_.assign(N + 1, std::vector<double>(9, 0.0));
    for (int i = 0; i <= N; ++i)
      for (int c = 0; c < 3; ++c) a->traj_curr_[i][c] = traj[3 * i + c];
  }
  for (int j = 0; j < n_rob; ++j) {
    multi_agent_planner_msgs::msg::Trajectory t;
    if (all_valid[j] && j != a->id_) {
      t.states.resize(N + 1);
      for (int k = 0; k <= N; ++k) {
        const double* q = all_pos + ((size_t)j * (N + 1) + k) * 3;
        t.states[k].position = {q[0], q[1], q[2]};
        t.states[k].velocity = {0.0, 0.0, 0.0};
        t.states[k].acceleration = {0.0, 0.0, 0.0};
      }
    }
    a->traj_other_agents_[j] = t;
  }
  int rows = -1;
  try {
    a->GenerateReferenceTrajectory();
    rows = (int)a->traj_ref_curr_.size();
    for (int i = 0; i < rows && i <= N; ++i)
      for (int c = 0; c < 6; ++c) ref_out[6 * i + c] = c < (int)a->traj_ref_curr_[i].size() ? a->traj_ref_curr_[i][c] : 0.0;
    *path_vel = a->path_vel_;
  } catch (...) {
    rows = -2;
  }
  std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
  return rows;
}
// The reference's own Agent::GenerateSafeCorridor (agent_class.cpp:1236-1447: kept polytopes, walk along the path, GetPolyOcta3D /
// GetPolyOcta3DNew, A / b conversion) on one agent.
//   grid / dim / origin / voxel : voxel_grid_;  pos [3] : state_curr_;  path [n_path][3] : path_curr_;  n_it : n_it_decomp_
//   previous update (prev_n > 0): prev_rows [prev_n], prev_A [prev_n][rmax][3], prev_b [prev_n][rmax], prev_seeds [prev_n][3],
//   prev_used [prev_n] = poly_const_vec_, poly_seeds_, poly_used_idx_;  prev_traj [n_traj][3] : positions of traj_curr_
// Outputs: rows_out [cap], A_out [cap][rmax][3], b_out [cap][rmax], seeds_out [cap][3].  Returns the number of polytopes, -1 when one
// exceeds rmax rows or cap, -2 when the reference threw.
int ref_agent_corridor(void* h, const int8_t* grid, const int32_t* dim, const double* origin, double voxel, const double* pos, const double* path,
                       int n_path, int n_it, int use_cvx_new, int prev_n, const int32_t* prev_rows, int rmax, const double* prev_A,
                       const double* prev_b, const double* prev_seeds, const uint8_t* prev_used, int n_traj, const double* prev_traj, int cap,
                       int32_t* rows_out, double* A_out, double* b_out, double* seeds_out) {
  Agent* a = static_cast<Agent*>(h);
  std::streambuf *keep_out = std::cout.rdbuf(), *keep_err = std::cerr.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf()), std::cerr.rdbuf(sink.rdbuf());
  a->n_it_decomp_ = n_it, a->use_cvx_ = true, a->use_cvx_new_ = use_cvx_new != 0;
  Eigen::Vector3d org(origin[0], origin[1], origin[2]);
  Eigen::Vector3i d(dim[0], dim[1], dim[2]);
  std::vector<voxel_grid_util::voxel_data_type> data(grid, grid + (size_t)dim[0] * dim[1] * dim[2]);
  a->voxel_grid_ = voxel_grid_util::VoxelGrid(org, d, voxel, data);
  a->state_curr_.assign(9, 0.0);
  for (int c = 0; c < 3; ++c) a->state_curr_[c] = pos[c];
  a->path_curr_.assign(n_path, std::vector<double>(3));
  for (int i = 0; i < n_path; ++i)
    for (int c = 0; c < 3; ++c) a->path_curr_[i][c] = path[3 * i + c];
  a->traj_curr_.assign(n_traj, std::vector<double>(9, 0.0));
  for (int i = 0; i < n_traj; ++i)
    for (int c = 0; c < 3; ++c) a->traj_curr_[i][c] = prev_traj[3 * i + c];
  a->poly_const_vec_.clear(), a->poly_seeds_.clear(), a->poly_vec_.clear(), a->poly_used_idx_.clear();
  for (int p = 0; p < prev_n; ++p) {
    MatDNf<3> Am(prev_rows[p], 3);
    VecDf bv(prev_rows[p]);
    for (int r = 0; r < prev_rows[p]; ++r) {
      for (int c = 0; c < 3; ++c) Am(r, c) = prev_A[((size_t)p * rmax + r) * 3 + c];
      bv(r) = prev_b[(size_t)p * rmax + r];
    }
    a->poly_const_vec_.push_back(LinearConstraint3D(Am, bv));
    a->poly_seeds_.push_back({prev_seeds[3 * p], prev_seeds[3 * p + 1], prev_seeds[3 * p + 2]});
    a->poly_vec_.push_back(Polyhedron3D());
  }
  a->poly_used_idx_.assign(a->poly_hor_, false);
  for (int p = 0; p < prev_n && p < a->poly_hor_; ++p) a->poly_used_idx_[p] = prev_used[p] != 0;
  int n_out = -2;
  try {
    a->GenerateSafeCorridor();
    n_out = (int)a->poly_const_vec_.size();
    if (n_out > cap) n_out = -1;
    for (int p = 0; p < n_out; ++p) {
      const LinearConstraint3D& lc = a->poly_const_vec_[p];
      rows_out[p] = lc.A_.rows();
      if (lc.A_.rows() > rmax) { n_out = -1; break; }
      for (int r = 0; r < lc.A_.rows(); ++r) {
        for (int c = 0; c < 3; ++c) A_out[((size_t)p * rmax + r) * 3 + c] = lc.A_(r, c);
        b_out[(size_t)p * rmax + r] = lc.b_(r);
      }
      for (int c = 0; c < 3; ++c) seeds_out[3 * p + c] = a->poly_seeds_[p][c];
    }
  } catch (...) {
    n_out = -2;
  }
  std::cout.rdbuf(keep_out), std::cerr.rdbuf(keep_err);
  return n_out;
}
}
