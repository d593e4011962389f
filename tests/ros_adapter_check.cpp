// Compiles include/hdsm_ros_adapter.hpp against stand-ins of the ROS2 message / geometry classes (same member names
// and shapes as the generated ones; ROS2 itself is not installed here) and runs the shim of INTEGRATION.md sections 3-4
// as a plain C++ caller of the C ABI - the code a maintainer pastes into agent_class.cpp, minus rclcpp.
//   ros_adapter_check pack    : packing checks only (no device needed)
//   ros_adapter_check solve   : additionally three closed-loop planning steps of two agents through hdsm_solve_batch
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "hdsm.h"
#include "hdsm_ros_adapter.hpp"

#define CHECK(c)                                                          \
  do {                                                                    \
    if (!(c)) {                                                           \
      std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c);   \
      return 1;                                                           \
    }                                                                     \
  } while (0)

namespace standin {  // multi_agent_planner_msgs/msg/{State,Trajectory}.msg, env_builder_msgs/msg/VoxelGrid.msg
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
struct State { std::vector<double> position, velocity, acceleration; };
struct Trajectory { Time stamp; double dt = 0; std::vector<State> states; double yaw = 0; };
struct VoxelGrid { std::array<double, 3> origin{}; std::array<uint32_t, 3> dimension{}; double voxel_size = 0; std::vector<int8_t> data; };
// LinearConstraint3D: A_ is an Eigen matrix with dynamic rows (column-major), b_ a dynamic vector (polyhedron.h:98-147)
struct Mat { int rows = 0; std::vector<double> v; double operator()(int r, int c) const { return v[(size_t)c * rows + r]; } };
struct Vec { std::vector<double> v; size_t size() const { return v.size(); } double operator()(int i) const { return v[i]; } };
struct LinearConstraint3D { Mat A_; Vec b_; };
LinearConstraint3D box(double cx, double cy, double cz, double half) {  // the six faces GetPolyOcta3D gives in free space
  LinearConstraint3D lc;
  lc.A_.rows = 6, lc.A_.v.assign(18, 0.0), lc.b_.v.assign(6, 0.0);
  const double c[3] = {cx, cy, cz};
  for (int a = 0; a < 3; ++a) {
    lc.A_.v[(size_t)a * 6 + 2 * a] = 1.0, lc.b_.v[2 * a] = c[a] + half;
    lc.A_.v[(size_t)a * 6 + 2 * a + 1] = -1.0, lc.b_.v[2 * a + 1] = -(c[a] - half);
  }
  return lc;
}
}  // namespace standin

// The planner-side members the shim touches (agent_class.hpp:407-432 and the state it reads), named as in the reference.
struct AgentShim {
  int id_ = 0, n_rob_ = 1, n_hor_ = 10, poly_hor_ = 4;
  double dt_ = 0.1, yaw_ = 0;
  std::vector<double> state_ini_, state_curr_;
  std::vector<std::vector<double>> traj_curr_, control_curr_, traj_ref_curr_;
  std::vector<standin::LinearConstraint3D> poly_const_vec_;
  std::vector<standin::Trajectory> traj_other_agents_;
  std::vector<bool> poly_used_idx_;
  bool optimization_failed_ = false;
  hdsm_handle* hdsm_ = nullptr;
  std::vector<double> all_pos_, prev_self_;
  std::vector<uint8_t> all_valid_;
  hdsm_result last_{};

  int CreateGurobiModel() {  // INTEGRATION.md section 3 (agent_agile_config.yaml values)
    hdsm_params p{};
    p.n_hor = n_hor_, p.poly_hor = poly_hor_, p.max_rows_per_poly = 18, p.rk4 = 0, p.prune = 1;
    p.dt = dt_, p.r_u = 0.01;
    for (int i = 0; i < 6; i++) p.r_x[i] = p.r_n[i] = i < 3 ? 100.0 : 1.0;
    p.max_vel = 20, p.max_jerk = 60, p.min_acc_xy = -15, p.max_acc_xy = 15, p.min_acc_z = -15, p.max_acc_z = 15;
    p.drone_radius = 0.25, p.drone_z_offset = 0.25, p.tilt = 0.1;
    return hdsm_create(&p, 1, n_rob_, 0, &hdsm_);
  }
  void GenerateTimeAwareSafeCorridor() {  // section 4: the snapshot is all that is left on the host
    hdsm_ros::pack_neighbour_table(traj_other_agents_, n_hor_, all_pos_, all_valid_);
    hdsm_ros::pack_prev_self(traj_curr_, state_ini_, n_hor_, prev_self_);
  }
  int SolveOptimizationProblem() {
    const int N = n_hor_, P = poly_hor_, R = 18;
    std::vector<double> ref, A, b, traj((N + 1) * 9), ctrl(N * 3);
    std::vector<int32_t> rows, assign(N);
    std::vector<uint8_t> used(P);
    hdsm_ros::pack_reference(traj_ref_curr_, N, ref);
    if (!hdsm_ros::pack_polytopes(poly_const_vec_, P, R, A, b, rows)) return HDSM_ERR_CAPACITY;
    int32_t gid = id_;
    const int rc = hdsm_solve_batch(hdsm_, 1, &gid, nullptr, nullptr, state_curr_.data(), ref.data(), A.data(), b.data(), rows.data(),
                                    prev_self_.data(), all_pos_.data(), all_valid_.data(), n_rob_, nullptr, traj.data(), ctrl.data(),
                                    used.data(), assign.data(), &last_);
    optimization_failed_ = rc != HDSM_OK || !(last_.status == HDSM_OPTIMAL || (last_.status == HDSM_NODE_LIMIT && std::isfinite(last_.obj)));
    if (!optimization_failed_)
      hdsm_ros::unpack_plan(traj.data(), ctrl.data(), used.data(), N, P, traj_curr_, control_curr_, poly_used_idx_);
    else
      hdsm_ros::shift_plan_on_failure(traj_curr_, control_curr_);  // :997-1019
    return rc;
  }
  standin::Trajectory PublishTrajectoryFull() const {
    standin::Trajectory m;
    hdsm_ros::fill_trajectory_msg(traj_curr_, dt_, yaw_, m);
    return m;
  }
};

static int check_packing() {
  // a plan message -> one row of the neighbour table, and back through fill_trajectory_msg
  std::vector<std::vector<double>> plan(11, std::vector<double>(9));
  for (int k = 0; k <= 10; ++k)
    for (int j = 0; j < 9; ++j) plan[k][j] = 100.0 * k + j + 0.25;
  standin::Trajectory msg;
  hdsm_ros::fill_trajectory_msg(plan, 0.1, 0.7, msg);
  CHECK(msg.states.size() == 11 && msg.dt == 0.1 && msg.yaw == 0.7);
  CHECK(msg.states[3].position[1] == 301.25 && msg.states[3].velocity[0] == 303.25 && msg.states[10].acceleration[2] == 1008.25);
  std::vector<standin::Trajectory> others(3);
  others[1] = msg;                    // slot 0 = self (empty), slot 2 = never received
  others[2].states.resize(4);        // a plan of another horizon is not usable
  std::vector<double> all_pos;
  std::vector<uint8_t> valid;
  hdsm_ros::pack_neighbour_table(others, 10, all_pos, valid);
  CHECK(all_pos.size() == 3 * 33 && valid[0] == 0 && valid[1] == 1 && valid[2] == 0);
  CHECK(all_pos[33 + 3 * 4 + 2] == 402.25 && all_pos[0] == 0.0);
  // own previous plan: state_ini_ before the first solve, traj_curr_ afterwards
  std::vector<double> prev;
  hdsm_ros::pack_prev_self({}, {1.0, 2.0, 3.0, 0, 0, 0, 0, 0, 0}, 10, prev);
  CHECK(prev.size() == 33 && prev[30] == 1.0 && prev[32] == 3.0);
  hdsm_ros::pack_prev_self(plan, {1.0, 2.0, 3.0}, 10, prev);
  CHECK(prev[3 * 7 + 1] == 701.25);
  // polytopes: column-major A_ -> row-major [P][R][3]; only the first poly_hor are taken; too many rows are refused
  std::vector<standin::LinearConstraint3D> polys = {standin::box(0, 0, 1.5, 2.25), standin::box(3, 0, 1.5, 2.25)};
  std::vector<double> A, b;
  std::vector<int32_t> rows;
  CHECK(hdsm_ros::pack_polytopes(polys, 4, 18, A, b, rows));
  CHECK(rows[0] == 6 && rows[1] == 6 && rows[2] == 0 && A.size() == 4 * 18 * 3 && b.size() == 4 * 18);
  CHECK(A[(0 * 18 + 0) * 3 + 0] == 1.0 && A[(0 * 18 + 1) * 3 + 0] == -1.0 && A[(0 * 18 + 4) * 3 + 2] == 1.0 && b[18 + 0] == 5.25 && b[5] == 0.75);
  CHECK(hdsm_ros::pack_polytopes(polys, 1, 18, A, b, rows) && rows.size() == 1);
  CHECK(!hdsm_ros::pack_polytopes(polys, 4, 5, A, b, rows));
  // reference rows and the failure fallback
  std::vector<double> ref;
  hdsm_ros::pack_reference(plan, 10, ref);
  CHECK(ref.size() == 60 && ref[6 * 9 + 5] == 905.25);
  std::vector<std::vector<double>> tc = plan, cc(10, std::vector<double>(3, 1.0));
  cc[9][0] = 9.0;
  hdsm_ros::shift_plan_on_failure(tc, cc);
  CHECK(tc.size() == 11 && tc[0][0] == 100.25 && tc[9][0] == 1000.25 && tc[10][0] == 1000.25 && cc.size() == 10 && cc[8][0] == 9.0 && cc[9][0] == 9.0);
  // results -> members
  std::vector<double> traj(99), ctrl(30);
  for (int i = 0; i < 99; ++i) traj[i] = i;
  for (int i = 0; i < 30; ++i) ctrl[i] = -i;
  const uint8_t used[4] = {1, 0, 1, 0};
  std::vector<bool> pu;
  hdsm_ros::unpack_plan(traj.data(), ctrl.data(), used, 10, 4, tc, cc, pu);
  CHECK(tc[10][8] == 98 && cc[9][2] == -29 && pu[0] && !pu[1] && pu[2] && !pu[3]);
  // voxel grid messages
  standin::VoxelGrid g;
  const int32_t dim[3] = {4, 3, 2};
  const double org[3] = {-1.5, 2.0, 0.0};
  std::vector<int8_t> data(24);
  for (int i = 0; i < 24; ++i) data[i] = (int8_t)(i % 3 == 0 ? 100 : (i % 3 == 1 ? 0 : -1));
  hdsm_ros::fill_grid_msg(data.data(), dim, org, 0.3, g);
  const hdsm_ros::GridView v = hdsm_ros::view_grid_msg(g);
  CHECK(v.data && v.voxels() == 24 && v.dim[1] == 3 && v.origin[0] == -1.5 && v.voxel_size == 0.3 && v.data[3] == 100 && v.data[5] == -1);
  g.data.pop_back();
  CHECK(hdsm_ros::view_grid_msg(g).data == nullptr);
  return 0;
}

// two agents flying towards each other's start in free space, three planning steps with message exchange in between
static int check_solve() {
  AgentShim ag[2];
  const double start[2][3] = {{0.0, 0.0, 1.5}, {6.0, 0.4, 1.5}};
  for (int i = 0; i < 2; ++i) {
    ag[i].id_ = i, ag[i].n_rob_ = 2;
    ag[i].state_ini_ = {start[i][0], start[i][1], start[i][2], 0, 0, 0, 0, 0, 0};
    ag[i].state_curr_ = ag[i].state_ini_;
    ag[i].traj_other_agents_.assign(2, standin::Trajectory());
    const int rc = ag[i].CreateGurobiModel();
    if (rc != HDSM_OK) {
      std::fprintf(stderr, "hdsm_create failed (%d)\n", rc);
      return 2;
    }
  }
  for (int step = 0; step < 3; ++step) {
    for (int i = 0; i < 2; ++i) {
      AgentShim& a = ag[i];
      const double dir = i == 0 ? 1.0 : -1.0, vel = 4.5;
      a.traj_ref_curr_.assign(11, std::vector<double>(6, 0.0));
      for (int k = 0; k <= 10; ++k) {  // straight-line reference towards the other side
        a.traj_ref_curr_[k][0] = a.state_curr_[0] + dir * vel * a.dt_ * (k + 1);
        a.traj_ref_curr_[k][1] = a.state_curr_[1], a.traj_ref_curr_[k][2] = a.state_curr_[2], a.traj_ref_curr_[k][3] = dir * vel;
      }
      a.poly_const_vec_ = {standin::box(a.state_curr_[0], a.state_curr_[1], 1.5, 2.25), standin::box(a.state_curr_[0] + dir * 3.0, a.state_curr_[1], 1.5, 2.25)};
      a.GenerateTimeAwareSafeCorridor();
      const int rc = a.SolveOptimizationProblem();
      CHECK(rc == HDSM_OK);
      CHECK(!a.optimization_failed_ && a.last_.status == HDSM_OPTIMAL && a.last_.kkt_res <= 1e-6);
      for (int j = 0; j < 9; ++j) CHECK(a.traj_curr_[0][j] == a.state_curr_[j]);            // x0 reproduced (:886-889)
      for (int j = 3; j < 9; ++j) CHECK(a.traj_curr_[10][j] == 0.0);                        // terminal v = a = 0 (:2078-2081)
      CHECK(a.poly_used_idx_[0] && !a.poly_used_idx_[2] && !a.poly_used_idx_[3]);
      CHECK(dir * (a.traj_curr_[10][0] - a.state_curr_[0]) > 0.5);                           // it moves towards its goal
    }
    for (int i = 0; i < 2; ++i) {  // "publish" and "receive", then advance one step (state_curr_ = traj_curr_[step_plan_])
      ag[1 - i].traj_other_agents_[i] = ag[i].PublishTrajectoryFull();
      ag[i].state_curr_ = ag[i].traj_curr_[1];
    }
  }
  // after the exchange the separating planes were active: the plans never come closer than the safety ellipsoid allows
  double dmin = 1e9;
  for (int k = 0; k <= 10; ++k) {
    double d2 = 0;
    for (int c = 0; c < 3; ++c) d2 += std::pow(ag[0].traj_curr_[k][c] - ag[1].traj_curr_[k][c], 2);
    dmin = std::fmin(dmin, std::sqrt(d2));
  }
  CHECK(dmin > 0.45);
  std::printf("solve ok: 2 agents x 3 steps, final objective %.6f / %.6f, min separation %.3f m\n", ag[0].last_.obj, ag[1].last_.obj, dmin);
  for (int i = 0; i < 2; ++i) hdsm_destroy(ag[i].hdsm_);
  return 0;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "pack";
  if (check_packing() != 0) return 1;
  std::printf("packing ok\n");
  if (mode == "solve") return check_solve();
  // without a device the shim must see a clean error code, not a crash
  AgentShim a;
  a.state_ini_ = a.state_curr_ = std::vector<double>(9, 0.0);
  const int rc = a.CreateGurobiModel();
  std::printf("hdsm_create returned %d\n", rc);
  if (rc == HDSM_OK) hdsm_destroy(a.hdsm_);
  return 0;
}
