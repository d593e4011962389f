// hdsm_ros_adapter.hpp - the message <-> array packing of the ROS2-side shim.
//
// Header only, no ROS2 dependency of its own: every function is a template over the message / geometry type, so the
// same code compiles inside multi_agent_planner against the generated message classes
//   multi_agent_planner_msgs::msg::Trajectory / State      (multi_agent_planner_msgs/msg/Trajectory.msg:1-11, State.msg:1-8)
//   env_builder_msgs::msg::VoxelGrid                        (env_builder_msgs/msg/VoxelGrid.msg:1-11)
//   LinearConstraint3D                                      (decomp_ros/decomp_util/include/decomp_geometry/polyhedron.h:98-147)
// and, in this repository's tests, against plain stand-in structs with the same member names (ROS2 is not installed
// here; tests/test_ros_adapter.py compiles and runs tests/ros_adapter_check.cpp).  The arrays are the layouts of
// include/hdsm.h.  INTEGRATION.md shows where each call goes in agent_class.cpp / map_builder.cpp.
#ifndef HDSM_ROS_ADAPTER_HPP_
#define HDSM_ROS_ADAPTER_HPP_
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace hdsm_ros {

// ---- neighbour table (replaces the per-(k, j) message copies of GenerateTimeAwareSafeCorridor, agent_class.cpp:1113-1134)
// One row of all_pos [(N+1)][3] from a received plan.  A plan is usable when it carries N + 1 states with 3-vectors;
// the agent's own slot holds an empty message (states.size() == 0, :1132-1134) and stays invalid.
template <class TrajectoryMsg>
inline bool pack_plan_positions(const TrajectoryMsg& msg, int n_hor, double* row) {
  if (msg.states.size() != static_cast<std::size_t>(n_hor + 1)) return false;
  for (int k = 0; k <= n_hor; ++k) {
    if (msg.states[k].position.size() < 3) return false;
    for (int c = 0; c < 3; ++c) row[3 * k + c] = msg.states[k].position[c];
  }
  return true;
}

// all_pos [n_rob][N+1][3] and all_valid [n_rob] from traj_other_agents_ (the caller holds traj_other_mtx_[j] around
// element j, or passes copies, exactly like the reference's snapshot at :1115-1118).
template <class TrajectoryMsg>
inline void pack_neighbour_table(const std::vector<TrajectoryMsg>& plans, int n_hor, std::vector<double>& all_pos,
                                 std::vector<std::uint8_t>& all_valid) {
  const std::size_t stride = static_cast<std::size_t>(n_hor + 1) * 3;
  all_pos.assign(plans.size() * stride, 0.0);
  all_valid.assign(plans.size(), 0);
  for (std::size_t j = 0; j < plans.size(); ++j) all_valid[j] = pack_plan_positions(plans[j], n_hor, &all_pos[j * stride]) ? 1 : 0;
}

// own previous plan positions [(N+1)][3]: traj_curr_ when there is one, state_ini_ repeated before the first solve (:1102-1110)
inline void pack_prev_self(const std::vector<std::vector<double>>& traj_curr, const std::vector<double>& state_ini, int n_hor,
                           std::vector<double>& prev_self_pos) {
  prev_self_pos.resize(static_cast<std::size_t>(n_hor + 1) * 3);
  for (int k = 0; k <= n_hor; ++k)
    for (int c = 0; c < 3; ++c)
      prev_self_pos[3 * k + c] = traj_curr.size() == static_cast<std::size_t>(n_hor + 1) ? traj_curr[k][c] : state_ini[c];
}

// ---- static corridor rows (poly_const_vec_, built at :1428-1437) -> poly_A [P][R][3], poly_b [P][R], poly_rows [P]
// Returns false when a polytope has more than R rows (the caller then takes the CPU path or a larger R).
template <class LinearConstraint>
inline bool pack_polytopes(const std::vector<LinearConstraint>& polys, int poly_hor, int max_rows, std::vector<double>& A,
                           std::vector<double>& b, std::vector<std::int32_t>& rows) {
  A.assign(static_cast<std::size_t>(poly_hor) * max_rows * 3, 0.0);
  b.assign(static_cast<std::size_t>(poly_hor) * max_rows, 0.0);
  rows.assign(poly_hor, 0);
  const int np = std::min<int>(poly_hor, static_cast<int>(polys.size()));  // only the first poly_hor are used (:909)
  for (int p = 0; p < np; ++p) {
    const int r_p = static_cast<int>(polys[p].b_.size());
    if (r_p > max_rows) return false;
    rows[p] = r_p;
    for (int r = 0; r < r_p; ++r) {
      for (int c = 0; c < 3; ++c) A[(static_cast<std::size_t>(p) * max_rows + r) * 3 + c] = polys[p].A_(r, c);
      b[static_cast<std::size_t>(p) * max_rows + r] = polys[p].b_(r);
    }
  }
  return true;
}

// reference trajectory traj_ref_curr_ (rows of >= 6 numbers, :1449-1553) -> ref [N][6] (only the first N rows are used, :870-883)
inline void pack_reference(const std::vector<std::vector<double>>& traj_ref, int n_hor, std::vector<double>& ref) {
  ref.assign(static_cast<std::size_t>(n_hor) * 6, 0.0);
  for (int i = 0; i < n_hor && i < static_cast<int>(traj_ref.size()); ++i)
    for (int j = 0; j < 6 && j < static_cast<int>(traj_ref[i].size()); ++j) ref[6 * i + j] = traj_ref[i][j];
}

// ---- results -> the node's members (read-back of SolveOptimizationProblem, :962-987)
inline void unpack_plan(const double* traj, const double* ctrl, const std::uint8_t* poly_used, int n_hor, int poly_hor,
                        std::vector<std::vector<double>>& traj_curr, std::vector<std::vector<double>>& control_curr,
                        std::vector<bool>& poly_used_idx) {
  traj_curr.assign(n_hor + 1, std::vector<double>(9));
  control_curr.assign(n_hor, std::vector<double>(3));
  for (int k = 0; k <= n_hor; ++k)
    for (int j = 0; j < 9; ++j) traj_curr[k][j] = traj[9 * k + j];
  for (int k = 0; k < n_hor; ++k)
    for (int j = 0; j < 3; ++j) control_curr[k][j] = ctrl[3 * k + j];
  poly_used_idx.assign(poly_hor, false);
  for (int p = 0; p < poly_hor; ++p) poly_used_idx[p] = poly_used[p] != 0;
}

// the reference's fallback when the optimisation failed (:997-1019): drop the first element, repeat the last
inline void shift_plan_on_failure(std::vector<std::vector<double>>& traj_curr, std::vector<std::vector<double>>& control_curr) {
  if (traj_curr.empty() || control_curr.empty()) return;
  traj_curr.erase(traj_curr.begin());
  control_curr.erase(control_curr.begin());
  traj_curr.push_back(std::vector<double>(traj_curr.back()));
  control_curr.push_back(std::vector<double>(control_curr.back()));
}

// traj_curr_ -> Trajectory message (PublishTrajectoryFull, :645-677, the n_x = 9 branch); stamp is left to the caller (now())
template <class TrajectoryMsg>
inline void fill_trajectory_msg(const std::vector<std::vector<double>>& traj_curr, double dt, double yaw, TrajectoryMsg& msg) {
  msg.yaw = yaw;
  msg.dt = dt;
  msg.states.clear();
  msg.states.resize(traj_curr.size());
  for (std::size_t i = 0; i < traj_curr.size(); ++i) {
    msg.states[i].position = {traj_curr[i][0], traj_curr[i][1], traj_curr[i][2]};
    msg.states[i].velocity = {traj_curr[i][3], traj_curr[i][4], traj_curr[i][5]};
    msg.states[i].acceleration = {traj_curr[i][6], traj_curr[i][7], traj_curr[i][8]};
  }
}

// ---- voxel grid messages (ConvertVGMsgToVGUtil / ConvertVGUtilToVGMsg, mapping_util/src/map_builder.cpp:648-690)
struct GridView {
  const std::int8_t* data;
  std::int32_t dim[3];
  double origin[3];
  double voxel_size;
  std::size_t voxels() const { return static_cast<std::size_t>(dim[0]) * dim[1] * dim[2]; }
};

// a view of the message's grid in the argument form of hdsm_sense_batch / hdsm_corridor_batch / hdsm_reftraj_batch;
// data == nullptr when the message is inconsistent (data.size() != product of the dimensions)
template <class VoxelGridMsg>
inline GridView view_grid_msg(const VoxelGridMsg& msg) {
  GridView v{};
  for (int a = 0; a < 3; ++a) v.dim[a] = static_cast<std::int32_t>(msg.dimension[a]), v.origin[a] = msg.origin[a];
  v.voxel_size = msg.voxel_size;
  v.data = msg.data.size() == v.voxels() ? reinterpret_cast<const std::int8_t*>(msg.data.data()) : nullptr;
  return v;
}

template <class VoxelGridMsg>
inline void fill_grid_msg(const std::int8_t* data, const std::int32_t dim[3], const double origin[3], double voxel_size, VoxelGridMsg& msg) {
  for (int a = 0; a < 3; ++a) msg.origin[a] = origin[a], msg.dimension[a] = static_cast<std::uint32_t>(dim[a]);
  msg.voxel_size = voxel_size;
  msg.data.assign(data, data + static_cast<std::size_t>(dim[0]) * dim[1] * dim[2]);
}

}  // namespace hdsm_ros
#endif  // HDSM_ROS_ADAPTER_HPP_
