// TEST INFRASTRUCTURE - C entry point around the reference's own map-builder node.
//
// Linked with /root/reference/mapping_util/src/map_builder.cpp, path_finding_util/src/path_tools.cpp and
// voxel_grid_util/src/{voxel_grid,raycast}.cpp, all compiled UNMODIFIED from where they lie, against the stand-ins
// in oracle/ref_shim/ (Eigen 3-vectors / 3x3 / quaternion; rclcpp, tf2, message classes with the .msg files' fields)
// into oracle/_ref/libref_map.so.  Used by tests/ to pin oracle/sense_oracle.c (frame, crop, RaycastAndClear,
// ClearLine, MergeVoxelGrids, ClearVoxelsCenter: map_builder.cpp:80-205, 242-329, 367-447) and the whole published
// grid (+ SetUncertainToUnknown :331-365, InflateObstacles, CreatePotentialField) of oracle/map_oracle.c.
//
// The node's members are private; this file - not the reference - opens them with the macro below after every
// standard header it needs has been included.  map_builder.cpp itself is a separate, untouched translation unit.
#include <stdint.h>
#include <string.h>

#include <array>
#include <chrono>
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#define private public
#include "map_builder.hpp"
#undef private

extern "C" {

// One call of MapBuilder::EnvironmentVoxelGridCallback on a freshly constructed node.
//   range[3], free_grid, inflation / potential parameters, limited_fov, fov_x, fov_y : the ROS parameters
//   env / dim_env / origin_env / voxel : the environment message;  pos : pos_curr_;  rot : rot_mat_cam_ (row-major)
//   old_grid / old_origin (NULL = first update) : voxel_grid_curr_ before the call (dimension = floor(range / voxel))
//   curr_out / curr_origin : voxel_grid_curr_ after the call;  pub_out : data of the published message
//   times_ms (may be NULL) [3] : the node's own wall-clock timers raycast_comp_time_, merge_comp_time_, tot_comp_time_ (:170-236)
// Returns the number of voxels of the local grid, or -1 when the node did not publish.
int ref_map_update(const double* range, int free_grid, double inflation_dist, double potential_dist, double potential_pow, int limited_fov,
                   double fov_x, double fov_y, const int8_t* env, const int32_t* dim_env, const double* origin_env, double voxel,
                   const double* pos, const double* rot, const int8_t* old_grid, const double* old_origin, int8_t* curr_out,
                   double* curr_origin, int8_t* pub_out, double* times_ms) {
  auto& ov = rclcpp::parameter_overrides();
  ov.clear();
  ov["voxel_grid_range"] = rclcpp::Parameter(std::vector<double>(range, range + 3));
  ov["free_grid"] = rclcpp::Parameter(free_grid != 0);
  ov["inflation_dist"] = rclcpp::Parameter(inflation_dist);
  ov["potential_dist"] = rclcpp::Parameter(potential_dist);
  ov["potential_pow"] = rclcpp::Parameter(potential_pow);
  ov["limited_fov"] = rclcpp::Parameter(limited_fov != 0);
  ov["fov_x"] = rclcpp::Parameter(fov_x);
  ov["fov_y"] = rclcpp::Parameter(fov_y);
  std::streambuf* keep = std::cout.rdbuf();  // the node logs to stdout
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());
  int n_out = -1;
  {
    mapping_util::MapBuilder mb;
    for (int a = 0; a < 3; ++a) mb.pos_curr_[a] = pos[a];
    if (rot)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) mb.rot_mat_cam_(r, c) = rot[3 * r + c];
    mb.first_transform_received_ = true;
    Eigen::Vector3i dim((int)std::floor(range[0] / voxel), (int)std::floor(range[1] / voxel), (int)std::floor(range[2] / voxel));
    const size_t n = (size_t)dim(0) * dim(1) * dim(2);
    if (old_grid) {
      std::vector<voxel_grid_util::voxel_data_type> data(old_grid, old_grid + n);
      Eigen::Vector3d oo(old_origin[0], old_origin[1], old_origin[2]);
      mb.voxel_grid_curr_ = voxel_grid_util::VoxelGrid(oo, dim, voxel, data);
    }
    auto msg = std::make_shared<env_builder_msgs::msg::VoxelGridStamped>();
    for (int a = 0; a < 3; ++a) msg->voxel_grid.origin[a] = origin_env[a], msg->voxel_grid.dimension[a] = (uint32_t)dim_env[a];
    msg->voxel_grid.voxel_size = voxel;
    msg->voxel_grid.data.assign(env, env + (size_t)dim_env[0] * dim_env[1] * dim_env[2]);
    mb.EnvironmentVoxelGridCallback(msg);
    if (mb.voxel_grid_pub_->count == 1) {
      const auto cur = mb.voxel_grid_curr_.GetData();
      const auto& pub = mb.voxel_grid_pub_->last.voxel_grid.data;
      if (cur.size() == n && pub.size() == n) {
        memcpy(curr_out, cur.data(), n);
        memcpy(pub_out, pub.data(), n);
        const Eigen::Vector3d o = mb.voxel_grid_curr_.GetOrigin();
        for (int a = 0; a < 3; ++a) curr_origin[a] = o(a);
        n_out = (int)n;
        if (times_ms) {
          times_ms[0] = mb.raycast_comp_time_.empty() ? 0.0 : mb.raycast_comp_time_.back();
          times_ms[1] = mb.merge_comp_time_.empty() ? 0.0 : mb.merge_comp_time_.back();
          times_ms[2] = mb.tot_comp_time_.empty() ? 0.0 : mb.tot_comp_time_.back();
        }
      }
    }
  }
  std::cout.rdbuf(keep);
  return n_out;
}
}
