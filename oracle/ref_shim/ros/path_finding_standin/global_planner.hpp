// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
// path_finding_util::GlobalPlanner (path_finding_util/include/global_planner.hpp) belongs to the out-of-scope path thread.
#ifndef HDSM_REF_SHIM_GLOBAL_PLANNER_HPP_
#define HDSM_REF_SHIM_GLOBAL_PLANNER_HPP_
#include <vector>
#include "voxel_grid.hpp"
namespace path_finding_util {
struct GlobalPlanner {
  std::vector<Eigen::Vector3d> PlanJPS(const Eigen::Vector3d&, const Eigen::Vector3d&, ::voxel_grid_util::VoxelGrid*) { return std::vector<Eigen::Vector3d>(); }
};
}
#endif
