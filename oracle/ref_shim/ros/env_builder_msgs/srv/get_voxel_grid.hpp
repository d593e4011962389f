// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
// fields of env_builder_msgs/srv/GetVoxelGrid.srv
#ifndef HDSM_REF_SHIM_ENV_SRV_HPP_
#define HDSM_REF_SHIM_ENV_SRV_HPP_
#include <array>
#include "env_builder_msgs/msg/voxel_grid_stamped.hpp"
namespace env_builder_msgs { namespace srv {
struct GetVoxelGrid {
  struct Request { std::array<double, 3> position{}; std::array<double, 3> range{}; };
  struct Response { env_builder_msgs::msg::VoxelGridStamped voxel_grid_stamped; };
};
} }
#endif
