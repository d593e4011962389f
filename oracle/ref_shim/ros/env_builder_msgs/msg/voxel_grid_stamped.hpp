// TEST INFRASTRUCTURE - stand-in with the fields of env_builder_msgs/msg/VoxelGrid.msg and VoxelGridStamped.msg
#ifndef HDSM_REF_SHIM_ENV_VG_HPP_
#define HDSM_REF_SHIM_ENV_VG_HPP_
#include <array>
#include "rclcpp/rclcpp.hpp"
namespace env_builder_msgs { namespace msg {
struct VoxelGrid { std::array<double, 3> origin{}; std::array<uint32_t, 3> dimension{}; double voxel_size = 0; std::vector<int8_t> data; };
struct VoxelGridStamped { typedef std::shared_ptr<VoxelGridStamped> SharedPtr; std_msgs::msg::Header header; VoxelGrid voxel_grid; };
} }
#endif
