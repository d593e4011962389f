// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp)
#ifndef HDSM_REF_SHIM_GEOM_TFS_HPP_
#define HDSM_REF_SHIM_GEOM_TFS_HPP_
#include "rclcpp/rclcpp.hpp"
namespace geometry_msgs { namespace msg {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::msg::Header header; std::string child_frame_id; Transform transform; };
} }
#endif
