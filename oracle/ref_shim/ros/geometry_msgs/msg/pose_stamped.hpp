// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_GEOM_POSESTAMPED_HPP_
#define HDSM_REF_SHIM_GEOM_POSESTAMPED_HPP_
#include "geometry_msgs/msg/point.hpp"
#include "geometry_msgs/msg/transform_stamped.hpp"
namespace geometry_msgs { namespace msg {
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
} }
#endif
