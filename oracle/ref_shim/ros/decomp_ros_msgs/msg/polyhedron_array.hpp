// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_DECOMP_MSG_HPP_
#define HDSM_REF_SHIM_DECOMP_MSG_HPP_
#include "geometry_msgs/msg/point.hpp"
#include "rclcpp/rclcpp.hpp"
namespace decomp_ros_msgs { namespace msg {
struct Polyhedron { std::vector<geometry_msgs::msg::Point> points, normals; };
struct PolyhedronArray { std_msgs::msg::Header header; std::vector<Polyhedron> polyhedrons; };
} }
#endif
