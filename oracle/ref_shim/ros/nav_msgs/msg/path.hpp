// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_NAV_PATH_HPP_
#define HDSM_REF_SHIM_NAV_PATH_HPP_
#include "geometry_msgs/msg/pose_stamped.hpp"
namespace nav_msgs { namespace msg { struct Path { std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses; }; } }
#endif
