// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_SENSOR_PC2_HPP_
#define HDSM_REF_SHIM_SENSOR_PC2_HPP_
#include "rclcpp/rclcpp.hpp"
namespace sensor_msgs { namespace msg { struct PointCloud2 { std_msgs::msg::Header header; size_t n_points = 0; }; } }
#endif
