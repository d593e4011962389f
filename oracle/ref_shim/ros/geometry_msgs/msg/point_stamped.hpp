// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_GEOM_POINTSTAMPED_HPP_
#define HDSM_REF_SHIM_GEOM_POINTSTAMPED_HPP_
#include "geometry_msgs/msg/point.hpp"
#include "rclcpp/rclcpp.hpp"
namespace geometry_msgs { namespace msg { struct PointStamped { typedef std::shared_ptr<PointStamped> SharedPtr; std_msgs::msg::Header header; Point point; }; } }
#endif
