// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_TF2_BROADCASTER_H_
#define HDSM_REF_SHIM_TF2_BROADCASTER_H_
#include "geometry_msgs/msg/transform_stamped.hpp"
namespace tf2_ros { struct TransformBroadcaster { template <class N> explicit TransformBroadcaster(N*) {} void sendTransform(const geometry_msgs::msg::TransformStamped& t) { last = t; } geometry_msgs::msg::TransformStamped last; }; }
#endif
