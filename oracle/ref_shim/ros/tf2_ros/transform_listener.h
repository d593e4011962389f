// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp)
#ifndef HDSM_REF_SHIM_TF2_LISTENER_H_
#define HDSM_REF_SHIM_TF2_LISTENER_H_
#include "geometry_msgs/msg/transform_stamped.hpp"
#include "tf2_ros/buffer.h"
namespace tf2_ros { struct TransformListener { TransformListener(Buffer&, rclcpp::Node*) {} }; }
namespace tf2_msgs { namespace msg {
struct TFMessage { typedef std::shared_ptr<TFMessage> SharedPtr; std::vector<geometry_msgs::msg::TransformStamped> transforms; };
} }
#endif
