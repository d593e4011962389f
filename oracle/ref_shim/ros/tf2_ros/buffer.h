// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp)
#ifndef HDSM_REF_SHIM_TF2_BUFFER_H_
#define HDSM_REF_SHIM_TF2_BUFFER_H_
#include "rclcpp/rclcpp.hpp"
namespace tf2_ros { struct Buffer { explicit Buffer(rclcpp::Clock::SharedPtr) {} }; }
#endif
