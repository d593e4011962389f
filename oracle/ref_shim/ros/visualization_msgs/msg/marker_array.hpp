// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp)
#ifndef HDSM_REF_SHIM_VIS_MARKER_HPP_
#define HDSM_REF_SHIM_VIS_MARKER_HPP_
#include "geometry_msgs/msg/point.hpp"
#include "geometry_msgs/msg/pose_stamped.hpp"
#include "geometry_msgs/msg/transform_stamped.hpp"
namespace visualization_msgs { namespace msg {
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, LINE_STRIP = 4, LINE_LIST = 5, ADD = 0 };
  std_msgs::msg::Header header; std::string ns; int id = 0, type = 0, action = 0;
  geometry_msgs::msg::Pose pose; geometry_msgs::msg::Vector3 scale; ColorRGBA color; std::vector<geometry_msgs::msg::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
} }
#endif
