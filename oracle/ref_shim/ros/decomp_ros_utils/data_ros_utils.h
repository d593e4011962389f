// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_DECOMP_UTILS_H_
#define HDSM_REF_SHIM_DECOMP_UTILS_H_
#include <decomp_geometry/polyhedron.h>
#include "decomp_ros_msgs/msg/polyhedron_array.hpp"
namespace DecompROS {
inline decomp_ros_msgs::msg::PolyhedronArray polyhedron_array_to_ros(const vec_E<Polyhedron3D>& vs) {
  decomp_ros_msgs::msg::PolyhedronArray msg;
  for (const auto& v : vs) {
    decomp_ros_msgs::msg::Polyhedron poly;
    for (const auto& p : v.hyperplanes()) {
      geometry_msgs::msg::Point pt, n;
      pt.x = p.p_(0), pt.y = p.p_(1), pt.z = p.p_(2), n.x = p.n_(0), n.y = p.n_(1), n.z = p.n_(2);
      poly.points.push_back(pt), poly.normals.push_back(n);
    }
    msg.polyhedrons.push_back(poly);
  }
  return msg;
}
}
#endif
