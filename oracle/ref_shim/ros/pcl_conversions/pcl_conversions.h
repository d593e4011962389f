// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_PCL_CONV_H_
#define HDSM_REF_SHIM_PCL_CONV_H_
#include <vector>
#include "sensor_msgs/msg/point_cloud2.hpp"
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
template <class P> struct PointCloud { std::vector<P> points; void push_back(const P& p) { points.push_back(p); } size_t size() const { return points.size(); } };
template <class P> void toROSMsg(const PointCloud<P>& c, sensor_msgs::msg::PointCloud2& m) { m.n_points = c.points.size(); }
}
#endif
