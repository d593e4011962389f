// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
// The path thread (JPS + distance-map planner) is out of scope; these classes have the reference's names and call
// signatures (jps3d/include/jps_planner/...) and return "no path", so Agent::GetPath compiles and is never relied on.
#ifndef HDSM_REF_SHIM_JPS_PLANNER_H_
#define HDSM_REF_SHIM_JPS_PLANNER_H_
#include <decomp_geometry/polyhedron.h>
#include <memory>
#include <vector>
namespace JPS {
struct VoxelMapUtil {
  template <class O, class D, class V> void setMap(const O&, const D&, const V&, double) {}
};
}
using JPS::VoxelMapUtil;
struct JPSPlanner3D {
  explicit JPSPlanner3D(bool) {}
  void setMapUtil(const std::shared_ptr<VoxelMapUtil>&) {}
  void updateMap() {}
  bool plan(const Vec3f&, const Vec3f&, double = 1, bool = true) { return false; }
  vec_E<Vec3f> getRawPath() const { return vec_E<Vec3f>(); }
};
#endif
