// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#ifndef HDSM_REF_SHIM_DMP_PLANNER_H_
#define HDSM_REF_SHIM_DMP_PLANNER_H_
#include "jps_planner/jps_planner/jps_planner.h"
struct IterativeDMPlanner3D {
  explicit IterativeDMPlanner3D(bool) {}
  void setSearchRadius(const Vec3f&) {}
  void setCweight(double) {}
  void setMap(const std::shared_ptr<VoxelMapUtil>&, const Vec3f&) {}
  bool iterativeComputePath(const Vec3f&, const Vec3f&, const vec_E<Vec3f>&, int) { return false; }
  vec_E<Vec3f> getRawPath() const { return vec_E<Vec3f>(); }
};
#endif
