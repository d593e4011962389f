// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
// fields of multi_agent_planner_msgs/msg/Trajectory.msg
#ifndef HDSM_REF_SHIM_MAP_TRAJ_HPP_
#define HDSM_REF_SHIM_MAP_TRAJ_HPP_
#include "multi_agent_planner_msgs/msg/state.hpp"
#include "rclcpp/rclcpp.hpp"
namespace multi_agent_planner_msgs { namespace msg {
struct Trajectory { typedef std::shared_ptr<Trajectory> SharedPtr; builtin_interfaces::msg::Time stamp; double dt = 0; std::vector<State> states; double yaw = 0; };
} }
#endif
