// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
// fields of multi_agent_planner_msgs/msg/State.msg
#ifndef HDSM_REF_SHIM_MAP_STATE_HPP_
#define HDSM_REF_SHIM_MAP_STATE_HPP_
#include <vector>
namespace multi_agent_planner_msgs { namespace msg { struct State { std::vector<double> position, velocity, acceleration; }; } }
#endif
