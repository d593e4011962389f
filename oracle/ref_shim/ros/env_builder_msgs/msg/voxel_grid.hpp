// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp): just what multi_agent_planner/src/agent_class.cpp needs to compile unmodified
#include "env_builder_msgs/msg/voxel_grid_stamped.hpp"
