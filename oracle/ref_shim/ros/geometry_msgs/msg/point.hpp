// TEST INFRASTRUCTURE - stand-in (see rclcpp/rclcpp.hpp)
#ifndef HDSM_REF_SHIM_GEOM_POINT_HPP_
#define HDSM_REF_SHIM_GEOM_POINT_HPP_
namespace geometry_msgs { namespace msg { struct Point { double x = 0, y = 0, z = 0; }; } }
#endif
