// oracle/ref_shim/geometry_msgs/Quaternion.h -- TEST INFRASTRUCTURE ONLY: plain stand-in for the ROS message struct that
// include/eth_mav_msgs/common.h names in inline helpers (ROS is absent here; none of those helpers is on the path).
#pragma once
namespace geometry_msgs { struct Quaternion { double x = 0, y = 0, z = 0, w = 1; }; }
