// Stand-in for <ros/ros.h>: see ../ros_shim.hpp (test infrastructure only).
#include "../ros_shim.hpp"
