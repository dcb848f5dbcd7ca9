// Stand-in for <sensor_msgs/CameraInfo.h>: only boost::array is needed by Frame.h:39.
#pragma once
#include <array>
#include <cstddef>
namespace boost { template <class T, std::size_t N> using array = std::array<T, N>; }
