// TEST INFRASTRUCTURE ONLY — stand-in for the slice of <ros/ros.h> the reference's EKF sources use (ros::Time, ROS_* logging
// and assert macros), so the unmodified /root/reference sources compile here (oracle/Makefile target `ref`).
// Logging keeps the reference's conventions observable: DEBUG/INFO/WARN are dropped (their arguments are not evaluated),
// ERROR and FATAL are counted (ekfvio_shim::error_count / fatal_count; checkSigma's pass criterion is "no ROS_FATAL"),
// ROS_ASSERT aborts as it does in the reference's Release build (NDEBUG is not defined there, CMakeLists.txt:17).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <list>
#include <sstream>
#include <string>
#include <vector>

namespace ekfvio_shim {
inline long& error_count() { static long c = 0; return c; }
inline long& fatal_count() { static long c = 0; return c; }
}  // namespace ekfvio_shim

namespace ros {
class Time {
    double s_ = 0.0;

public:
    Time() {}
    explicit Time(double s) : s_(s) {}
    double toSec() const { return s_; }
    static Time now() { return Time(0.0); }
};
class Duration {
    double s_ = 0.0;

public:
    Duration() {}
    explicit Duration(double s) : s_(s) {}
    double toSec() const { return s_; }
};
inline Duration operator-(const Time& a, const Time& b) { return Duration(a.toSec() - b.toSec()); }
}  // namespace ros

#define EKFVIO_SHIM_NOP(...) do { } while (0)
#define ROS_DEBUG(...) EKFVIO_SHIM_NOP()
#define ROS_DEBUG_STREAM(x) EKFVIO_SHIM_NOP()
#define ROS_DEBUG_COND(c, ...) EKFVIO_SHIM_NOP()
#define ROS_INFO(...) EKFVIO_SHIM_NOP()
#define ROS_INFO_STREAM(x) EKFVIO_SHIM_NOP()
#define ROS_WARN(...) EKFVIO_SHIM_NOP()
#define ROS_WARN_STREAM(x) EKFVIO_SHIM_NOP()
#define ROS_ERROR(...) do { ++ekfvio_shim::error_count(); } while (0)
#define ROS_ERROR_STREAM(x) do { ++ekfvio_shim::error_count(); } while (0)
#define ROS_ERROR_COND(c, ...) do { if (c) ++ekfvio_shim::error_count(); } while (0)
#define ROS_FATAL(...) do { ++ekfvio_shim::fatal_count(); } while (0)
#define ROS_FATAL_STREAM(x) do { ++ekfvio_shim::fatal_count(); } while (0)
#define ROS_FATAL_STREAM_COND(c, x) do { if (c) ++ekfvio_shim::fatal_count(); } while (0)
#define ROS_ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); std::abort(); } } while (0)
