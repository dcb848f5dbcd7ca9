// Stand-in for <tf2_ros/transform_listener.h>: Feature.h includes it, nothing on the EKF path uses it.
#pragma once
