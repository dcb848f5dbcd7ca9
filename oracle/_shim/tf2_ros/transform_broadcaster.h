// Stand-in for <tf2_ros/transform_broadcaster.h>: Feature.h includes it, nothing on the EKF path uses it.
#pragma once
