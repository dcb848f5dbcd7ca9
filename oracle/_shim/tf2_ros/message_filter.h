// Stand-in for <tf2_ros/message_filter.h>: Feature.h includes it, nothing on the EKF path uses it.
#pragma once
