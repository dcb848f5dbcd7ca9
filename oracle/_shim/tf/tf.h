// Stand-in for <tf/tf.h>: Feature.h includes it, nothing on the EKF path uses it.
#pragma once
