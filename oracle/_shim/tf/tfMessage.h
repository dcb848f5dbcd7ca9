// Stand-in for <tf/tfMessage.h>: Feature.h includes it, nothing on the EKF path uses it.
#pragma once
