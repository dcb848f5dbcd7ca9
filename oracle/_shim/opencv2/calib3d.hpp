// Stand-in for <opencv2/calib3d.hpp>: see cv_shim.hpp (test infrastructure only).
#include "../cv_shim.hpp"
