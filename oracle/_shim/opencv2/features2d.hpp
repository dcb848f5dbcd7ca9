// Stand-in for <opencv2/features2d.hpp>: see cv_shim.hpp (test infrastructure only).
#include "../cv_shim.hpp"
