// Stand-in for <opencv2/core/core.hpp>: see cv_shim.hpp (test infrastructure only).
#include "../../cv_shim.hpp"
