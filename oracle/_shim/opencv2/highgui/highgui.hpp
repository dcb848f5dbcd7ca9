// Stand-in for <opencv2/highgui/highgui.hpp>: see cv_shim.hpp (test infrastructure only).
#include "../../cv_shim.hpp"
