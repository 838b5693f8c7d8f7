// stands in for <opencv2/calib3d.hpp>: see minicv.hpp
#include "../minicv.hpp"
