// stands in for <opencv2/core/affine.hpp>: see minicv.hpp
#include "../../minicv.hpp"
