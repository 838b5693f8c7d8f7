// stands in for <opencv2/core/mat.hpp>: see minicv.hpp
#include "../../minicv.hpp"
