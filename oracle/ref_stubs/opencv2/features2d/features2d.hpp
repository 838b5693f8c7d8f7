// stands in for <opencv2/features2d/features2d.hpp>: see minicv.hpp
#include "../../minicv.hpp"
