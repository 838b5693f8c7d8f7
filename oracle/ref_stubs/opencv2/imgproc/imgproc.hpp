// stands in for <opencv2/imgproc/imgproc.hpp>: see minicv.hpp
#include "../../minicv.hpp"
