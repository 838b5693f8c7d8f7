// stands in for <opencv2/highgui/highgui.hpp>: see minicv.hpp
#include "../../minicv.hpp"
