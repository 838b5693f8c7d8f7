// stands in for <opencv2/opencv.hpp>: see minicv.hpp
#include "../minicv.hpp"
