// stands in for <opencv/cv.hpp>: see minicv.hpp
#include "../minicv.hpp"
