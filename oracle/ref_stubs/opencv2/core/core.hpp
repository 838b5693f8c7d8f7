// stands in for <opencv2/core/core.hpp>: see minicv.hpp
#include "../../minicv.hpp"
