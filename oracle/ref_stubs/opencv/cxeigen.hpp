// stands in for <opencv/cxeigen.hpp>: see minicv.hpp
#include "../minicv.hpp"
