// stands in for include/convert.h (Eigen/g2o conversions; src/pnpmatch.cc includes it and uses nothing of it)
