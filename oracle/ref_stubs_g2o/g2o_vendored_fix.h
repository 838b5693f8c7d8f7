// Force-included by `make ref_g2o`.  The vendored Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp does not compile as
// shipped: line 126 reads `c(0,0) = x*y/z_2 *fx;` where upstream g2o has `_jacobianOplusXj(0,0) = ...` (inside
// EdgeSE3ProjectXYZ::linearizeOplus, the point-and-pose edge that the pose-only optimisation of src/Optimizer.cc never
// creates).  Declaring a matrix of that name in namespace g2o lets the file compile UNMODIFIED; the function it sits in
// is never called on the pinned path.
#pragma once
#ifdef __cplusplus
#include <Eigen/Core>
namespace g2o { static Eigen::Matrix<double, 2, 6> c; }
#endif
