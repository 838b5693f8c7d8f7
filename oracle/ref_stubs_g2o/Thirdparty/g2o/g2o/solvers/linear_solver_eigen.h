// shadows Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h (sparse Eigen solver): src/Optimizer.cc includes it and uses
// nothing of it (PoseOptimization builds a LinearSolverDense, src/Optimizer.cc:18)
