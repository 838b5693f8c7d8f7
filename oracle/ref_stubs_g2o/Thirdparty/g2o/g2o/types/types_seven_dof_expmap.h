// shadows Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h (Sim3 types): include/Optimizer.h includes it and
// PoseOptimization uses nothing of it
