// Optimizer.h — drop-in for the reference's include/Optimizer.h: the pose-only g2o Levenberg-Marquardt of
// Optimizer::PoseOptimization (src/Optimizer.cc:15-86) runs on the device (svo_pose_optimize); no g2o or
// Eigen is needed on the host for it.
#pragma once
#include "frame.h"

class Optimizer {
public:
    // Optimises pFrame->Tcw over the frame's matched map points (EdgeSE3ProjectXYZOnlyPose, Huber
    // sqrt(5.991), optimize(10)), calls pFrame->SetPose and returns nInitialCorrespondences.
    static int PoseOptimization(frame *pFrame);
};
