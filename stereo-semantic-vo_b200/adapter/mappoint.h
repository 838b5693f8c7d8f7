// mappoint.h — host-side map point as the reference's matchers see it (include/mappoint.h):
// a world position, ONE descriptor row frozen at birth (src/mappoint.cc:12), a bad flag and
// the observation table.  Pure bookkeeping; nothing here runs on the device.
#pragma once
#include <map>
#include "cv_shim.h"

class frame;

class mappoint {
public:
    mappoint(cv::Mat &pos, frame *pFrame, int id);
    void AddObservation(frame *fm, size_t idx);

    cv::Mat worldpos;
    cv::Mat m_descriptor;   // 1 x 32, CV_8U
    bool bad;
    int observation_num;
    int create_id;
    int octave;             // pyramid level of the keypoint the point was created from (projection windows; not in the reference)
    std::map<frame *, int> observations;
};
