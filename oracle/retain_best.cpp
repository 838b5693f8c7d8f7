// retain_best.cpp — replay of cv::KeyPointsFilter::retainBest (OpenCV features2d,
// called twice per level by cv::ORB, itself called from src/frame.cc:77-78).
// TEST INFRASTRUCTURE ONLY (see svo_oracle.c).  The kept SET is defined by the
// response threshold; the kept ORDER is whatever libstdc++'s introselect and
// partition leave behind, so this file calls exactly those two std algorithms
// (SURVEY.md Appendix A.6).
#include <algorithm>
#include <cstdint>
#include <vector>
#include "svo_oracle.h"

namespace {
struct Item { float response; int32_t idx; };
struct Greater { bool operator()(const Item& a, const Item& b) const { return a.response > b.response; } };
}

extern "C" int svo_o_retain_best(float* resp, int32_t* idx, int n, int n_points)
{
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    std::vector<Item> v(n);
    for (int i = 0; i < n; ++i) v[i] = Item{resp[i], idx[i]};
    std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), Greater());
    const float amb = v[n_points - 1].response;
    auto new_end = std::partition(v.begin() + n_points, v.end(), [amb](const Item& k) { return k.response >= amb; });
    const int kept = int(new_end - v.begin());
    for (int i = 0; i < n; ++i) { resp[i] = v[i].response; idx[i] = v[i].idx; }
    return kept;
}

// Test shims: call libstdc++'s own internals so the GPU replay can be checked
// on forced depth limits (heap-select fallback) and on bare key arrays.
extern "C" void svo_o_introselect(float* resp, int32_t* idx, int n, int nth, int depth_limit)
{
    std::vector<Item> v(n);
    for (int i = 0; i < n; ++i) v[i] = Item{resp[i], idx[i]};
    if (n > 0 && nth < n) {
        if (depth_limit < 0) depth_limit = std::__lg(n) * 2;
        std::__introselect(v.begin(), v.begin() + nth, v.end(), depth_limit,
                           __gnu_cxx::__ops::__iter_comp_iter(Greater()));
    }
    for (int i = 0; i < n; ++i) { resp[i] = v[i].response; idx[i] = v[i].idx; }
}
