// stands in for <pangolin/pangolin.h> (the viewer: out of scope).  src/Tracking.cc names one type and one call.
#pragma once
namespace pangolin {
struct OpenGlMatrix {
    double m[16];
    void SetIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
};
inline void FinishFrame() {}
}
