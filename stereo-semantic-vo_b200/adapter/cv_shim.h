// cv_shim.h — the few OpenCV value types the reference's frame/pnpmatch interface exposes
// (cv::Mat, cv::KeyPoint, cv::Point2f, cv::DMatch), for building the adapter where OpenCV's
// C++ headers are not installed (this image).  With OpenCV present, define
// SVO_ADAPTER_USE_OPENCV and the real headers are used instead; the adapter only touches the
// members below.
#pragma once
#ifdef SVO_ADAPTER_USE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace cv {

enum { CV_8U_ = 0, CV_32F_ = 5 };
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16

struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// Row-major, reference-counted 2-D array; type = depth | (channels-1) << 3 like OpenCV.
class Mat {
public:
    int rows, cols;
    size_t step;
    uint8_t *data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, double fill) { create(r, c, type); setTo(fill); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uint8_t>(new uint8_t[step * (size_t)r + 16], std::default_delete<uint8_t[]>());
        data = buf_.get();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * (depth() == 6 ? 8 : depth() == 5 ? 4 : 1); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Mat clone() const
    {
        Mat m;
        if (!empty()) { m.create(rows, cols, type_); for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize()); }
        return m;
    }
    Mat row(int r) const
    {
        Mat m(*this);
        m.rows = 1; m.data = data + (size_t)r * step;
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    uint8_t *ptr(int r = 0) { return data + (size_t)r * step; }
    const uint8_t *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(data + (size_t)r * step)[c]; }
    template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(data + (size_t)r * step)[c]; }
    void setTo(double v)
    {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * channels(); ++c) {
                if (depth() == 6) ptr<double>(r)[c] = v;
                else if (depth() == 5) ptr<float>(r)[c] = (float)v;
                else ptr(r)[c] = (uint8_t)v;
            }
    }
private:
    int type_;
    std::shared_ptr<uint8_t> buf_;
};

}  // namespace cv
#endif
