// minicv.hpp — TEST INFRASTRUCTURE.  The handful of OpenCV types and calls that the reference's
// src/pnpmatch.cc, src/frame.cc and src/mappoint.cc touch, so that those three files compile UNMODIFIED,
// from where they lie under /root/reference, into oracle/_ref/libsvo_ref.so (recipe: oracle/Makefile, target
// `ref`).  No OpenCV C++ headers exist in this image; the arithmetic that OpenCV itself would do (cv::ORB,
// findFundamentalMat, solvePnPRansac, Rodrigues) is NOT restated here: those calls are forwarded through C
// callbacks that oracle/ref.py points at the real cv2.  What is implemented here is container plumbing
// (Mat views, clone, at<>), the 3x3 / 4x4 float products of frame::SetPose / UnprojectStereo (OpenCV's small
// gemm path: float products summed left to right, checked against cv2.gemm in tests/test_ref_pin.py), a
// brute-force Hamming matcher with BFMatcher's first-minimum rule (checked against cv2.BFMatcher there too)
// and no-op drawing / GUI calls (cv::circle, cv::line, imshow, waitKey — SURVEY.md Appendix D).
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_FM_8POINT 2

namespace cv {
typedef unsigned char uchar;
typedef std::string String;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
struct Vec3b {                      // Thirdparty/MB/MSA.cpp paints a debug image through at<Vec3b>
    unsigned char val[3];
    Vec3b() { val[0] = val[1] = val[2] = 0; }
    Vec3b(int a, int b, int c) { val[0] = (unsigned char)a; val[1] = (unsigned char)b; val[2] = (unsigned char)c; }
    unsigned char &operator[](int i) { return val[i]; }
    const unsigned char &operator[](int i) const { return val[i]; }
};
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
};
typedef Rect_<int> Rect;
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
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

class Mat;
template <typename T> class MatCommaInitializer_;

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar *data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[step * (size_t)(r > 0 ? r : 0) + 64](), std::default_delete<uchar[]>());
        data = buf_.get();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { static const int sz[7] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)sz[depth()]; }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Mat clone() const
    {
        Mat m;
        if (data) { m.create(rows, cols, type_); for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize()); }
        return m;
    }
    // views share the buffer
    Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    void copyTo(Mat &dst) const
    {
        if (dst.data && dst.rows == rows && dst.cols == cols && dst.type_ == type_) {
            for (int r = 0; r < rows; ++r) std::memcpy(dst.ptr(r), ptr(r), (size_t)cols * elemSize());
        } else dst = clone();
    }
    void copyTo(Mat &&dst) const { Mat d(dst); copyTo(d); }   // into a temporary view (outImg.rowRange(..))
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const
    {
        Mat out(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * channels(); ++c) out.put(r, c, get(r, c) * alpha + beta);
        dst = out;
    }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return reinterpret_cast<T *>(data + (size_t)r * step)[c]; }
    template <typename T> const T &at(int r, int c) const { return reinterpret_cast<const T *>(data + (size_t)r * step)[c]; }
    // single index: element i of a row or column vector (src/Optimizer.cc:67-69 reads Xw.at<float>(k) of a 3x1)
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    void setTo(const Scalar &s)
    {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * channels(); ++c) put(r, c, s.val[c % channels() < 4 ? c % channels() : 0]);
    }
    double get(int r, int c) const
    {
        switch (depth()) {
        case CV_8U: return ptr<uchar>(r)[c];
        case CV_32S: return ptr<int>(r)[c];
        case CV_32F: return ptr<float>(r)[c];
        case CV_64F: return ptr<double>(r)[c];
        case CV_16S: return ptr<short>(r)[c];
        case CV_16U: return ptr<unsigned short>(r)[c];
        default: return ptr<signed char>(r)[c];
        }
    }
    void put(int r, int c, double v)
    {
        switch (depth()) {
        case CV_8U: ptr<uchar>(r)[c] = (uchar)(v < 0 ? 0 : (v > 255 ? 255 : std::lrint(v))); break;
        case CV_32S: ptr<int>(r)[c] = (int)std::lrint(v); break;
        case CV_32F: ptr<float>(r)[c] = (float)v; break;
        case CV_64F: ptr<double>(r)[c] = v; break;
        case CV_16S: ptr<short>(r)[c] = (short)std::lrint(v); break;
        case CV_16U: ptr<unsigned short>(r)[c] = (unsigned short)std::lrint(v); break;
        default: ptr<signed char>(r)[c] = (signed char)std::lrint(v); break;
        }
    }
    Mat t() const
    {
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) std::memcpy(m.data + (size_t)c * m.step + (size_t)r * elemSize(), data + (size_t)r * step + (size_t)c * elemSize(), elemSize());
        return m;
    }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type, Scalar(0));
        for (int i = 0; i < r && i < c; ++i) m.put(i, i, 1.0);
        return m;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }
private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

// OpenCV's gemm for the small float matrices of this path (len <= 4): every output is the float sum of float
// products, accumulated left to right.  CV_64F operands accumulate in double.
inline Mat operator*(const Mat &a, const Mat &b)
{
    Mat d(a.rows, b.cols, a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < b.cols; ++c) {
            if (a.depth() == CV_32F) {
                float s = a.at<float>(r, 0) * b.at<float>(0, c);
                for (int k = 1; k < a.cols; ++k) { const float p = a.at<float>(r, k) * b.at<float>(k, c); s = s + p; }
                d.at<float>(r, c) = s;
            } else {
                double s = a.get(r, 0) * b.get(0, c);
                for (int k = 1; k < a.cols; ++k) { const double p = a.get(r, k) * b.get(k, c); s = s + p; }
                d.put(r, c, s);
            }
        }
    return d;
}
inline Mat operator+(const Mat &a, const Mat &b)
{
    Mat d(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            if (a.depth() == CV_32F) d.at<float>(r, c) = a.at<float>(r, c) + b.at<float>(r, c);
            else d.put(r, c, a.get(r, c) + b.get(r, c));
        }
    return d;
}
inline Mat operator-(const Mat &a)
{
    Mat d(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            if (a.depth() == CV_32F) d.at<float>(r, c) = -a.at<float>(r, c);
            else d.put(r, c, -a.get(r, c));
        }
    return d;
}
inline std::ostream &operator<<(std::ostream &o, const Mat &m)
{
    o << "[";
    for (int r = 0; r < m.rows; ++r) {
        for (int c = 0; c < m.cols * m.channels(); ++c) o << (c ? ", " : "") << m.get(r, c);
        o << (r + 1 < m.rows ? ";\n " : "");
    }
    return o << "]";
}

template <typename T> struct DepthOf;
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
};
// (cv::Mat_<float>(3,1) << x, y, z): values fill the matrix in row-major order, converted to T
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(const Mat_<T> &m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_<T> &operator,(U v)
    {
        const int r = i_ / m_.cols, c = i_ % m_.cols;
        m_.template at<T>(r, c) = (T)v;
        ++i_;
        return *this;
    }
    operator Mat() const { return m_; }
    Mat_<T> m_;
    int i_;
};
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, U v)
{
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

template <typename T> using Ptr = std::shared_ptr<T>;

// ---- hooks: the real OpenCV (cv2) behind C callbacks, installed by oracle/ref.py -------------------------
extern "C" {
// what: 0 detectAndCompute, 1 detect, 2 compute (kps in/out; n_in valid entries).  kps: n x 7 floats
// (x, y, size, angle, response, octave, class_id).  Returns the keypoint count.
typedef int (*minicv_orb_fn)(const uchar *img, int rows, int cols, int step, int channels, int what, float *kps, uchar *desc,
                             int cap, int n_in);
typedef int (*minicv_fund_fn)(const float *p1, const float *p2, int n, double *F9);            // returns 1 when F was found
typedef int (*minicv_pnp_fn)(const float *p3, const float *p2, int n, const float *K9, double *rvec, double *tvec, int *inl, int cap);
typedef void (*minicv_rodrigues_fn)(const double *rvec, double *R9);
}
struct MiniCvHooks { minicv_orb_fn orb; minicv_fund_fn fund; minicv_pnp_fn pnp; minicv_rodrigues_fn rodrigues; };
inline MiniCvHooks &minicv_hooks() { static MiniCvHooks h = {nullptr, nullptr, nullptr, nullptr}; return h; }
#define MINICV_ORB_CAP 8192

class Feature2D {
public:
    virtual ~Feature2D() {}
    void detectAndCompute(const Mat &img, const Mat &mask, std::vector<KeyPoint> &kps, Mat &desc) { (void)mask; run(img, 0, kps, desc); }
    void detect(const Mat &img, std::vector<KeyPoint> &kps) { Mat d; run(img, 1, kps, d); }
    void compute(const Mat &img, std::vector<KeyPoint> &kps, Mat &desc) { run(img, 2, kps, desc); }
private:
    void run(const Mat &img, int what, std::vector<KeyPoint> &kps, Mat &desc)
    {
        if (!minicv_hooks().orb) { std::fprintf(stderr, "minicv: no ORB hook installed\n"); std::abort(); }
        std::vector<float> k((size_t)MINICV_ORB_CAP * 7);
        std::vector<uchar> d((size_t)MINICV_ORB_CAP * 32);
        int n_in = 0;
        if (what == 2) {
            n_in = (int)kps.size();
            for (int i = 0; i < n_in; ++i) {
                float *p = &k[(size_t)i * 7];
                p[0] = kps[i].pt.x; p[1] = kps[i].pt.y; p[2] = kps[i].size; p[3] = kps[i].angle; p[4] = kps[i].response;
                p[5] = (float)kps[i].octave; p[6] = (float)kps[i].class_id;
            }
        }
        const int n = minicv_hooks().orb(img.data, img.rows, img.cols, (int)img.step, img.channels(), what, k.data(), d.data(), MINICV_ORB_CAP, n_in);
        kps.resize((size_t)n);
        for (int i = 0; i < n; ++i) {
            const float *p = &k[(size_t)i * 7];
            kps[i].pt = Point2f(p[0], p[1]); kps[i].size = p[2]; kps[i].angle = p[3]; kps[i].response = p[4];
            kps[i].octave = (int)p[5]; kps[i].class_id = (int)p[6];
        }
        if (what != 1) {
            desc = Mat(n, 32, CV_8U);
            if (n) std::memcpy(desc.data, d.data(), (size_t)n * 32);
        }
    }
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
class ORB : public Feature2D {
public:
    static Ptr<ORB> create() { return std::make_shared<ORB>(); }   // defaults: 500 features, 1.2, 8 levels
};

// cv::DescriptorMatcher::create("BruteForce-Hamming")->match: per query row the FIRST minimum over the train rows
class DescriptorMatcher {
public:
    static Ptr<DescriptorMatcher> create(const std::string &) { return std::make_shared<DescriptorMatcher>(); }
    void match(const Mat &q, const Mat &t, std::vector<DMatch> &out)
    {
        out.clear();
        if (t.rows == 0) return;
        for (int i = 0; i < q.rows; ++i) {
            int best = INT_MAX, bi = -1;
            for (int j = 0; j < t.rows; ++j) {
                int d = 0;
                const uchar *a = q.ptr(i), *b = t.ptr(j);
                for (int k = 0; k < q.cols; ++k) d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
                if (d < best) { best = d; bi = j; }
            }
            out.push_back(DMatch(i, bi, (float)best));
        }
    }
};

inline Mat findFundamentalMat(const std::vector<Point2f> &p1, const std::vector<Point2f> &p2, int method)
{
    (void)method;
    double F[9];
    if (!minicv_hooks().fund) { std::fprintf(stderr, "minicv: no findFundamentalMat hook\n"); std::abort(); }
    const int n = (int)p1.size();
    if (!minicv_hooks().fund(n ? &p1[0].x : nullptr, n ? &p2[0].x : nullptr, n, F)) return Mat();
    Mat m(3, 3, CV_64F);
    std::memcpy(m.data, F, sizeof F);
    return m;
}
inline bool solvePnPRansac(const std::vector<Point3f> &p3, const std::vector<Point2f> &p2, const Mat &K, const Mat &dist, Mat &rvec,
                           Mat &tvec, bool guess, int iters, float err, double conf, Mat &inliers)
{
    (void)dist; (void)guess; (void)iters; (void)err; (void)conf;   // ref.py passes the reference's literals: false, 100, 8.0, 0.99
    if (!minicv_hooks().pnp) { std::fprintf(stderr, "minicv: no solvePnPRansac hook\n"); std::abort(); }
    const int n = (int)p3.size();
    std::vector<int> inl((size_t)n + 1);
    double r[3] = {0, 0, 0}, t[3] = {0, 0, 0};
    float Kf[9];
    for (int i = 0; i < 9; ++i) Kf[i] = K.at<float>(i / 3, i % 3);
    const int ni = minicv_hooks().pnp(n ? &p3[0].x : nullptr, n ? &p2[0].x : nullptr, n, Kf, r, t, inl.data(), n);
    rvec = Mat(3, 1, CV_64F); tvec = Mat(3, 1, CV_64F);
    for (int i = 0; i < 3; ++i) { rvec.at<double>(i, 0) = r[i]; tvec.at<double>(i, 0) = t[i]; }
    inliers = Mat(ni > 0 ? ni : 0, 1, CV_32S);
    for (int i = 0; i < ni; ++i) inliers.at<int>(i, 0) = inl[(size_t)i];
    return ni > 0;
}
inline void Rodrigues(const Mat &src, Mat &dst)
{
    if (!minicv_hooks().rodrigues) { std::fprintf(stderr, "minicv: no Rodrigues hook\n"); std::abort(); }
    double r[3] = {src.get(0, 0), src.get(1, 0), src.get(2, 0)}, R[9];
    minicv_hooks().rodrigues(r, R);
    dst = Mat(3, 3, CV_64F);
    std::memcpy(dst.data, R, sizeof R);
}

// drawing / GUI: no-ops (SURVEY.md Appendix D lists them as defects the baseline strips)
inline void circle(const Mat &, Point2f, int, const Scalar &, int = 1) {}
inline void line(const Mat &, Point2f, Point2f, const Scalar &, int = 1) {}
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int = 0) { return -1; }
inline Mat imread(const std::string &, int = 1) { return Mat(); }
inline bool imwrite(const std::string &, const Mat &) { return true; }

// frame::ElasMatch (never called, src/Tracking.cc:226 uses MB) still has to compile
class StereoSGBM {
public:
    enum { MODE_SGBM = 0 };
    static Ptr<StereoSGBM> create(int, int, int) { return std::make_shared<StereoSGBM>(); }
    void setPreFilterCap(int) {} void setBlockSize(int) {} void setP1(int) {} void setP2(int) {} void setMinDisparity(int) {}
    void setNumDisparities(int) {} void setUniquenessRatio(int) {} void setSpeckleWindowSize(int) {} void setSpeckleRange(int) {}
    void setDisp12MaxDiff(int) {} void setMode(int) {}
    void compute(const Mat &, const Mat &, Mat &) {}
};
// cv::FileStorage(path, READ)["Camera.fx"] as src/Tracking.cc:24-38 uses it: "key: number" lines of an OpenCV YAML file
class FileNode {
    double v_;
public:
    explicit FileNode(double v = 0) : v_(v) {}
    operator float() const { return (float)v_; }
    operator double() const { return v_; }
    operator int() const { return (int)v_; }
};
class FileStorage {
    std::map<std::string, double> kv_;
    bool ok_;
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string &path, int) : ok_(false)
    {
        std::ifstream f(path.c_str());
        std::string line;
        while (std::getline(f, line)) {
            const size_t c = line.find(':');
            if (c == std::string::npos || line[0] == '%' || line[0] == '#') continue;
            char *end = nullptr;
            const std::string val = line.substr(c + 1);
            const double d = std::strtod(val.c_str(), &end);
            if (end != val.c_str()) { kv_[line.substr(0, c)] = d; ok_ = true; }
        }
    }
    bool isOpened() const { return ok_; }
    FileNode operator[](const std::string &k) const { auto it = kv_.find(k); return FileNode(it == kv_.end() ? 0.0 : it->second); }
    FileNode operator[](const char *k) const { return (*this)[std::string(k)]; }
    void release() {}
};

}  // namespace cv

// legacy C struct used by include/YOLOv3SE.h (not on the path)
struct IplImage {
    int nChannels, width, height, widthStep;
    char *imageData;
    IplImage(const cv::Mat &m) : nChannels(m.channels()), width(m.cols), height(m.rows), widthStep((int)m.step), imageData((char *)m.data) {}
};
