// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Header-only stand-in for the part of OpenCV's C++ API that the reference's hot-path sources use, so that the
// reference's OWN translation units (src/ORBextractor.cc, src/LineMatcher.cpp, src/Config.cpp, ...) and function bodies
// sliced out of its larger files compile UNMODIFIED here (this image has no OpenCV C++).  Built only by
// oracle/Makefile.ref into oracle/_ref/libref.so; used only by tests/ to pin the oracle to the reference itself.
//
// Two kinds of content:
//  * containers (Mat, Point_, KeyPoint, Ptr, InputArray...): plain data structures with OpenCV's field names;
//  * un-vendored OpenCV ARITHMETIC (FAST, GaussianBlur, resize, Sobel, fastAtan2, LSD, BFMatcher, small float gemm):
//    forwarded to oracle/cvprim.hpp / oracle/line.cpp, whose semantics are pinned bit-for-bit to cv2 4.13 by
//    tests/test_oracle_golden.py.  Everything the reference VENDORS is the reference's own source.
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
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>
#include "../cvprim.hpp"

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_16SC1 CV_16S
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_PI 3.1415926535897932384626433832795
#define CV_OUT
#define CV_IN_OUT
#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_Assert(x) assert(x)
#define CV_DbgAssert(x) assert(x)

typedef unsigned char uchar;
typedef unsigned short ushort;

static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return (int)std::floor(v); }
static inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {
using std::vector; using std::min; using std::max; using std::abs; using std::swap; using std::sqrt; using std::exp; using std::pow; using std::log;
typedef std::string String;
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_LINEAR_EXACT = 5 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };
enum { COLOR_BGR2GRAY = 6 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
    Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_ operator*(float s) const { return Point_((T)(x * s), (T)(y * s)); }
};
// cv::Point2i(float, float) truncates like saturate_cast<int> of a float -> cvRound in OpenCV; the reference only passes
// values through int conversions of already integral floats (hX*i is not integral: Point2i(float,int) -> int(float)).
template <> template <> inline Point_<int>::Point_(const Point_<float>& o) : x(cvRound(o.x)), y(cvRound(o.y)) {}
typedef Point_<int> Point2i; typedef Point2i Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
    T area() const { return width * height; }
};
typedef Size_<int> Size;
template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b) { static_assert(N >= 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c, T d) { static_assert(N >= 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f; typedef Vec<int, 4> Vec4i; typedef Vec<float, 2> Vec2f; typedef Vec<double, 2> Vec2d;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } static Scalar all(double v) { return Scalar(v, v, v, v); } };

struct KeyPoint {
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(std::numeric_limits<float>::max()) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
struct KeyPointsFilter { static void retainBest(vector<KeyPoint>&, int) { fprintf(stderr, "cvstub: KeyPointsFilter::retainBest is not provided\n"); abort(); } };

template <typename T> struct Ptr : public std::shared_ptr<T> {
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const Ptr<U>& p) : std::shared_ptr<T>(std::static_pointer_cast<T>((const std::shared_ptr<U>&)p)) {}
    bool empty() const { return !this->get(); }
    void release() { this->reset(); }
    operator T*() const { return this->get(); }
};
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

// ---- Mat: reference-counted 2-D array with ROI views ----------------------------------------------------------------
struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
    size_t operator[](int i) const { return i == 0 ? v : elem; }
    size_t elem = 1;
};
class Mat;
struct MatExpr;
class Mat {
public:
    int flags = 0, dims = 2, rows = 0, cols = 0;
    uchar* data = nullptr;
    MatStep step;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)ext) { step = MatStep(step_ ? step_ : (size_t)c * esz(type)); step.elem = esz(type); }
    Mat(const Mat& m, const Rect& r) { *this = m; data = m.data + (size_t)r.y * m.step + (size_t)r.x * m.elemSize(); rows = r.height; cols = r.width; }
    template <typename T> explicit Mat(const vector<T>& v) { create((int)v.size(), 1, mtype((T*)0)); if (!v.empty()) memcpy(data, v.data(), v.size() * sizeof(T)); }
    static size_t esz(int type) { switch (type & 7) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_64F: return 8; default: return 4; } }
    static int mtype(uchar*) { return CV_8U; } static int mtype(float*) { return CV_32F; } static int mtype(double*) { return CV_64F; }
    static int mtype(int*) { return CV_32S; } static int mtype(short*) { return CV_16S; }
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && this->type() == type && buf_) return;
        flags = type; rows = r; cols = c; step = MatStep((size_t)c * esz(type)); step.elem = esz(type);
        buf_ = std::make_shared<vector<uchar>>((size_t)r * step + 64);
        data = buf_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = MatStep(0); }
    // Mat::zeros is a MatExpr in OpenCV: assigned to a Mat of the same size and type it is evaluated IN PLACE (create() keeps the
    // buffer), which src/ORBextractor.cc:1039 relies on (`descriptors = Mat::zeros(...)` on a row range of the output matrix)
    struct Zeros { int r, c, type; operator Mat() const { Mat m(r, c, type); m.fill0(); return m; } };
    static Zeros zeros(int r, int c, int type) { return Zeros{r, c, type}; }
    static Zeros zeros(Size s, int type) { return Zeros{s.height, s.width, type}; }
    Mat& operator=(const Zeros& z) { create(z.r, z.c, z.type); fill0(); return *this; }
    Mat(const Zeros& z) { create(z.r, z.c, z.type); fill0(); }
    void fill0() { for (int r = 0; r < rows; ++r) memset(ptr(r), 0, (size_t)cols * elemSize()); }
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(1)); return m; }
    void convertTo(Mat& o, int rtype) const {             // 8U -> 32F / same type (the SAD windows of Frame::ComputeStereoMatches)
        Mat d(rows, cols, rtype);
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) {
            double v = depth() == CV_8U ? (double)ptr(r)[c] : depth() == CV_32F ? (double)at<float>(r, c) : at<double>(r, c);
            if (rtype == CV_32F) d.at<float>(r, c) = (float)v; else if (rtype == CV_64F) d.at<double>(r, c) = v; else d.ptr(r)[c] = (uchar)v;
        }
        o = d;
    }
    static Mat eye(int r, int c, int type) { Mat m(r, c, type); m.fill0(); for (int i = 0; i < std::min(r, c); ++i) { if (type == CV_32F) m.at<float>(i, i) = 1.f; else if (type == CV_64F) m.at<double>(i, i) = 1.0; else m.ptr(i)[i] = 1; } return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return 1; }
    size_t elemSize() const { return esz(flags); }
    size_t elemSize1() const { return esz(flags); }
    size_t step1() const { return step / esz(flags); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows == 1; }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    Mat row(int r) const { return Mat(*this, Rect(0, r, cols, 1)); }
    Mat col(int c) const { return Mat(*this, Rect(c, 0, 1, rows)); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat rowRange(const Range& r) const { return rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return colRange(r.start, r.end); }
    Mat operator()(const Rect& r) const { return Mat(*this, r); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& o) const {
        if (empty()) { o.release(); return; }
        if (o.data == data && o.rows == rows && o.cols == cols) return;
        o.create(rows, cols, type());
        for (int r = 0; r < rows; ++r) memcpy(o.ptr(r), ptr(r), (size_t)cols * elemSize());
    }
    void setTo(const Scalar& s) {
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) {
            switch (depth()) { case CV_8U: ptr(r)[c] = (uchar)s.val[0]; break; case CV_16S: ptr<short>(r)[c] = (short)s.val[0]; break;
                case CV_32S: ptr<int>(r)[c] = (int)s.val[0]; break; case CV_32F: ptr<float>(r)[c] = (float)s.val[0]; break; default: ptr<double>(r)[c] = s.val[0]; }
        }
    }
    Mat& operator=(const Scalar& s) { setTo(s); return *this; }
    void reserve(size_t) {}
    void push_back(const Mat& m) {                       // append rows
        if (m.empty()) return;
        Mat n;
        const int r0 = empty() ? 0 : rows;
        n.create(r0 + m.rows, m.cols, m.type());
        for (int r = 0; r < r0; ++r) memcpy(n.ptr(r), ptr(r), (size_t)cols * elemSize());
        for (int r = 0; r < m.rows; ++r) memcpy(n.ptr(r0 + r), m.ptr(r), (size_t)m.cols * m.elemSize());
        *this = n;
    }
    // small float algebra (pose arithmetic of the matchers); products accumulate in source order like the oracle's
    // mat3_mul_vec, whose agreement with cv2.gemm on 3x3 * 3x1 float is pinned by tests/test_oracle_golden.py
    Mat t() const { Mat m(cols, rows, type()); for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) memcpy(m.ptr(c) + (size_t)r * elemSize(), ptr(r) + (size_t)c * elemSize(), elemSize()); return m; }
    double dot(const Mat& o) const { double s = 0; for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) s += depth() == CV_32F ? (double)at<float>(r, c) * o.at<float>(r, c) : at<double>(r, c) * o.at<double>(r, c); return s; }
private:
    std::shared_ptr<vector<uchar>> buf_;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, Mat::mtype((T*)0)) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
};
inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.type() == b.type());
    Mat m(a.rows, b.cols, a.type());
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < b.cols; ++c) {
        if (a.depth() == CV_32F) { float s = a.at<float>(r, 0) * b.at<float>(0, c); for (int k = 1; k < a.cols; ++k) s = s + a.at<float>(r, k) * b.at<float>(k, c); m.at<float>(r, c) = s; }
        else { double s = a.at<double>(r, 0) * b.at<double>(0, c); for (int k = 1; k < a.cols; ++k) s = s + a.at<double>(r, k) * b.at<double>(k, c); m.at<double>(r, c) = s; }
    }
    return m;
}
template <typename F> inline Mat mat_zip(const Mat& a, const Mat& b, F f) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat m(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) {
        if (a.depth() == CV_32F) m.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c)); else m.at<double>(r, c) = f(a.at<double>(r, c), b.at<double>(r, c));
    }
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return mat_zip(a, b, [](auto x, auto y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return mat_zip(a, b, [](auto x, auto y) { return x - y; }); }
inline Mat operator-(const Mat& a) { Mat m(a.rows, a.cols, a.type()); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) { if (a.depth() == CV_32F) m.at<float>(r, c) = -a.at<float>(r, c); else m.at<double>(r, c) = -a.at<double>(r, c); } return m; }
// MatExpr scaling (Mat * double, Mat / double): OpenCV evaluates it as a.convertTo(dst, -1, alpha) whose CV_32F kernel (cvtScale32f) multiplies
// by the alpha ROUNDED TO FLOAT, one float product per element; a / s is a * (1. / s) (MatExpr operator/ in matop.cpp)
inline Mat operator*(const Mat& a, double s) { Mat m(a.rows, a.cols, a.type()); const float sf = (float)s; for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) { if (a.depth() == CV_32F) m.at<float>(r, c) = a.at<float>(r, c) * sf; else m.at<double>(r, c) = a.at<double>(r, c) * s; } return m; }
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator/(const Mat& a, double s) { return a * (1. / s); }
// cv::norm(Mat) L2 of a small float vector: double accumulation, sqrt (OpenCV's normL2_32f accumulates in double)
inline double norm(const Mat& a) { double s = 0; for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) { const double v = a.depth() == CV_32F ? (double)a.at<float>(r, c) : a.at<double>(r, c); s += v * v; } return std::sqrt(s); }
// cv::norm(a, b, NORM_L1) on 8-bit windows (Frame.cc:820): integer sum of absolute differences
inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(normType == NORM_L1 && a.rows == b.rows && a.cols == b.cols);
    long s = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) {
        if (a.depth() == CV_8U) s += std::abs((int)a.ptr(r)[c] - (int)b.ptr(r)[c]);
        else if (a.depth() == CV_32F) return [&] { double t = 0; for (int rr = 0; rr < a.rows; ++rr) for (int cc = 0; cc < a.cols; ++cc) t += std::fabs((double)a.at<float>(rr, cc) - (double)b.at<float>(rr, cc)); return t; }();
    }
    return (double)s;
}

// ---- InputArray / OutputArray with the real proxies' semantics (getMat / create) -------------------------------------
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
    Size size() const { return m_ ? m_->size() : Size(); }
    int type() const { return m_ ? m_->type() : 0; }
protected:
    const Mat* m_ = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m), w_(&m) {}
    void create(int r, int c, int type) const { if (w_) w_->create(r, c, type); }
    void create(Size s, int type) const { create(s.height, s.width, type); }
    void release() const { if (w_) w_->release(); }
    bool needed() const { return w_ != nullptr; }
    Mat& getMatRef() const { return *w_; }
private:
    Mat* w_ = nullptr;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline const _OutputArray& noArray() { static _OutputArray a; return a; }

// ---- un-vendored OpenCV arithmetic -> oracle/cvprim.hpp (pinned to cv2 4.13) -------------------------------------------
inline float fastAtan2(float y, float x) { return orc::fast_atan2_deg(y, x); }
inline orc::Image8 to_image8(const Mat& m) { orc::Image8 im(m.cols, m.rows); for (int r = 0; r < m.rows; ++r) memcpy(im.row(r), m.ptr(r), m.cols); return im; }
inline void from_image8(const orc::Image8& im, Mat& m) { m.create(im.h, im.w, CV_8UC1); for (int r = 0; r < im.h; ++r) memcpy(m.ptr(r), im.row(r), im.w); }
inline void FAST(InputArray image, vector<KeyPoint>& kps, int threshold, bool nms = true) {
    const Mat m = image.getMat();
    vector<orc::FastKp> out;
    orc::fast_detect(m.ptr(), m.cols, m.rows, (int)(size_t)m.step, threshold, nms, out);
    kps.clear();
    for (const orc::FastKp& k : out) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1.f, (float)k.score));
}
inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    assert(ksize.width == ksize.height && (sigmaY == 0 || sigmaY == sigmaX) && (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    const orc::Image8 in = to_image8(src.getMat());
    orc::Image8 out(in.w, in.h);
    orc::gaussian_blur_q8(in, orc::gauss_kernel_q8(ksize.width, sigmaX), out);
    dst.create(in.h, in.w, CV_8UC1);
    Mat d = dst.getMat();
    for (int r = 0; r < in.h; ++r) memcpy(d.ptr(r), out.row(r), in.w);
}
inline void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    assert(interpolation == INTER_LINEAR && dsize.width > 0);
    const orc::Image8 in = to_image8(src.getMat());
    orc::Image8 out(dsize.width, dsize.height);
    orc::resize_linear(in, out);
    Mat d = dst.getMat();                                  // the reference resizes INTO a pre-allocated ROI of the bordered level
    if (d.rows != dsize.height || d.cols != dsize.width) { dst.create(dsize.height, dsize.width, CV_8UC1); d = dst.getMat(); }
    for (int r = 0; r < out.h; ++r) memcpy(d.ptr(r), out.row(r), out.w);
}
// copyMakeBorder(src, dst, ...) as the reference uses it: dst is the bordered parent buffer of the level; BORDER_REFLECT_101
// (+BORDER_ISOLATED for level > 0).  Border pixels are never read on the hot path but are filled faithfully.
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar()) {
    const Mat src = src_.getMat();
    const int h = src.rows + top + bottom, w = src.cols + left + right;
    Mat tmp(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) {
        const int sy = orc::reflect101(y - top, src.rows);
        for (int x = 0; x < w; ++x) tmp.ptr(y)[x] = src.ptr(sy)[orc::reflect101(x - left, src.cols)];
    }
    (void)borderType;
    Mat d = dst_.getMat();
    if (d.rows != h || d.cols != w) { dst_.create(h, w, CV_8UC1); d = dst_.getMat(); }
    for (int y = 0; y < h; ++y) memcpy(d.ptr(y), tmp.ptr(y), w);
}
inline void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize = 3, double = 1, double = 0, int = BORDER_DEFAULT) {
    assert(ksize == 3 && ddepth == CV_16S && dx + dy == 1);
    const orc::Image8 in = to_image8(src_.getMat());
    vector<int16_t> gx, gy;
    orc::sobel3(in, gx, gy);
    dst_.create(in.h, in.w, CV_16SC1);
    Mat d = dst_.getMat();
    const vector<int16_t>& g = dx ? gx : gy;
    for (int r = 0; r < in.h; ++r) memcpy(d.ptr(r), g.data() + (size_t)r * in.w, (size_t)in.w * 2);
}
// cv::LineIterator: only .count is consumed (LSDDetector_custom.cpp:295-296), 8-connected
class LineIterator {
public:
    int count;
    LineIterator(const Mat&, Point p1, Point p2, int connectivity = 8, bool = false) { assert(connectivity == 8); count = std::max(std::abs(p2.x - p1.x), std::abs(p2.y - p1.y)) + 1; }
};

// cv::BFMatcher(NORM_HAMMING).knnMatch: brute force, ascending distance, ties -> lowest train index (pinned by the
// cv2 golden vectors of oracle's knn2)
class DescriptorMatcher { public: virtual ~DescriptorMatcher() {} };
class BFMatcher : public DescriptorMatcher {
public:
    BFMatcher(int normType = NORM_L2, bool crossCheck = false) : norm_(normType) { (void)crossCheck; }
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false) { return Ptr<BFMatcher>(new BFMatcher(normType, crossCheck)); }
    void knnMatch(InputArray q_, InputArray t_, vector<vector<DMatch>>& out, int k, InputArray = noArray(), bool = false) const {
        assert(norm_ == NORM_HAMMING);
        const Mat q = q_.getMat(), t = t_.getMat();
        out.assign(q.rows, vector<DMatch>());
        for (int i = 0; i < q.rows; ++i) {
            vector<std::pair<int, int>> d(t.rows);
            for (int j = 0; j < t.rows; ++j) d[j] = std::make_pair(orc::hamming256(q.ptr(i), t.ptr(j)), j);
            const int kk = std::min(k, t.rows);
            std::partial_sort(d.begin(), d.begin() + kk, d.end());
            for (int j = 0; j < kk; ++j) out[i].push_back(DMatch(i, d[j].second, (float)d[j].first));
        }
    }
private:
    int norm_;
};

// cv::FileStorage: only needed so that src/Config.cpp (loadFromFile) compiles; never opened here
class FileNode {
public:
    enum { NONE = 0, INT = 1, REAL = 2, FLOAT = REAL, STR = 3, STRING = STR, SEQ = 4, MAP = 5 };
    int type() const { return NONE; }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
    std::string string() const { return std::string(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline void operator>>(const FileNode&, T&) {}
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }  // never opened: writes go nowhere

class Algorithm { public: virtual ~Algorithm() {} };
}  // namespace cv

// cv::LineSegmentDetector is un-vendored OpenCV: the oracle's LSD (oracle/line.cpp, pinned bit-for-bit to
// cv2.createLineSegmentDetector on 60+ images by tests/test_oracle_golden.py / test_oracle_vs_cv2.py)
#include "../oracle.h"
namespace cv {
inline void pyrDown(InputArray, OutputArray, Size = Size()) { fprintf(stderr, "cvstub: pyrDown is not provided (numOctaves == 1 on the hot path)\n"); abort(); }
inline void cvtColor(InputArray, OutputArray, int) { fprintf(stderr, "cvstub: cvtColor is not provided (8UC1 input only)\n"); abort(); }
class LineSegmentDetector : public Algorithm {
public:
    olf_line_params P;
    void detect(InputArray image, vector<Vec4f>& lines) {
        const Mat m = image.getMat();
        orc_line* h = orc_line_create(&P);
        vector<float> segs((size_t)4 * 1 << 18);
        int n = 0;
        const int rc = orc_lsd_detect(h, m.ptr(), m.cols, m.rows, (int)(size_t)m.step, segs.data(), 1 << 18, &n);
        orc_line_destroy(h);
        if (rc) throw std::runtime_error("cvstub: oracle LSD failed");
        lines.clear();
        for (int i = 0; i < n; ++i) lines.push_back(Vec4f(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]));
    }
};
inline Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8, double sigma_scale = 0.6, double quant = 2.0,
                                                          double ang_th = 22.5, double log_eps = 0, double density_th = 0.7, int n_bins = 1024) {
    Ptr<LineSegmentDetector> p(new LineSegmentDetector());
    memset(&p->P, 0, sizeof(p->P));
    p->P.lsd_nfeatures = 0; p->P.min_line_length = 0; p->P.lsd_refine = refine; p->P.lsd_scale = scale; p->P.lsd_sigma_scale = sigma_scale;
    p->P.lsd_quant = quant; p->P.lsd_ang_th = ang_th; p->P.lsd_log_eps = log_eps; p->P.lsd_density_th = density_th; p->P.lsd_n_bins = n_bins;
    return p;
}
}  // namespace cv
