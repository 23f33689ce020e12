// Minimal stand-in for the handful of OpenCV types that appear in the reference's front-end class surfaces
// (cv::Mat, cv::KeyPoint, cv::Point2f, cv::InputArray/OutputArray, cv::line_descriptor::KeyLine).  Used ONLY when the
// shim is built without OpenCV (this image has no OpenCV C++); define OLF_HAVE_OPENCV to compile the very same shim
// sources against the real headers inside the reference tree.  InputArray / OutputArray are proxy classes with the real
// ones' semantics (getMat(), create(), release(), empty()) so that the shim sources use exactly the calls that exist on
// cv::_InputArray / cv::_OutputArray (the reference does `Mat image = _image.getMat()`, `_descriptors.create(n,32,CV_8U)`,
// src/ORBextractor.cc:1051,1070).
#pragma once
#ifdef OLF_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <line_descriptor_custom.hpp>
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
namespace cv {
typedef unsigned char uchar;
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct KeyPoint {                      // cv::KeyPoint, same field names and defaults
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
class Mat {
public:
    int rows = 0, cols = 0; size_t step = 0; uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)ext), type_(type) {}
    void create(int r, int c, int type) {
        if (data && buf_ && rows == r && cols == c && type_ == type) return;
        type_ = type; rows = r; cols = c; step = (size_t)c * elem();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step);
        data = buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }                  // vector access, as cv::Mat::at(int)
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat rowRange(int a, int b) const { Mat m = row(a); m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + (size_t)a * elem(); return m; }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat t() const { Mat m(cols, rows, type_); for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c); return m; }
    double dot(const Mat& o) const { double s = 0; for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * o.at<float>(r, c); return s; }
    Mat row(int r) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.type_ = type_; m.buf_ = buf_; return m; }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; ++r) memcpy(m.ptr(r), ptr(r), (size_t)cols * elem()); return m; }
    void copyTo(Mat& o) const { o = clone(); }
private:
    size_t elem() const { return type_ == CV_32F ? 4 : 1; }
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<uchar>> buf_;
};
// The CV_32F matrix algebra the ORBmatcher overloads use on 3x3 / 3x1 / 4x4 matrices, with OpenCV's arithmetic: small float gemm = row . column
// accumulated left to right in float (pinned by tests/golden: cv2.gemm), MatExpr scaling multiplies by the scalar rounded to float (cvtScale32f),
// a / s = a * (1. / s), cv::norm and Mat::dot accumulate in double.
inline Mat operator*(const Mat& a, const Mat& b) {
    Mat m(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < b.cols; ++c) { float s = a.at<float>(r, 0) * b.at<float>(0, c); for (int k = 1; k < a.cols; ++k) s = s + a.at<float>(r, k) * b.at<float>(k, c); m.at<float>(r, c) = s; }
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { Mat m(a.rows, a.cols, CV_32F); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) + b.at<float>(r, c); return m; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat m(a.rows, a.cols, CV_32F); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c); return m; }
inline Mat operator-(const Mat& a) { Mat m(a.rows, a.cols, CV_32F); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = -a.at<float>(r, c); return m; }
inline Mat operator*(const Mat& a, double s) { Mat m(a.rows, a.cols, CV_32F); const float sf = (float)s; for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) * sf; return m; }
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator/(const Mat& a, double s) { return a * (1. / s); }
inline double norm(const Mat& a) { double s = 0; for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) { const double v = a.at<float>(r, c); s += v * v; } return std::sqrt(s); }
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
protected:
    const Mat* m_ = nullptr;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m), w_(&m) {}
    void create(int rows, int cols, int type) const { if (w_) w_->create(rows, cols, type); }
    void release() const { if (w_) w_->release(); }
private:
    Mat* w_ = nullptr;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
namespace line_descriptor {
struct KeyLine {                       // Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:105-145
    float angle; int class_id; int octave; Point2f pt; float response; float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int numOfPixels;
};
}  // namespace line_descriptor
}  // namespace cv
#endif
