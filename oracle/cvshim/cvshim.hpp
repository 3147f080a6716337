// oracle/cvshim/cvshim.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal stand-in for the slice of the OpenCV C++ API that the reference's
// ORBextractor.cc / ORBmatcher.cc / Frame.cc use, so those files compile UNMODIFIED from
// /root/reference (OpenCV C++ headers and libraries do not exist in this image).
// Image primitives forward to oracle/cv_prims.c, which is pinned bit-for-bit to
// Python cv2 4.13.0 (tests/test_cv_prims.py).  Semantics follow OpenCV 4.13:
//  * Mat copies / ROIs share one ref-counted buffer (malloc'ed; never operator new,
//    so the harness's bump arena only sees the reference's own containers);
//  * OutputArray::create is a no-op when size and type already match, so resize()
//    and copyMakeBorder() write INTO an existing ROI (ORBextractor.cc:1166-1169);
//  * Mat::operator=(MatExpr) fills in place when the destination already has the
//    right shape (ORBextractor.cc:167 assigns zeros() to a rowRange view);
//  * small CV_32F algebra accumulates in float, k ascending, C added last, which is
//    what cv::gemm does for 3x3 * 3x1; cv::norm accumulates in double.
#ifndef ORACLE_CVSHIM_HPP
#define ORACLE_CVSHIM_HPP

#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <iostream>
#include <list>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../cv_prims.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

static inline int cvRound(double v) { return cvp_round(v); }
static inline int cvRound(float v) { return cvp_round((double)v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

using ::uchar;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b)
{ a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0,
             int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class Mat;

// zeros / ones / eye placeholder: materialises on assignment or conversion
struct MatExpr {
    int rows, cols, type, kind; // kind: 0 zeros, 1 ones, 2 eye
    operator Mat() const;
};

struct MatStep {
    size_t v;
    MatStep() : v(0) {}
    MatStep(size_t s) : v(s) {}
    operator size_t() const { return v; }
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;

    Mat() : rows(0), cols(0), data(0), type_(CV_8U), blk_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(0), type_(CV_8U), blk_(0) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(0), type_(CV_8U), blk_(0) { create(s.height, s.width, type); }
    // header over user memory (not owned)
    Mat(int r, int c, int type, void* ext, size_t stp = 0)
        : rows(r), cols(c), data((uchar*)ext), type_(type), blk_(0) { step = stp ? stp : (size_t)c * esz(type); }
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), type_(m.type_), blk_(m.blk_) { retain(); }
    ~Mat() { drop(); }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) {
            m.retain();
            drop();
            rows = m.rows; cols = m.cols; data = m.data; step = m.step; type_ = m.type_; blk_ = m.blk_;
        }
        return *this;
    }
    Mat& operator=(const MatExpr& e)
    {
        create(e.rows, e.cols, e.type); // no-op if the shape matches: fill in place
        fill(e.kind);
        return *this;
    }

    static size_t esz(int type) { return type == CV_8U ? 1 : type == CV_64F ? 8 : 4; }
    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && type_ == type) return;
        drop();
        rows = r; cols = c; type_ = type;
        step = (size_t)c * esz(type);
        size_t bytes = (size_t)r * step.v;
        blk_ = (int*)std::malloc(bytes + 64);
        *blk_ = 1;
        data = (uchar*)blk_ + 64;
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { drop(); rows = cols = 0; data = 0; blk_ = 0; }

    int type() const { return type_; }
    int depth() const { return type_; }
    int channels() const { return 1; }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t elemSize() const { return esz(type_); }
    size_t step1() const { return step.v / esz(type_); }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step.v == (size_t)cols * esz(type_); }

    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.v))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.v))[c]; }
    // single index: element i of a row or column vector
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }

    Mat operator()(const Rect& r) const
    {
        Mat m(*this);
        m.data = data + (size_t)r.y * step.v + (size_t)r.x * esz(type_);
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }

    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_));
        return m;
    }
    void copyTo(Mat& dst) const
    {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step.v == step.v) return;
        Mat src(*this); // keep alive if dst aliases
        dst.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memmove(dst.ptr(r), src.ptr(r), (size_t)cols * esz(type_));
    }
    void convertTo(Mat& dst, int rtype) const
    {
        Mat src(*this);
        Mat out(rows, cols, rtype); // fresh buffer: dst may be *this (Frame.cc:614)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                double v = src.type_ == CV_8U ? (double)src.at<uchar>(r, c)
                         : src.type_ == CV_32F ? (double)src.at<float>(r, c)
                         : src.type_ == CV_32S ? (double)src.at<int>(r, c) : src.at<double>(r, c);
                if (rtype == CV_32F) out.at<float>(r, c) = (float)v;
                else if (rtype == CV_64F) out.at<double>(r, c) = v;
                else if (rtype == CV_32S) out.at<int>(r, c) = cvRound(v);
                else out.at<uchar>(r, c) = (uchar)std::min(255, std::max(0, cvRound(v)));
            }
        dst = out;
    }
    Mat t() const
    {
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (type_ == CV_32F) m.at<float>(c, r) = at<float>(r, c);
                else if (type_ == CV_64F) m.at<double>(c, r) = at<double>(r, c);
                else m.at<uchar>(c, r) = at<uchar>(r, c);
            }
        return m;
    }
    Mat reshape(int /*cn*/, int newrows = 0) const
    {
        assert(isContinuous());
        Mat m(*this);
        if (newrows > 0) { m.cols = (int)(total() / (size_t)newrows); m.rows = newrows; m.step = (size_t)m.cols * esz(type_); }
        return m;
    }
    double dot(const Mat& b) const
    {
        double s = 0;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)b.at<float>(r, c);
        return s;
    }

    static MatExpr zeros(int r, int c, int type) { MatExpr e = { r, c, type, 0 }; return e; }
    static MatExpr ones(int r, int c, int type) { MatExpr e = { r, c, type, 1 }; return e; }
    static MatExpr eye(int r, int c, int type) { MatExpr e = { r, c, type, 2 }; return e; }

    void fill(int kind)
    {
        for (int r = 0; r < rows; ++r) {
            std::memset(ptr(r), 0, (size_t)cols * esz(type_));
            for (int c = 0; c < cols; ++c) {
                if (kind == 1 || (kind == 2 && r == c)) {
                    if (type_ == CV_32F) at<float>(r, c) = 1.f;
                    else if (type_ == CV_64F) at<double>(r, c) = 1.0;
                    else if (type_ == CV_32S) at<int>(r, c) = 1;
                    else at<uchar>(r, c) = 1;
                }
            }
        }
    }

private:
    int type_;
    int* blk_; // ref-counted malloc block; 0 for external memory
    void retain() const { if (blk_) __atomic_add_fetch(blk_, 1, __ATOMIC_RELAXED); }
    void drop()
    {
        if (blk_ && __atomic_sub_fetch(blk_, 1, __ATOMIC_ACQ_REL) == 0) std::free(blk_);
        blk_ = 0;
    }
};

inline MatExpr::operator Mat() const { Mat m(rows, cols, type); m.fill(kind); return m; }

// Mat_<float>(r,c) << a, b, c   (Frame.cc:738)
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_64F) {}
};
template <typename T> struct MatCommaInit_ {
    Mat_<T> m; int i;
    MatCommaInit_(const Mat_<T>& _m) : m(_m), i(0) {}
    MatCommaInit_& operator,(T v) { m.template at<T>(i / m.cols, i % m.cols) = v; ++i; return *this; }
    operator Mat() const { return m; }
};
template <typename T> static inline MatCommaInit_<T> operator<<(const Mat_<T>& m, T v)
{ MatCommaInit_<T> ci(m); return (ci, v); }

// ---- small float algebra (CV_32F only) -----------------------------------------
static inline Mat operator*(const Mat& a, const Mat& b)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.cols == b.rows);
    Mat d(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            float s = 0.f;
            for (int k = 0; k < a.cols; ++k) {
                float p = a.at<float>(i, k) * b.at<float>(k, j);
                s = k == 0 ? p : s + p;
            }
            d.at<float>(i, j) = s;
        }
    return d;
}
#define CVSHIM_EW(NAME, EXPR)                                                        \
    static inline Mat NAME(const Mat& a, const Mat& b)                               \
    {                                                                                \
        assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols); \
        Mat d(a.rows, a.cols, CV_32F);                                               \
        for (int i = 0; i < a.rows; ++i)                                             \
            for (int j = 0; j < a.cols; ++j) {                                       \
                float x = a.at<float>(i, j), y = b.at<float>(i, j);                  \
                d.at<float>(i, j) = (EXPR);                                          \
            }                                                                        \
        return d;                                                                    \
    }
CVSHIM_EW(operator+, x + y)
CVSHIM_EW(operator-, x - y)
#undef CVSHIM_EW
static inline Mat scaled_(const Mat& a, double s)
{
    Mat d(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) d.at<float>(i, j) = (float)((double)a.at<float>(i, j) * s);
    return d;
}
static inline Mat operator*(double s, const Mat& a) { return scaled_(a, s); }
static inline Mat operator*(const Mat& a, double s) { return scaled_(a, s); }
static inline Mat operator/(const Mat& a, double s) { return scaled_(a, 1.0 / s); }
static inline Mat operator-(const Mat& a) { return scaled_(a, -1.0); }

static inline double norm(const Mat& a, int type = NORM_L2)
{
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            double v = a.type() == CV_32F ? (double)a.at<float>(i, j) : a.type() == CV_64F ? a.at<double>(i, j) : (double)a.at<uchar>(i, j);
            s += type == NORM_L1 ? std::fabs(v) : v * v;
        }
    return type == NORM_L1 ? s : std::sqrt(s);
}
static inline double norm(const Mat& a, const Mat& b, int type = NORM_L2)
{
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            double v = a.type() == CV_32F ? (double)a.at<float>(i, j) - (double)b.at<float>(i, j)
                                          : (double)a.at<uchar>(i, j) - (double)b.at<uchar>(i, j);
            s += type == NORM_L1 ? std::fabs(v) : v * v;
        }
    return type == NORM_L1 ? s : std::sqrt(s);
}

// ---- array proxies ---------------------------------------------------------------
class _InputArray {
public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    Mat getMat() const { return *m_; }
    Mat& getMatRef() const { return *m_; }
    bool empty() const { return m_->empty(); }
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size s, int type) const { m_->create(s.height, s.width, type); }
    void release() const { m_->release(); }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// ---- image primitives (forward to cv_prims.c) --------------------------------------
static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double = 0, double = 0,
                          int interpolation = INTER_LINEAR)
{
    assert(interpolation == INTER_LINEAR);
    (void)interpolation;
    Mat src = _src.getMat();
    assert(src.type() == CV_8U);
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    cvp_resize_linear_8u(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                                  int borderType)
{
    Mat src = _src.getMat();
    assert(src.type() == CV_8U && (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    assert(top == bottom && top == left && top == right);
    (void)bottom; (void)left; (void)right; (void)borderType;
    _dst.create(src.rows + 2 * top, src.cols + 2 * top, src.type());
    Mat dst = _dst.getMat();
    cvp_border_reflect101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top);
}

static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy = 0,
                                int borderType = BORDER_DEFAULT)
{
    Mat src = _src.getMat();
    assert(src.type() == CV_8U && ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 &&
           borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    cvp_gaussian7x7_s2(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}

static inline void FAST(InputArray _img, std::vector<KeyPoint>& kps, int threshold, bool nms = true)
{
    assert(nms);
    (void)nms;
    Mat img = _img.getMat();
    kps.clear();
    int cap = ((img.cols + 1) / 2) * ((img.rows + 1) / 2) + 16;
    cvp_corner* buf = (cvp_corner*)std::malloc(sizeof(cvp_corner) * (size_t)cap);
    int n = cvp_fast9_nms(img.data, img.cols, img.rows, img.step, threshold, buf, cap);
    assert(n <= cap);
    kps.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        kps.push_back(KeyPoint((float)buf[i].x, (float)buf[i].y, 7.f, -1.f, (float)buf[i].score));
    std::free(buf);
}

static inline float fastAtan2(float y, float x) { return cvp_fast_atan2(y, x); }

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) { std::abort(); } // only in dead ComputeKeyPointsOld
};

static inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&)
{ std::abort(); } // Frame.cc:456,485 -- never executed when distCoef[0]==0

} // namespace cv

#endif
