/*
 * oracle/cvshim/opencv2/core.hpp -- minimal OpenCV 4.0 SOURCE-COMPATIBILITY shim.
 *
 * TEST INFRASTRUCTURE ONLY.  OpenCV C++ is not installed in this image, so the reference
 * (/root/reference/src/sift.cpp, include/sift.hpp) cannot be built as shipped.  This header declares just
 * enough of namespace cv -- with OpenCV 4.0 semantics restated from its public documentation -- for the
 * UNMODIFIED reference translation unit to compile into oracle/_ref/libsift_ref.so, and for the drop-in
 * host shim (sift-gpu_b200/host/sift_dropin.cpp) to be compile- and run-checked here.  On a box with real
 * OpenCV 4 this directory is simply left off the include path.
 *
 * Numerical primitives (cvRound, fastAtan2, Matx solve, ...) forward to oracle/oracle_prims.h, the same
 * restatement the C oracle uses, so oracle-vs-_ref comparisons isolate sift.cpp's own logic.
 */
#ifndef CVSHIM_OPENCV2_CORE_HPP_
#define CVSHIM_OPENCV2_CORE_HPP_

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <map>
#include <vector>

#include "oracle_prims.h"

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 511) + 1)

typedef unsigned char uchar;

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
#define CV_Assert(expr) \
    do { if (!(expr)) throw cv::Exception(std::string("CV_Assert failed: ") + #expr); } while (0)
#define CV_Error(code, msg) throw cv::Exception(msg)

enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_CHOLESKY = 3 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_L1 = 2, NORM_L2 = 4 };

static inline int cvRound(double v) { return oracle_cv_round(v); }
static inline int cvRound(float v) { return oracle_cv_round((double)v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return oracled_cv_floor(v); }
static inline int cvFloor(float v) { return oraclef_cv_floor(v); }
static inline int cvFloor(int v) { return v; }
template <typename T> static inline T saturate_cast(float v);
template <> inline uchar saturate_cast<uchar>(float v) { return oraclef_saturate_u8(v); }

static inline int64_t getTickCount() {
    return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline double getTickFrequency() { return 1e9; }

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

/* cv::KeyPoint: same field order / defaults as OpenCV (28-byte POD: pt, size, angle, response, octave, class_id). */
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

/* Matx: fixed-size row-major matrix; solve()/dot() follow OpenCV's Matx_FastSolveOp / Matx::dot. */
template <typename T, int m, int n> class Matx {
public:
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
    Matx(T v0, T v1, T v2) { static_assert(m * n == 3, "3 values"); val[0] = v0; val[1] = v1; val[2] = v2; }
    Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8) {
        static_assert(m * n == 9, "9 values");
        T v[9] = {v0, v1, v2, v3, v4, v5, v6, v7, v8};
        for (int i = 0; i < 9; ++i) val[i] = v[i];
    }
    const T& operator()(int i, int j) const { return val[i * n + j]; }
    T& operator()(int i, int j) { return val[i * n + j]; }
    T dot(const Matx<T, m, n>& M) const {
        T s = 0;
        for (int i = 0; i < m * n; ++i) s += val[i] * M.val[i];
        return s;
    }
    template <int l> Matx<T, n, l> solve(const Matx<T, m, l>& rhs, int method = DECOMP_LU) const;
};
template <typename T, int cn> class Vec : public Matx<T, cn, 1> {
public:
    Vec() {}
    Vec(T v0, T v1, T v2) : Matx<T, cn, 1>(v0, v1, v2) {}
    Vec(const Matx<T, cn, 1>& a) : Matx<T, cn, 1>(a) {}
    const T& operator[](int i) const { return this->val[i]; }
    T& operator[](int i) { return this->val[i]; }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 3, 1> Matx31f;
typedef Vec<float, 3> Vec3f;

template <> template <> inline Matx<float, 3, 1> Matx<float, 3, 3>::solve<1>(const Matx<float, 3, 1>& rhs, int method) const {
    CV_Assert(method == DECOMP_LU || method == DECOMP_CHOLESKY);
    Matx<float, 3, 1> x;
    oraclef_solve3(val, rhs.val, x.val);
    return x;
}

template <typename T> class Scalar_ { public: T val[4]; Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; } };
typedef Scalar_<double> Scalar;

struct MatSize {
    const int* p;
    explicit MatSize(const int* p_) : p(p_) {}
    bool operator==(const MatSize& o) const { return p[0] == o.p[0] && p[1] == o.p[1]; }
    bool operator!=(const MatSize& o) const { return !(*this == o); }
};

/* cv::Mat: 2-D, reference-counted, always continuous (step == cols*elemSize).  Copying is shallow, as in
 * OpenCV.  Doubles as InputArray/OutputArray (getMat / create), see the typedefs below. */
class Mat {
public:
    int flags, rows, cols;
    uchar* data;
    MatSize size;
    Mat() : flags(0), rows(0), cols(0), data(nullptr), size(&rows) {}
    Mat(int r, int c, int type) : flags(0), rows(0), cols(0), data(nullptr), size(&rows) { create(r, c, type); }
    /* user-data header (no copy, not owned) */
    Mat(int r, int c, int type, void* ext) : flags(type), rows(r), cols(c), data((uchar*)ext), size(&rows) {}
    Mat(const Mat& o) : flags(o.flags), rows(o.rows), cols(o.cols), data(o.data), size(&rows), buf_(o.buf_) {}
    Mat& operator=(const Mat& o) {
        if (this != &o) { flags = o.flags; rows = o.rows; cols = o.cols; data = o.data; buf_ = o.buf_; }
        return *this;
    }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()] * channels(); }
    size_t step1() const { return (size_t)cols * channels(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags && buf_) return;
        flags = type; rows = r; cols = c;
        size_t bytes = (size_t)r * c * elemSize();
        buf_ = std::shared_ptr<uchar>(bytes ? (uchar*)std::malloc(bytes) : nullptr, std::free);
        data = buf_.get();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; }
    Mat clone() const {
        Mat m(rows, cols, flags);
        if (data) std::memcpy(m.data, data, (size_t)rows * cols * elemSize());
        return m;
    }
    template <typename T> T& at(int r, int c) { return ((T*)data)[(size_t)r * cols + c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)data)[(size_t)r * cols + c]; }
    template <typename T> T* ptr(int r = 0) { return (T*)data + (size_t)r * cols; }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)data + (size_t)r * cols; }
    /* _InputArray / _OutputArray surface used by the reference */
    Mat getMat() const { return *this; }
    void convertTo(Mat& dst, int type) const {
        CV_Assert(channels() == 1 && CV_MAT_CN(type) == 1);
        Mat out(rows, cols, type);
        size_t n = total();
        if (depth() == CV_8U && CV_MAT_DEPTH(type) == CV_32F) for (size_t i = 0; i < n; ++i) ((float*)out.data)[i] = (float)data[i];
        else if (depth() == CV_32F && CV_MAT_DEPTH(type) == CV_32F) std::memcpy(out.data, data, n * 4);
        else CV_Error(0, "cvshim: convertTo combination not supported");
        dst = out;
    }

private:
    std::shared_ptr<uchar> buf_;
};

/* MatExpr `a - b` (src/sift.cpp:280), CV_32FC1 only. */
static inline Mat operator-(const Mat& a, const Mat& b) {
    CV_Assert(a.rows == b.rows && a.cols == b.cols && a.type() == CV_32FC1 && b.type() == CV_32FC1);
    Mat d(a.rows, a.cols, CV_32FC1);
    const float *pa = (const float*)a.data, *pb = (const float*)b.data;
    float* pd = (float*)d.data;
    for (size_t i = 0, n = a.total(); i < n; ++i) pd[i] = pa[i] - pb[i];
    return d;
}

typedef const Mat& InputArray;
typedef Mat& OutputArray;

/* cv::resize, INTER_NEAREST only: dst(y,x) = src(min(floor(y*fy), rows-1), min(floor(x*fx), cols-1)),
 * fx = 1/((double)dcols/scols) -- OpenCV's resizeNN index rule (SURVEY App. B pins it against cv2 4.13). */
static inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR) {
    CV_Assert(interpolation == INTER_NEAREST && src.type() == CV_32FC1);
    if (dsize.width <= 0 || dsize.height <= 0) throw Exception("cvshim resize: empty dsize");
    Mat out(dsize.height, dsize.width, src.type());
    double ifx = 1. / ((double)dsize.width / src.cols), ify = 1. / ((double)dsize.height / src.rows);
    for (int y = 0; y < dsize.height; ++y) {
        int sy = std::min(cvFloor(y * ify), src.rows - 1);
        for (int x = 0; x < dsize.width; ++x) {
            int sx = std::min(cvFloor(x * ifx), src.cols - 1);
            out.at<float>(y, x) = src.at<float>(sy, sx);
        }
    }
    dst = out;
}

template <typename T, size_t fixed = 1024 / sizeof(T) + 8> class AutoBuffer {
public:
    explicit AutoBuffer(size_t n) : v_(n) {}
    operator T*() { return v_.data(); }
    operator const T*() const { return v_.data(); }
    T* data() { return v_.data(); }
private:
    std::vector<T> v_;
};

/* TLSData<T>: one T per thread, gather() returns them in creation order. */
template <typename T> class TLSData {
public:
    T* get() const {
        std::lock_guard<std::mutex> g(mu_);
        auto id = std::this_thread::get_id();
        auto it = idx_.find(id);
        if (it == idx_.end()) { slots_.emplace_back(new T()); it = idx_.emplace(id, slots_.size() - 1).first; }
        return slots_[it->second].get();
    }
    void gather(std::vector<T*>& out) const {
        std::lock_guard<std::mutex> g(mu_);
        out.clear();
        for (auto& s : slots_) out.push_back(s.get());
    }
private:
    mutable std::mutex mu_;
    mutable std::map<std::thread::id, size_t> idx_;
    mutable std::vector<std::unique_ptr<T>> slots_;
};

template <typename T> using Ptr = std::shared_ptr<T>;

namespace hal {
static inline void exp32f(const float* src, float* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = oraclef_exp(src[i]); }
static inline void fastAtan2(const float* y, const float* x, float* dst, int n, bool angleInDegrees) {
    for (int i = 0; i < n; ++i) { float a = oraclef_fast_atan2(y[i], x[i]); dst[i] = angleInDegrees ? a : a * (float)(CV_PI / 180); }
}
static inline void magnitude32f(const float* x, const float* y, float* dst, int n) { for (int i = 0; i < n; ++i) dst[i] = oraclef_magnitude(x[i], y[i]); }
}  // namespace hal

/* third-party CPU SIFT (cv::xfeatures2d::SIFT): out of scope.  Declared so that
 * SITF_BuildIn_OpenCV compiles; calling it throws. */
namespace xfeatures2d {
class SIFT {
public:
    static Ptr<SIFT> create() { return Ptr<SIFT>(new SIFT()); }
    void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false) {
        throw Exception("cvshim: cv::xfeatures2d::SIFT is not available without OpenCV contrib");
    }
    void clear() {}
};
}  // namespace xfeatures2d

}  // namespace cv
#endif
