// cv_shim.h -- the sliver of the OpenCV 2.4 C++ API that the reference's hot-path translation units name,
// so that Duke/{mfreconstruct,reconstruct,graycodes,multifrequency,utilities,pointcloudimage,virtualcamera,
// stereorect}.cpp compile UNMODIFIED with g++ (oracle/Makefile target `ref`).  Ours, not the reference's and
// not OpenCV's: a reference-counted dense matrix, small fixed vectors/points with OpenCV's conversion rules,
// and inert stubs for the image-IO / calibration calls that the pinned functions never reach.
//
// Third-party arithmetic that the path does reach and that therefore stays UNPINNED (see README.md):
//   Mat * Mat      accumulates in double, left to right, narrowed to the matrix type
//   Vec / float    multiplies by (1.f / alpha), as OpenCV 2.4's operator/ does
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;

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
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)

struct CvScalar { double val[4]; };
struct CvSize { int width, height; };
struct _IplImage;
typedef struct _IplImage IplImage;
static inline CvScalar cvScalar(double a, double b = 0, double c = 0, double d = 0) { CvScalar s = {{a, b, c, d}}; return s; }
static inline CvSize cvSize(int w, int h) { CvSize s = {w, h}; return s; }

namespace cv {

using std::vector;

static inline void shim_unreachable(const char *what)
{
    fprintf(stderr, "cv_shim: %s is not implemented (not on the pinned hot path)\n", what);
    abort();
}

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { long i = lrint(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline schar saturate_cast<schar>(double v) { long i = lrint(v); return (schar)(i < -128 ? -128 : i > 127 ? 127 : i); }
template <> inline ushort saturate_cast<ushort>(double v) { long i = lrint(v); return (ushort)(i < 0 ? 0 : i > 65535 ? 65535 : i); }
template <> inline short saturate_cast<short>(double v) { long i = lrint(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }
template <> inline int saturate_cast<int>(double v) { return (int)lrint(v); }
template <> inline float saturate_cast<float>(double v) { return (float)v; }
template <> inline double saturate_cast<double>(double v) { return v; }

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    Scalar(const CvScalar &s) { for (int i = 0; i < 4; i++) val[i] = s.val[i]; }
};
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    Size(const CvSize &s) : width(s.width), height(s.height) {}
};

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a) { for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; }
    Vec(T a, T b, T c) { for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; val[2] = c; }
    template <typename U> Vec(const Vec<U, N> &o) { for (int i = 0; i < N; i++) val[i] = saturate_cast<T>((double)o.val[i]); }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    T dot(const Vec &o) const { T s = 0; for (int i = 0; i < N; i++) s += val[i] * o.val[i]; return s; }  // Matx::dot
};
template <typename T, int N> static inline Vec<T, N> operator+(const Vec<T, N> &a, const Vec<T, N> &b) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>((double)a.val[i] + (double)b.val[i]); return r; }
template <typename T, int N> static inline Vec<T, N> operator-(const Vec<T, N> &a, const Vec<T, N> &b) { Vec<T, N> r; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>((double)a.val[i] - (double)b.val[i]); return r; }
// OpenCV 2.4: operator/(Vec, float alpha) == Vec(a, 1.f/alpha, Matx_ScaleOp())
template <typename T, int N> static inline Vec<T, N> operator/(const Vec<T, N> &a, float alpha) { Vec<T, N> r; const float inv = 1.f / alpha; for (int i = 0; i < N; i++) r.val[i] = saturate_cast<T>(a.val[i] * inv); return r; }
typedef Vec<uchar, 3> Vec3b;
typedef Vec<ushort, 3> Vec3w;
typedef Vec<short, 3> Vec3s;
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    Point3_(const Vec<T, 3> &v) : x(v.val[0]), y(v.val[1]), z(v.val[2]) {}
    template <typename U> Point3_(const Vec<U, 3> &v) : x(saturate_cast<T>((double)v.val[0])), y(saturate_cast<T>((double)v.val[1])), z(saturate_cast<T>((double)v.val[2])) {}
    operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
};
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
template <typename T> static inline Point3_<T> operator+(const Point3_<T> &a, const Point3_<T> &b) { return Point3_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y), saturate_cast<T>(a.z + b.z)); }
template <typename T> static inline Point3_<T> operator-(const Point3_<T> &a, const Point3_<T> &b) { return Point3_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y), saturate_cast<T>(a.z - b.z)); }
template <typename T> static inline Point3_<T> operator*(float a, const Point3_<T> &b) { return Point3_<T>(saturate_cast<T>(a * b.x), saturate_cast<T>(a * b.y), saturate_cast<T>(a * b.z)); }
template <typename T> static inline Point3_<T> operator*(double a, const Point3_<T> &b) { return Point3_<T>(saturate_cast<T>(a * b.x), saturate_cast<T>(a * b.y), saturate_cast<T>(a * b.z)); }
template <typename T> static inline Point3_<T> operator*(const Point3_<T> &b, float a) { return a * b; }

enum { INTER_LINEAR = 1 };

class Mat {
public:
    int rows, cols;
    unsigned char *data;
    size_t step;

    Mat() : rows(0), cols(0), data(0), step(0), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); *this = s; }
    Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((unsigned char *)ext), type_(type) { step = (size_t)c * elemSize(); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    // Mat(const IplImage*) without making `mat = NULL` ambiguous (NULL is not an IplImage* for deduction)
    template <typename P, typename = typename std::enable_if<std::is_same<P, IplImage *>::value>::type>
    Mat(P) : rows(0), cols(0), data(0), step(0), type_(0) { shim_unreachable("Mat(IplImage*)"); }

    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<unsigned char> >((size_t)r * step + 16, (unsigned char)0xCD);
        data = buf_->data();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize1() const { static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 0}; return (size_t)sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == 0 || rows * cols == 0; }
    void release() { buf_.reset(); data = 0; rows = cols = 0; step = 0; }

    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *(const T *)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T &at(int i) { if (rows == 1) return at<T>(0, i); if (cols == 1) return at<T>(i, 0); return at<T>(i / cols, i % cols); }
    template <typename T> T &at(int y, int x, int ch) { return *(T *)(data + (size_t)y * step + ((size_t)x * channels() + ch) * sizeof(T)); }

    // Mat = scalar: fill (cv::Mat::operator=(const Scalar&)); on an empty Mat (e.g. `mask = NULL;`) a no-op
    Mat &operator=(const Scalar &s)
    {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int ch = 0; ch < channels(); ch++) set(r, c, ch, s.val[ch]);
        return *this;
    }
    double get(int r, int c, int ch = 0) const
    {
        const unsigned char *p = data + (size_t)r * step + ((size_t)c * channels() + ch) * elemSize1();
        switch (depth()) {
        case CV_8U: return *(const uchar *)p;
        case CV_8S: return *(const schar *)p;
        case CV_16U: return *(const ushort *)p;
        case CV_16S: return *(const short *)p;
        case CV_32S: return *(const int *)p;
        case CV_32F: return *(const float *)p;
        default: return *(const double *)p;
        }
    }
    void set(int r, int c, int ch, double v)
    {
        unsigned char *p = data + (size_t)r * step + ((size_t)c * channels() + ch) * elemSize1();
        switch (depth()) {
        case CV_8U: *(uchar *)p = saturate_cast<uchar>(v); break;
        case CV_8S: *(schar *)p = saturate_cast<schar>(v); break;
        case CV_16U: *(ushort *)p = saturate_cast<ushort>(v); break;
        case CV_16S: *(short *)p = saturate_cast<short>(v); break;
        case CV_32S: *(int *)p = saturate_cast<int>(v); break;
        case CV_32F: *(float *)p = (float)v; break;
        default: *(double *)p = v; break;
        }
    }
    Mat t() const
    {
        Mat r(cols, rows, type_);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) r.set(j, i, 0, get(i, j));
        return r;
    }
    Mat operator-() const
    {
        Mat r(rows, cols, type_);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) r.set(i, j, 0, -get(i, j));
        return r;
    }
    Mat &operator-=(double) { shim_unreachable("Mat -= scalar"); return *this; }
    Mat &operator*=(double) { shim_unreachable("Mat *= scalar"); return *this; }

private:
    int type_;
    std::shared_ptr<std::vector<unsigned char> > buf_;
};

// matrix product: double accumulator, k ascending, narrowed to the left operand's type (UNPINNED third-party op)
static inline Mat operator*(const Mat &a, const Mat &b)
{
    Mat r(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.get(i, k) * b.get(k, j);
            r.set(i, j, 0, s);
        }
    return r;
}

static inline Mat imread(const std::string &, int = 1) { return Mat(); }
static inline bool imwrite(const std::string &, const Mat &) { return true; }
static inline void remap(const Mat &, Mat &, const Mat &, const Mat &, int) { shim_unreachable("cv::remap"); }
static inline void stereoRectify(const Mat &, const Mat &, const Mat &, const Mat &, Size, const Mat &, const Mat &, Mat &, Mat &,
                                 Mat &, Mat &, Mat &, int, double) { shim_unreachable("cv::stereoRectify"); }
static inline void initUndistortRectifyMap(const Mat &, const Mat &, const Mat &, const Mat &, Size, int, Mat &, Mat &) { shim_unreachable("cv::initUndistortRectifyMap"); }
static inline void split(const Mat &, std::vector<Mat> &) { shim_unreachable("cv::split"); }
static inline void merge(const std::vector<Mat> &, Mat &) { shim_unreachable("cv::merge"); }
static inline void minMaxIdx(const Mat &, double *, double *) { shim_unreachable("cv::minMaxIdx"); }

}  // namespace cv
