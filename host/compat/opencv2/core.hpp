// Minimal stand-in for the OpenCV types that appear in the SIGNATURES of the hot-path classes
// (cv::Mat, cv::KeyPoint, cv::DMatch, cv::Point*, cv::Vec*).  Selected only when the build has no
// real OpenCV (this image: SURVEY section 8c); never mixed with it.  Layouts follow OpenCV 3.4:
// KeyPoint 28 bytes, DMatch 16 bytes, Mat row-major with a byte step.
#pragma once
#ifndef ERP_OPENCV_COMPAT
#define ERP_OPENCV_COMPAT 1
#endif

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef int64_t int64;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b) { static_assert(N == 2, ""); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < N; i++) r.val[i] = val[i] - o.val[i]; return r; }
    Vec operator+(const Vec& o) const { Vec r; for (int i = 0; i < N; i++) r.val[i] = val[i] + o.val[i]; return r; }
};
typedef Vec<int, 2> Vec2i;
typedef Vec<double, 2> Vec2d;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

struct KeyPoint {          // 28 bytes, pt first
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s) : pt(x, y), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {            // 16 bytes
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");

// dense, reference-counted 2-D matrix (the subset the hot path touches)
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;        // bytes between rows

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    // borrowed storage (no ownership), as cv::Mat(rows, cols, type, void* data, size_t step)
    Mat(int r, int c, int type, void* ext, size_t step_bytes = 0)
        : rows(r), cols(c), data(static_cast<unsigned char*>(ext)), type_(type)
    {
        step = step_bytes ? step_bytes : (size_t)c * elemSize();
    }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        store_ = std::shared_ptr<unsigned char>(new unsigned char[step * (size_t)(r > 0 ? r : 0) + 16], std::default_delete<unsigned char[]>());
        data = store_.get();
        std::memset(data, 0, step * (size_t)(r > 0 ? r : 0));
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const
    {
        static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)sz[depth()] * channels();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step * (size_t)r); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + step * (size_t)r); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + m.step * r, data + step * r, (size_t)cols * elemSize());
        return m;
    }

private:
    int type_ = 0;
    std::shared_ptr<unsigned char> store_;
};

inline int64 getTickCount() { return (int64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline double getTickFrequency() { return 1e9; }

} // namespace cv
