// Minimal stand-in for the part of OpenCV (>= 4.2) the reference's front-end path touches.  TEST INFRASTRUCTURE ONLY
// (oracle/ref_build.py compiles the reference's own sources against it; the OpenCV C++ headers are not in this image).
// The four numerical routines the extractor calls (initUndistortRectifyMap, undistortPoints and their fisheye twins,
// remap INTER_LINEAR) forward to the restatements in oracle/ppg_oracle.c, which tests/test_oracle_cv.py pins bit for
// bit against cv2 4.13 outputs (tests/golden/cv_kat.npz).  Everything else is plain container plumbing.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

extern "C" {
void ppgo_undistort_points_pinhole(const float* K, const float* D, const float* xy, int n, float* out);
void ppgo_undistort_points_fisheye(const float* K, const float* D, const float* xy, int n, float* out);
void ppgo_init_undistort_map_pinhole(const float* K, const float* D, int W, int H, float* mx, float* my);
void ppgo_init_undistort_map_fisheye(const float* K, const float* D, int W, int H, float* mx, float* my);
void ppgo_remap_linear(const float* src, int W, int H, const float* mx, const float* my, float* dst);
}

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_2PI 6.283185307179586476925286766559

namespace cv {

typedef unsigned char uchar;
typedef std::string String;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
struct Range {
    int start, end;
};
class noArrayT {};
inline noArrayT noArray() { return noArrayT(); }

enum { INTER_NEAREST = 0, INTER_LINEAR = 1, NORM_MINMAX = 32, BORDER_CONSTANT = 0 };

class Mat {
   public:
    int rows = 0, cols = 0, flags = 0;
    uchar* data = nullptr;
    size_t step = 0;  // bytes per row
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext) : rows(r), cols(c), flags(type), data((uchar*)ext) {
        step = (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        flags = type;
        step = (size_t)c * elemSize();
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step, (uchar)0);
        data = buf->data();
    }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> 3) + 1; }
    size_t elemSize1() const {
        switch (depth()) {
            case CV_8U: return 1;
            case CV_64F: return 8;
            default: return 4;
        }
    }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    Size size() const { return Size(cols, rows); }
    template <typename T>
    T& at(int i, int j) {
        return *reinterpret_cast<T*>(data + (size_t)i * step + (size_t)j * sizeof(T));
    }
    template <typename T>
    const T& at(int i, int j) const {
        return *reinterpret_cast<const T*>(data + (size_t)i * step + (size_t)j * sizeof(T));
    }
    template <typename T>
    T& at(int i) {
        return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
    }
    template <typename T>
    const T& at(int i) const {
        return rows == 1 ? at<T>(0, i) : at<T>(i, 0);
    }
    template <typename T>
    T* ptr(int i = 0) {
        return reinterpret_cast<T*>(data + (size_t)i * step);
    }
    template <typename T>
    const T* ptr(int i = 0) const {
        return reinterpret_cast<const T*>(data + (size_t)i * step);
    }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    // same data, other channel count: r x c x cn  ->  r x (c * cn / new_cn) x new_cn
    Mat reshape(int cn, int new_rows = 0) const {
        Mat m = *this;
        const int total_scalars = cols * channels();
        assert(total_scalars % cn == 0 && new_rows == 0);
        (void)new_rows;
        m.cols = total_scalars / cn;
        m.flags = CV_MAKETYPE(depth(), cn);
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, flags);
        for (int i = 0; i < rows; i++)
            memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& o) const { o = clone(); }
    Mat row(int i) const {
        Mat m = *this;
        m.rows = 1;
        m.data = data + (size_t)i * step;
        return m;
    }
    Mat rowRange(int a, int b) const {
        Mat m = *this;
        m.rows = b - a;
        m.data = data + (size_t)a * step;
        return m;
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < (r < c ? r : c); i++) m.set(i, i, 1.0);
        return m;
    }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) m.set(i, j, 1.0);
        return m;
    }
    void set(int i, int j, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(i, j) = (uchar)v; break;
            case CV_64F: at<double>(i, j) = v; break;
            case CV_32S: at<int>(i, j) = (int)v; break;
            default: at<float>(i, j) = (float)v; break;
        }
    }
    double get(int i, int j) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(i, j);
            case CV_64F: return at<double>(i, j);
            case CV_32S: return at<int>(i, j);
            default: return at<float>(i, j);
        }
    }
    // debug-only arithmetic (PPGExtractor::showTensor): saturating conversion is not modelled
    void convertTo(Mat& o, int type, double a = 1.0, double b = 0.0) const {
        Mat m(rows, cols, type);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) m.set(i, j, get(i, j) * a + b);
        o = m;
    }
    Mat operator*(double s) const {
        Mat m = clone();
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) m.set(i, j, get(i, j) * s);
        return m;
    }
};

// cv::Mat_<float>(r, c) << a, b, ...  (sensors/src/Pinhole.cpp:70, :76): row-major fill
template <typename T>
class Mat_ : public Mat {
   public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {}
};
template <typename T>
struct MatCommaInit {
    Mat m;
    int k;
    MatCommaInit operator,(T v) {
        m.set(k / m.cols, k % m.cols, (double)v);
        return MatCommaInit{m, k + 1};
    }
    operator Mat() const { return m; }
};
template <typename T>
inline MatCommaInit<T> operator<<(const Mat_<T>& m, T v) {
    Mat mm = m;
    mm.set(0, 0, (double)v);
    return MatCommaInit<T>{mm, 1};
}

// camera matrix K (3x3 CV_32F) / distortion D (4x1 CV_32F) -> the plain arrays ppg_oracle.c takes
inline void kd_arrays(const Mat& K, const Mat& D, float* k9, float* d4) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) k9[3 * i + j] = (float)K.get(i, j);
    for (int i = 0; i < 4; i++) d4[i] = (float)(D.rows == 1 ? D.get(0, i) : D.get(i, 0));
}
// cv::undistortPoints(src, dst, K, D, R = noArray / Mat(), P = K) on an N x 1 CV_32FC2 array (R must be empty / identity)
template <typename RT>
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const RT&, const Mat& P) {
    (void)P;
    float k9[9], d4[4];
    kd_arrays(K, D, k9, d4);
    const int n = src.rows * src.cols;
    Mat out(src.rows, src.cols, CV_32FC2);
    ppgo_undistort_points_pinhole(k9, d4, src.ptr<float>(), n, out.ptr<float>());
    dst = out;
}
inline void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size sz, int m1type,
                                    Mat& map1, Mat& map2) {
    (void)R;
    (void)newK;
    assert(m1type == CV_32F);
    (void)m1type;
    float k9[9], d4[4];
    kd_arrays(K, D, k9, d4);
    map1 = Mat(sz.height, sz.width, CV_32F);
    map2 = Mat(sz.height, sz.width, CV_32F);
    ppgo_init_undistort_map_pinhole(k9, d4, sz.width, sz.height, map1.ptr<float>(), map2.ptr<float>());
}
namespace fisheye {
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const Mat& R, const Mat& P) {
    (void)R;
    (void)P;
    float k9[9], d4[4];
    kd_arrays(K, D, k9, d4);
    const int n = src.rows * src.cols;
    Mat out(src.rows, src.cols, CV_32FC2);
    ppgo_undistort_points_fisheye(k9, d4, src.ptr<float>(), n, out.ptr<float>());
    dst = out;
}
inline void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size sz, int m1type,
                                    Mat& map1, Mat& map2) {
    (void)R;
    (void)newK;
    assert(m1type == CV_32F);
    (void)m1type;
    float k9[9], d4[4];
    kd_arrays(K, D, k9, d4);
    map1 = Mat(sz.height, sz.width, CV_32F);
    map2 = Mat(sz.height, sz.width, CV_32F);
    ppgo_init_undistort_map_fisheye(k9, d4, sz.width, sz.height, map1.ptr<float>(), map2.ptr<float>());
}
}  // namespace fisheye
// cv::remap(src, dst, mapx, mapy, INTER_LINEAR): CV_32FC1, border constant 0; in place allowed (src copied first)
inline void remap(const Mat& src, Mat& dst, const Mat& mapx, const Mat& mapy, int interp) {
    assert(interp == INTER_LINEAR && src.type() == CV_32F);
    (void)interp;
    Mat s = src.clone();
    Mat out(src.rows, src.cols, CV_32F);
    ppgo_remap_linear(s.ptr<float>(), src.cols, src.rows, mapx.ptr<float>(), mapy.ptr<float>(), out.ptr<float>());
    dst = out;
}
// debug helpers of PPGExtractor::showTensor (dead code, never called)
inline void normalize(const Mat&, Mat&, double, double, int) {}
inline void resize(const Mat&, Mat&, Size, int = INTER_LINEAR) {}
inline void resize(const Mat&, Mat&, Size, double, double, int = INTER_LINEAR) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

}  // namespace cv
