// Minimal stand-in for the part of Eigen (>= 3.3) the reference's front-end path touches.  TEST INFRASTRUCTURE ONLY
// (oracle/ref_build.py compiles the reference's own sources against it; Eigen itself is not in this image).
//
// What it pins and what it does not: the reference's control flow (NMS walk, overlap filter, line scoring,
// colinearity, ExtendMapMatches) is compiled from /root/reference unmodified; the ARITHMETIC of the few Eigen
// expressions on the path is this file's reading of Eigen's semantics, stated here once:
//   * coefficient-wise expressions are evaluated per coefficient, left to right as written, in the matrix's Scalar;
//     a scalar operand of another arithmetic type is converted to Scalar first (Eigen's promote_scalar_arg), so
//     `v * 0.2` on a Vector2f multiplies by (float)0.2;
//   * norm() = sqrt(squaredNorm()), squaredNorm() = sum of coefficient squares from index 0 upwards in Scalar (exact
//     for the 2-vectors of the extractor; for the 256-wide descriptor difference Eigen's vectorised reduction order is
//     unspecified -- DescriptorDistance below uses the fixed order the oracle and the CUDA kernels share);
//   * no FMA contraction (the TU is built with -ffp-contract=off, as the reference's CMakeLists.txt sets no -march).
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3

namespace Eigen {

constexpr int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
typedef std::ptrdiff_t Index;

template <typename T>
using aligned_allocator = std::allocator<T>;

namespace detail {
template <typename S, int R, int C, bool Fixed = (R > 0 && C > 0)>
struct Store {
    S d[R * C];
    Store() {
        for (int i = 0; i < R * C; i++) d[i] = S();
    }
    int rows() const { return R; }
    int cols() const { return C; }
    void resize(int r, int c) { assert(r == R && c == C); }
    S* data() { return d; }
    const S* data() const { return d; }
};
template <typename S, int R, int C>
struct Store<S, R, C, false> {
    std::vector<S> d;
    int r_ = (R > 0 ? R : 0), c_ = (C > 0 ? C : 0);
    int rows() const { return r_; }
    int cols() const { return c_; }
    void resize(int r, int c) {
        r_ = r;
        c_ = c;
        d.assign((size_t)r * c, S());  // Eigen leaves the coefficients uninitialised; nothing on the path reads them
    }
    S* data() { return d.data(); }
    const S* data() const { return d.data(); }
};
}  // namespace detail

template <typename S, int R, int C, int Opt = 0, int MR = R, int MC = C>
class Matrix;

template <typename M>
struct CommaInit {
    M& m;
    int k;                        // scalars placed so far (row by row, as Eigen's comma initialiser)
    int br = 0, bc = 0, bh = 0;   // block cursor: top row, next column, height of the current band of blocks
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    CommaInit& operator,(T v) {
        const int c = m.cols();
        m(k / c, k % c) = (typename M::Scalar)v;
        k++;
        return *this;
    }
    // a matrix operand: blocks fill a band left to right, the next band starts below (KannalaBrandt8.cpp:203, :207:
    // `Tcw << R, t` on a 3 x 4)
    template <typename S2, int R2, int C2>
    CommaInit& operator,(const Matrix<S2, R2, C2>& b) {
        if (bc + b.cols() > m.cols()) {
            br += bh;
            bc = 0;
        }
        for (int i = 0; i < b.rows(); i++)
            for (int j = 0; j < b.cols(); j++) m(br + i, bc + j) = (typename M::Scalar)b(i, j);
        bh = b.rows();
        bc += b.cols();
        return *this;
    }
};

// `m.row(i)` of a non-const matrix: assignable, readable (KannalaBrandt8.cpp:211, :228-231)
template <typename M>
struct RowRef {
    M& m;
    int i;
    typedef typename M::Scalar S;
    typedef Matrix<S, 1, M::ColsAtCompileTime> Row;
    operator Row() const {
        Row r;
        for (int j = 0; j < m.cols(); j++) r[j] = m(i, j);
        return r;
    }
    RowRef& operator=(const Row& r) {
        for (int j = 0; j < m.cols(); j++) m(i, j) = r[j];
        return *this;
    }
    template <int N>
    S dot(const Matrix<S, N, 1>& o) const {  // (a0*b0 + a1*b1) + a2*b2, as Matrix::dot
        S acc = m(i, 0) * o[0];
        for (int j = 1; j < N; j++) acc = acc + m(i, j) * o[j];
        return acc;
    }
};

// `v.head(n) / s` with a run-time n, assigned to a fixed vector (KannalaBrandt8.cpp:235)
template <typename S>
struct HeadDyn {
    S v[4];
    int n;
    HeadDyn operator/(S s) const {
        HeadDyn r = *this;
        for (int i = 0; i < n; i++) r.v[i] = v[i] / s;
        return r;
    }
    template <int N>
    operator Matrix<S, N, 1, 0, N, 1>() const {
        Matrix<S, N, 1, 0, N, 1> r;
        for (int i = 0; i < N; i++) r[i] = v[i];
        return r;
    }
};

template <typename S, int R, int C, int Opt, int MR, int MC>
class Matrix {
   public:
    typedef S Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
    detail::Store<S, R, C> st;

    Matrix() {}
    Matrix(int r, int c) { st.resize(r, c); }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 2, int>::type = 0>
    Matrix(S a, S b) {
        st.d[0] = a;
        st.d[1] = b;
    }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 3, int>::type = 0>
    Matrix(S a, S b, S c) {
        st.d[0] = a;
        st.d[1] = b;
        st.d[2] = c;
    }
    template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 4 && (RR == 1 || CC == 1), int>::type = 0>
    Matrix(S a, S b, S c, S d) {
        st.d[0] = a;
        st.d[1] = b;
        st.d[2] = c;
        st.d[3] = d;
    }
    int rows() const { return st.rows(); }
    int cols() const { return st.cols(); }
    int size() const { return rows() * cols(); }
    void resize(int r, int c) { st.resize(r, c); }
    S* data() { return st.data(); }
    const S* data() const { return st.data(); }
    // column-major storage (Eigen's default)
    S& operator()(int i, int j) { return st.data()[(size_t)j * rows() + i]; }
    const S& operator()(int i, int j) const { return st.data()[(size_t)j * rows() + i]; }
    S& operator()(int i) { return st.data()[i]; }
    const S& operator()(int i) const { return st.data()[i]; }
    S& operator[](int i) { return st.data()[i]; }
    const S& operator[](int i) const { return st.data()[i]; }
    S& x() { return st.data()[0]; }
    S& y() { return st.data()[1]; }
    S& z() { return st.data()[2]; }
    S& w() { return st.data()[3]; }
    const S& x() const { return st.data()[0]; }
    const S& y() const { return st.data()[1]; }
    const S& z() const { return st.data()[2]; }
    const S& w() const { return st.data()[3]; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { return Matrix(r, c); }
    static Matrix Ones() {
        Matrix m;
        for (int i = 0; i < m.size(); i++) m[i] = S(1);
        return m;
    }
    static Matrix Identity() {
        Matrix m;
        for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = S(1);
        return m;
    }
    Matrix& setZero() {
        for (int i = 0; i < size(); i++) st.data()[i] = S();
        return *this;
    }
    Matrix& setConstant(S v) {
        for (int i = 0; i < size(); i++) st.data()[i] = v;
        return *this;
    }
    Matrix& setIdentity() {
        *this = Identity();
        return *this;
    }
    CommaInit<Matrix> operator<<(S v) {
        (*this)(0, 0) = v;
        return CommaInit<Matrix>{*this, 1};
    }
    CommaInit<Matrix> operator<<(const Matrix& o) {  // `a << b` with a whole matrix: plain assignment
        *this = o;
        return CommaInit<Matrix>{*this, size()};
    }
    template <typename S2, int R2, int C2, typename std::enable_if<(R2 != R || C2 != C), int>::type = 0>
    CommaInit<Matrix> operator<<(const Matrix<S2, R2, C2>& o) {  // first block of a row of blocks
        CommaInit<Matrix> ci{*this, 0};
        ci, o;
        return ci;
    }
    RowRef<Matrix> row(int i) { return RowRef<Matrix>{*this, i}; }
    Matrix<S, 1, C> row(int i) const {
        Matrix<S, 1, C> r;
        for (int j = 0; j < cols(); j++) r[j] = (*this)(i, j);
        return r;
    }
    Matrix<S, R, 1> col(int j) const {
        Matrix<S, R, 1> r;
        for (int i = 0; i < rows(); i++) r[i] = (*this)(i, j);
        return r;
    }
    HeadDyn<S> head(int n) const {
        HeadDyn<S> h;
        h.n = n;
        for (int i = 0; i < n && i < 4; i++) h.v[i] = (*this)[i];
        return h;
    }

    Matrix operator+(const Matrix& o) const {
        Matrix r = *this;
        for (int i = 0; i < size(); i++) r[i] = (*this)[i] + o[i];
        return r;
    }
    Matrix operator-(const Matrix& o) const {
        Matrix r = *this;
        for (int i = 0; i < size(); i++) r[i] = (*this)[i] - o[i];
        return r;
    }
    Matrix operator-() const {
        Matrix r = *this;
        for (int i = 0; i < size(); i++) r[i] = -(*this)[i];
        return r;
    }
    Matrix& operator+=(const Matrix& o) {
        for (int i = 0; i < size(); i++) (*this)[i] = (*this)[i] + o[i];
        return *this;
    }
    Matrix& operator-=(const Matrix& o) {
        for (int i = 0; i < size(); i++) (*this)[i] = (*this)[i] - o[i];
        return *this;
    }
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    Matrix operator*(T s) const {
        Matrix r = *this;
        const S f = (S)s;
        for (int i = 0; i < size(); i++) r[i] = (*this)[i] * f;
        return r;
    }
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    Matrix operator/(T s) const {
        Matrix r = *this;
        const S f = (S)s;
        for (int i = 0; i < size(); i++) r[i] = (*this)[i] / f;
        return r;
    }
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    Matrix& operator*=(T s) {
        const S f = (S)s;
        for (int i = 0; i < size(); i++) (*this)[i] = (*this)[i] * f;
        return *this;
    }
    template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    Matrix& operator/=(T s) {
        const S f = (S)s;
        for (int i = 0; i < size(); i++) (*this)[i] = (*this)[i] / f;
        return *this;
    }
    // matrix product: (a0*b0 + a1*b1) + a2*b2 ..., summed left to right
    template <int C2>
    Matrix<S, R, C2> operator*(const Matrix<S, C, C2>& o) const {
        Matrix<S, R, C2> r;
        if (R < 0 || C2 < 0) r.resize(rows(), o.cols());
        for (int i = 0; i < rows(); i++)
            for (int j = 0; j < o.cols(); j++) {
                S acc = (*this)(i, 0) * o(0, j);
                for (int k = 1; k < cols(); k++) acc = acc + (*this)(i, k) * o(k, j);
                r(i, j) = acc;
            }
        return r;
    }
    Matrix<S, C, R> transpose() const {
        Matrix<S, C, R> r;
        if (R < 0 || C < 0) r.resize(cols(), rows());
        for (int i = 0; i < rows(); i++)
            for (int j = 0; j < cols(); j++) r(j, i) = (*this)(i, j);
        return r;
    }
    S dot(const Matrix& o) const {
        S acc = (*this)[0] * o[0];
        for (int i = 1; i < size(); i++) acc = acc + (*this)[i] * o[i];
        return acc;
    }
    S squaredNorm() const { return dot(*this); }
    S norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { return *this / norm(); }
    void normalize() { *this = normalized(); }
    Matrix cross(const Matrix& o) const {
        Matrix r;
        r[0] = (*this)[1] * o[2] - (*this)[2] * o[1];
        r[1] = (*this)[2] * o[0] - (*this)[0] * o[2];
        r[2] = (*this)[0] * o[1] - (*this)[1] * o[0];
        return r;
    }
    template <typename T>
    Matrix<T, R, C> cast() const {
        Matrix<T, R, C> r;
        if (R < 0 || C < 0) r.resize(rows(), cols());
        for (int i = 0; i < size(); i++) r[i] = (T)(*this)[i];
        return r;
    }
    template <int BR, int BC>
    Matrix<S, BR, BC> block(int i0, int j0) const {
        Matrix<S, BR, BC> r;
        for (int i = 0; i < BR; i++)
            for (int j = 0; j < BC; j++) r(i, j) = (*this)(i0 + i, j0 + j);
        return r;
    }
    template <int N>
    Matrix<S, N, 1> head() const {
        Matrix<S, N, 1> r;
        for (int i = 0; i < N; i++) r[i] = (*this)[i];
        return r;
    }
    // 3x3 only (sensors/src/Pinhole.cpp:103: K.transpose().inverse(), K.inverse()).  Written as Eigen 3.3 / 3.4 evaluate
    // a fixed 3x3 inverse (Eigen/src/LU/InverseImpl.h, compute_inverse<.., 3>): cofactor(i, j) = m(i1, j1) * m(i2, j2) -
    // m(i1, j2) * m(i2, j1) with i1 = (i + 1) % 3, i2 = (i + 2) % 3; det = (cof(0,0) * m(0,0) + cof(1,0) * m(1,0)) +
    // cof(2,0) * m(2,0); every entry = cofactor(j, i) * (1 / det).  ASSUMPTION about Eigen's arithmetic, like the rest
    // of this file; the product path never inverts a matrix itself (the caller hands it F12).
    Matrix inverse() const {
        static_assert(R == 3 && C == 3, "stand-in: only the fixed 3x3 inverse is provided");
        const Matrix& m = *this;
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
        };
        const S c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        const S det = (c0 * m(0, 0) + c1 * m(1, 0)) + c2 * m(2, 0);
        const S invdet = S(1) / det;
        Matrix r;
        r(0, 0) = c0 * invdet;
        r(0, 1) = c1 * invdet;
        r(0, 2) = c2 * invdet;
        r(1, 0) = cof(0, 1) * invdet;
        r(1, 1) = cof(1, 1) * invdet;
        r(1, 2) = cof(2, 1) * invdet;
        r(2, 0) = cof(0, 2) * invdet;
        r(2, 1) = cof(1, 2) * invdet;
        r(2, 2) = cof(2, 2) * invdet;
        return r;
    }
    S determinant() const;
    S trace() const {
        S t = S();
        for (int i = 0; i < (R < C ? R : C); i++) t = t + (*this)(i, i);
        return t;
    }
    bool operator==(const Matrix& o) const {
        for (int i = 0; i < size(); i++)
            if (!((*this)[i] == o[i])) return false;
        return true;
    }
};

template <typename T, typename S, int R, int C, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
Matrix<S, R, C> operator*(T s, const Matrix<S, R, C>& m) {
    Matrix<S, R, C> r = m;
    const S f = (S)s;
    for (int i = 0; i < m.size(); i++) r[i] = f * m[i];
    return r;
}
template <typename S, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C>& m) {
    for (int i = 0; i < m.rows(); i++) {
        for (int j = 0; j < m.cols(); j++) os << m(i, j) << " ";
        os << "\n";
    }
    return os;
}

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;

// JacobiSVD: only what KannalaBrandt8::Triangulate uses (sensors/src/KannalaBrandt8.cpp:233-234) -- the right singular
// vector of the SMALLEST singular value of a 4 x 4 float matrix as matrixV().col(3).  NOT Eigen's two-sided Jacobi
// iteration: the vector comes from oracle/ppg_oracle.c::ppgo_null_vector4 (cyclic Jacobi on A^T A in double, rounded to
// float; checked against numpy.linalg.svd in tests/test_oracle_triangulation.py).  Eigen's own float iteration agrees with
// it to a few float ulps of the vector; everything downstream of this vector in the reference is compiled unmodified.
// The other columns of matrixV() are not computed (zero).
extern "C" void ppgo_null_vector4(const float* A_rowmajor, float* v4);
enum { ComputeFullU = 4, ComputeThinU = 8, ComputeFullV = 16, ComputeThinV = 32 };
template <typename M>
class JacobiSVD {
   public:
    JacobiSVD(const M& A, unsigned int = 0) {
        static_assert(M::RowsAtCompileTime == 4 && M::ColsAtCompileTime == 4, "stand-in: 4 x 4 only");
        float a[16], v[4];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) a[4 * i + j] = (float)A(i, j);
        ppgo_null_vector4(a, v);
        for (int i = 0; i < 4; i++) V(i, 3) = (typename M::Scalar)v[i];
    }
    const M& matrixV() const { return V; }

   private:
    M V;
};

template <typename S, int N>
class DiagonalMatrix {
   public:
    Matrix<S, N, 1> d;
    Matrix<S, N, 1>& diagonal() { return d; }
    const Matrix<S, N, 1>& diagonal() const { return d; }
    void setIdentity() { d = Matrix<S, N, 1>::Ones(); }
};

// Rotations: off the front-end path (poses, IMU); enough for the reference's headers to instantiate.
template <typename S>
class AngleAxis;
template <typename S>
class Quaternion {
   public:
    S w_, x_, y_, z_;
    Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
    Quaternion(S w, S x, S y, S z) : w_(w), x_(x), y_(y), z_(z) {}
    Quaternion(const Matrix<S, 3, 3>& R) {  // Shepperd
        const S tr = R(0, 0) + R(1, 1) + R(2, 2);
        if (tr > 0) {
            S s = std::sqrt(tr + S(1)) * 2;
            w_ = s / 4;
            x_ = (R(2, 1) - R(1, 2)) / s;
            y_ = (R(0, 2) - R(2, 0)) / s;
            z_ = (R(1, 0) - R(0, 1)) / s;
        } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
            S s = std::sqrt(S(1) + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
            w_ = (R(2, 1) - R(1, 2)) / s;
            x_ = s / 4;
            y_ = (R(0, 1) + R(1, 0)) / s;
            z_ = (R(0, 2) + R(2, 0)) / s;
        } else if (R(1, 1) > R(2, 2)) {
            S s = std::sqrt(S(1) + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
            w_ = (R(0, 2) - R(2, 0)) / s;
            x_ = (R(0, 1) + R(1, 0)) / s;
            y_ = s / 4;
            z_ = (R(1, 2) + R(2, 1)) / s;
        } else {
            S s = std::sqrt(S(1) + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
            w_ = (R(1, 0) - R(0, 1)) / s;
            x_ = (R(0, 2) + R(2, 0)) / s;
            y_ = (R(1, 2) + R(2, 1)) / s;
            z_ = s / 4;
        }
    }
    Quaternion(const AngleAxis<S>& aa);
    static Quaternion Identity() { return Quaternion(); }
    S w() const { return w_; }
    S x() const { return x_; }
    S y() const { return y_; }
    S z() const { return z_; }
    S norm() const { return std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_); }
    Quaternion normalized() const {
        const S n = norm();
        return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n);
    }
    void normalize() { *this = normalized(); }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const {
        const S n2 = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_;
        return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    }
    Quaternion operator*(const Quaternion& o) const {
        return Quaternion(w_ * o.w_ - x_ * o.x_ - y_ * o.y_ - z_ * o.z_, w_ * o.x_ + x_ * o.w_ + y_ * o.z_ - z_ * o.y_,
                          w_ * o.y_ - x_ * o.z_ + y_ * o.w_ + z_ * o.x_, w_ * o.z_ + x_ * o.y_ - y_ * o.x_ + z_ * o.w_);
    }
    Matrix<S, 3, 3> toRotationMatrix() const {
        Matrix<S, 3, 3> R;
        R(0, 0) = 1 - 2 * (y_ * y_ + z_ * z_);
        R(0, 1) = 2 * (x_ * y_ - z_ * w_);
        R(0, 2) = 2 * (x_ * z_ + y_ * w_);
        R(1, 0) = 2 * (x_ * y_ + z_ * w_);
        R(1, 1) = 1 - 2 * (x_ * x_ + z_ * z_);
        R(1, 2) = 2 * (y_ * z_ - x_ * w_);
        R(2, 0) = 2 * (x_ * z_ - y_ * w_);
        R(2, 1) = 2 * (y_ * z_ + x_ * w_);
        R(2, 2) = 1 - 2 * (x_ * x_ + y_ * y_);
        return R;
    }
    Matrix<S, 3, 1> operator*(const Matrix<S, 3, 1>& v) const { return toRotationMatrix() * v; }
    template <typename T>
    Quaternion<T> cast() const {
        return Quaternion<T>((T)w_, (T)x_, (T)y_, (T)z_);
    }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;

template <typename S>
class AngleAxis {
   public:
    S angle_;
    Matrix<S, 3, 1> axis_;
    AngleAxis() : angle_(0) { axis_[0] = 1; }
    AngleAxis(S a, const Matrix<S, 3, 1>& ax) : angle_(a), axis_(ax) {}
    AngleAxis(const Quaternion<S>& q) {
        const S n = std::sqrt(q.x_ * q.x_ + q.y_ * q.y_ + q.z_ * q.z_);
        if (n < S(1e-12)) {
            angle_ = 0;
            axis_[0] = 1;
        } else {
            angle_ = S(2) * std::atan2(n, q.w_);
            axis_[0] = q.x_ / n;
            axis_[1] = q.y_ / n;
            axis_[2] = q.z_ / n;
        }
    }
    S angle() const { return angle_; }
    const Matrix<S, 3, 1>& axis() const { return axis_; }
    Matrix<S, 3, 3> toRotationMatrix() const { return Quaternion<S>(*this).toRotationMatrix(); }
};
template <typename S>
Quaternion<S>::Quaternion(const AngleAxis<S>& aa) {
    const S h = aa.angle_ / 2;
    const S s = std::sin(h);
    w_ = std::cos(h);
    x_ = aa.axis_[0] * s;
    y_ = aa.axis_[1] * s;
    z_ = aa.axis_[2] * s;
}
typedef AngleAxis<float> AngleAxisf;
typedef AngleAxis<double> AngleAxisd;

// Map over a caller's buffer: only what DescriptorDistance (feature/src/MapPoint.cpp:22-29) does with it.
template <typename M, int Opt = 0>
class Map {
   public:
    typedef typename std::remove_const<M>::type Plain;
    typedef typename Plain::Scalar Scalar;
    const Scalar* p;
    int r, c;
    Map(const Scalar* ptr, int rows, int cols) : p(ptr), r(rows), c(cols) {}
    int size() const { return r * c; }
    struct Diff {
        const Scalar *a, *b;
        int n;
        // The fixed summation order the oracle (ppgo_descriptor_distance) and the CUDA kernels share: 32 strided partial
        // sums, then a xor butterfly (Eigen's own vectorised order is unspecified; only thresholds / orderings matter).
        Scalar norm() const {
            Scalar part[32];
            for (int l = 0; l < 32; l++) {
                Scalar s = 0;
                for (int k = l; k < n; k += 32) {
                    const Scalar d = a[k] - b[k];
                    s = s + d * d;
                }
                part[l] = s;
            }
            for (int o = 16; o >= 1; o >>= 1)
                for (int l = 0; l < 32; l++)
                    if ((l & o) == 0) {
                        const Scalar t = part[l] + part[l ^ o];
                        part[l] = t;
                        part[l ^ o] = t;
                    }
            return std::sqrt(part[0]);
        }
    };
    Diff operator-(const Map& o) const { return Diff{p, o.p, size()}; }
};

}  // namespace Eigen
