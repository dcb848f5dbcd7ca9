/* compat.h — minimal stand-ins for the third-party types that appear in the reference's public
 * interfaces (Eigen fixed-size vectors/matrices, cv::Point2f / cv::Mat, ros::Time).
 *
 * The reference headers (include/ekf_vio/TightlyCoupledEKF.h:18-23, Feature.h:11-35,
 * Frame.h:11-22, KLTTracker.h:11-23) include Eigen, OpenCV and ROS.  When those libraries are
 * installed the facade uses them (define EKFVIO_USE_REAL_DEPS, or let __has_include find them);
 * this image has none of them, so the facade compiles against the small PODs below, which keep
 * the member names and operators the reference's call sites use.
 */
#ifndef EKFVIO_COMPAT_H_
#define EKFVIO_COMPAT_H_

#if !defined(EKFVIO_NO_REAL_DEPS) && defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<opencv2/core/core.hpp>) && __has_include(<ros/ros.h>)
#define EKFVIO_USE_REAL_DEPS 1
#endif
#endif

#ifdef EKFVIO_USE_REAL_DEPS
#include <Eigen/Core>
#include <Eigen/Sparse>
#include <opencv2/core/core.hpp>
#include <ros/ros.h>
namespace ekfvio_compat { typedef Eigen::MatrixXf SigmaMatrix; }
#else

#include <cstddef>
#include <cstdint>
#include <vector>

namespace Eigen {

template <typename T, int R, int C>
class Matrix {
public:
    T v[R * C];
    Matrix() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    Matrix(T a, T b) { static_assert(R * C == 2, "two-coefficient constructor"); v[0] = a; v[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "three-coefficient constructor"); v[0] = a; v[1] = b; v[2] = c; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    void setZero() { for (int i = 0; i < R * C; ++i) v[i] = T(0); }
    int rows() const { return R; }
    int cols() const { return C; }
    int size() const { return R * C; }
    /* column-major storage, as Eigen's default: operator()(i) is the LINEAR index — the reference's
     * K(2)/K(5) reads (Feature.h:60-66) depend on exactly this */
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T& operator()(int r, int c) { return v[c * R + r]; }
    const T& operator()(int r, int c) const { return v[c * R + r]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& x() { return v[0]; }
    T& y() { return v[1]; }
    T& z() { return v[2]; }
    const T& x() const { return v[0]; }
    const T& y() const { return v[1]; }
    const T& z() const { return v[2]; }
    bool operator==(const Matrix& o) const { for (int i = 0; i < R * C; ++i) if (v[i] != o.v[i]) return false; return true; }
};
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;

/* dense dynamic float matrix: stands in for both MatrixXf and SparseMatrix<float> (Sigma, F, H, Q) */
class MatrixXf {
public:
    MatrixXf() : r_(0), c_(0) {}
    MatrixXf(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.f) {}
    void resize(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.f); }
    void conservativeResize(int r, int c) {
        std::vector<float> n((size_t)r * c, 0.f);
        for (int j = 0; j < (c < c_ ? c : c_); ++j) for (int i = 0; i < (r < r_ ? r : r_); ++i) n[(size_t)j * r + i] = d_[(size_t)j * r_ + i];
        d_.swap(n); r_ = r; c_ = c;
    }
    int rows() const { return r_; }
    int cols() const { return c_; }
    float& operator()(int i, int j) { return d_[(size_t)j * r_ + i]; }
    const float& operator()(int i, int j) const { return d_[(size_t)j * r_ + i]; }
    float coeff(int i, int j) const { return (*this)(i, j); }
    float& coeffRef(int i, int j) { return (*this)(i, j); }
    MatrixXf toDense() const { return *this; }
    long nonZeros() const { long n = 0; for (float x : d_) n += x != 0.f; return n; }
    Matrix2f block2(int i, int j) const { Matrix2f m; m(0, 0) = (*this)(i, j); m(0, 1) = (*this)(i, j + 1); m(1, 0) = (*this)(i + 1, j); m(1, 1) = (*this)(i + 1, j + 1); return m; }
    bool operator==(const MatrixXf& o) const { return r_ == o.r_ && c_ == o.c_ && d_ == o.d_; }
    const float* data() const { return d_.data(); }
private:
    int r_, c_;
    std::vector<float> d_;   /* column-major */
};
template <typename T> using SparseMatrix = MatrixXf;   /* Sigma is ~90 % dense after one step */

}  // namespace Eigen

namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
/* 8-bit single-channel image view (what Frame::img holds after cvtColor, klt_test.cpp:24-26) */
struct Mat {
    int rows, cols;
    size_t step;
    const uint8_t* data;
    std::vector<uint8_t> owned;
    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, const uint8_t* p, size_t s) : rows(r), cols(c), step(s), data(p) {}
    static Mat copyOf(int r, int c, const uint8_t* p, size_t s) {
        Mat m; m.rows = r; m.cols = c; m.step = (size_t)c; m.owned.resize((size_t)r * c);
        for (int y = 0; y < r; ++y) for (int x = 0; x < c; ++x) m.owned[(size_t)y * c + x] = p[(size_t)y * s + x];
        m.data = m.owned.data(); return m;
    }
};
}  // namespace cv

namespace ros {
struct Time {
    double sec;
    Time() : sec(0) {}
    explicit Time(double s) : sec(s) {}
    double toSec() const { return sec; }
    bool operator==(const Time& o) const { return sec == o.sec; }
};
}  // namespace ros

namespace ekfvio_compat { typedef Eigen::MatrixXf SigmaMatrix; }
#endif /* EKFVIO_USE_REAL_DEPS */

#endif /* EKFVIO_COMPAT_H_ */
