// TEST INFRASTRUCTURE ONLY — a stand-in for the slice of Eigen 3 that the reference's EKF sources use, so that the
// UNMODIFIED /root/reference/include/ekf_vio/TightlyCoupledEKF.cpp + Feature.cpp compile in this image (Eigen itself is not
// installed and there is no network).  oracle/Makefile target `ref` builds oracle/_ref/libekf_ref_{f32,f64}.so from them.
//
// Every operation is evaluated eagerly (no expression templates) with Eigen's documented semantics:
//   * dense Matrix<T,R,C>: column-major, linear operator()(i) walks the column-major storage (Feature.h:60-66 relies on it, E1);
//     scalar arguments of operator* / operator/ / compound assignments are narrowed to T first (Eigen's promote_scalar_arg);
//   * Quaternion<T>: (w,x,y,z) constructor, Hamilton product, q*v = v + w*uv + qv x uv with uv = 2 (qv x v) (no normalisation),
//     inverse() = conjugate / squaredNorm (zero if the norm is zero), normalize() divides by norm();
//   * SparseMatrix<T>: stored densely here — structural zeros and explicit zeros carry the same value, and the reference never
//     reads the pattern except in debug prints; prune(ref, eps) / sparseView(ref, eps) drop |v| <= |ref|*eps;
//   * SimplicialLDLT: LDL^T WITHOUT numerical pivoting of the LOWER triangle of its argument, up-looking (row by row) as
//     Eigen's simplicial factorisation does; D_kk == 0 -> info() == NumericalIssue.  Eigen's AMD fill-reducing permutation is
//     NOT reproduced (natural ordering): on the dense S of this filter it only changes rounding.
// The scalar behind the *f typedefs (Vector3f, Quaternionf, ...) is EKFVIO_SHIM_SCALAR (default float); the f64 build sets
// it to double and compiles the reference sources with `float` mapped to `double` (oracle/ref_build/ref_capi.cpp).
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <ostream>
#include <type_traits>
#include <vector>

#ifndef EKFVIO_SHIM_SCALAR
#define EKFVIO_SHIM_SCALAR float
#endif

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { Dynamic = -1 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

template <class T> class SparseMatrix;

namespace shim {
template <class T, int N> struct Store {                 // fixed size
    std::array<T, (N > 0 ? N : 1)> v{};
    void resize(std::size_t) {}
    std::size_t size() const { return (std::size_t)N; }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
};
template <class T> struct Store<T, -1> {                 // dynamic
    std::vector<T> v;
    void resize(std::size_t n) { v.assign(n, T(0)); }
    std::size_t size() const { return v.size(); }
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
};
template <class S> using if_scalar = typename std::enable_if<std::is_arithmetic<S>::value, int>::type;
template <class S> using if_integral = typename std::enable_if<std::is_integral<S>::value, int>::type;
}  // namespace shim

template <class T, int R, int C>
class Matrix {
    static constexpr int kN = (R > 0 && C > 0) ? R * C : -1;
    shim::Store<T, kN> s_;
    Index rows_ = (R > 0 ? R : 0), cols_ = (C > 0 ? C : 0);

public:
    typedef T Scalar;
    Matrix() {}
    // VectorXf(n): dynamic vector of n coefficients (zero-filled here; Eigen leaves them uninitialised)
    template <class I, shim::if_integral<I> = 0>
    explicit Matrix(I n) {
        static_assert(R == Dynamic || C == Dynamic, "size constructor on a fixed-size type");
        if (C == 1) resize((Index)n, 1); else resize(1, (Index)n);
    }
    // two arguments: (rows, cols) of a dynamic matrix, or the two coefficients of a 2-vector
    template <class A, class B, shim::if_scalar<A> = 0, shim::if_scalar<B> = 0>
    Matrix(A a, B b) {
        if constexpr (R == Dynamic && C == Dynamic) resize((Index)a, (Index)b);
        else { static_assert(R * C == 2, "two-coefficient constructor"); s_.v[0] = T(a); s_.v[1] = T(b); }
    }
    template <class A, class B, class D, shim::if_scalar<A> = 0>
    Matrix(A x, B y, D z) { static_assert(R * C == 3, "three-coefficient constructor"); s_.v[0] = T(x); s_.v[1] = T(y); s_.v[2] = T(z); }
    // conversion between shapes of the same scalar (e.g. a dynamic block into a Matrix2f)
    template <int R2, int C2>
    Matrix(const Matrix<T, R2, C2>& o) { resize(o.rows(), o.cols()); for (Index i = 0; i < size(); ++i) s_.data()[i] = o.data()[i]; }

    void resize(Index r, Index c) {
        if (R > 0 && C > 0) { if (r != R || c != C) std::abort(); return; }
        rows_ = r; cols_ = c; s_.resize((std::size_t)(r * c));
    }
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index size() const { return rows_ * cols_; }
    T* data() { return s_.data(); }
    const T* data() const { return s_.data(); }

    T& operator()(Index i) { return s_.data()[i]; }
    const T& operator()(Index i) const { return s_.data()[i]; }
    T& operator()(Index i, Index j) { return s_.data()[i + j * rows_]; }
    const T& operator()(Index i, Index j) const { return s_.data()[i + j * rows_]; }
    T& operator[](Index i) { return s_.data()[i]; }
    const T& operator[](Index i) const { return s_.data()[i]; }
    T x() const { return s_.data()[0]; }
    T y() const { return s_.data()[1]; }
    T z() const { return s_.data()[2]; }

    Matrix& setZero() { for (Index i = 0; i < size(); ++i) s_.data()[i] = T(0); return *this; }
    Matrix& noalias() { return *this; }

    Matrix& operator+=(const Matrix& o) { for (Index i = 0; i < size(); ++i) s_.data()[i] += o.s_.data()[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (Index i = 0; i < size(); ++i) s_.data()[i] -= o.s_.data()[i]; return *this; }
    template <class S, shim::if_scalar<S> = 0> Matrix& operator/=(S sc) { const T t = T(sc); for (Index i = 0; i < size(); ++i) s_.data()[i] /= t; return *this; }
    template <class S, shim::if_scalar<S> = 0> Matrix& operator*=(S sc) { const T t = T(sc); for (Index i = 0; i < size(); ++i) s_.data()[i] *= t; return *this; }
    Matrix operator+(const Matrix& o) const { Matrix r(*this); r += o; return r; }
    Matrix operator-(const Matrix& o) const { Matrix r(*this); r -= o; return r; }
    Matrix operator-() const { Matrix r(*this); for (Index i = 0; i < size(); ++i) r.s_.data()[i] = -s_.data()[i]; return r; }
    template <class S, shim::if_scalar<S> = 0> Matrix operator/(S sc) const { Matrix r(*this); r /= sc; return r; }
    template <class S, shim::if_scalar<S> = 0> Matrix operator*(S sc) const { Matrix r(*this); r *= sc; return r; }

    T squaredNorm() const { T a = T(0); for (Index i = 0; i < size(); ++i) a += s_.data()[i] * s_.data()[i]; return a; }
    T norm() const { return std::sqrt(squaredNorm()); }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t; t.resize(cols_, rows_);
        for (Index i = 0; i < rows_; ++i) for (Index j = 0; j < cols_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }

    // v.segment(start, n) = / -= another vector
    struct Segment {
        Matrix& m; Index start, n;
        template <int R2, int C2> Segment& operator=(const Matrix<T, R2, C2>& o) { for (Index i = 0; i < n; ++i) m(start + i) = o(i); return *this; }
        template <int R2, int C2> Segment& operator-=(const Matrix<T, R2, C2>& o) { for (Index i = 0; i < n; ++i) m(start + i) -= o(i); return *this; }
        template <int R2, int C2> Segment& operator+=(const Matrix<T, R2, C2>& o) { for (Index i = 0; i < n; ++i) m(start + i) += o(i); return *this; }
    };
    Segment segment(Index start, Index n) { return Segment{*this, start, n}; }

    SparseMatrix<T> sparseView(const T& reference = T(0), const T& epsilon = T(1e-5)) const;   // defined below
};

template <class S, class T, int R, int C, shim::if_scalar<S> = 0>
Matrix<T, R, C> operator*(S sc, const Matrix<T, R, C>& m) { Matrix<T, R, C> r(m); r *= sc; return r; }

template <class T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
    for (Index i = 0; i < m.rows(); ++i) { for (Index j = 0; j < m.cols(); ++j) os << m(i, j) << (j + 1 < m.cols() ? " " : ""); if (i + 1 < m.rows()) os << "\n"; }
    return os;
}

typedef Matrix<EKFVIO_SHIM_SCALAR, 2, 1> Vector2f;
typedef Matrix<EKFVIO_SHIM_SCALAR, 3, 1> Vector3f;
typedef Matrix<EKFVIO_SHIM_SCALAR, 2, 2> Matrix2f;
typedef Matrix<EKFVIO_SHIM_SCALAR, 3, 3> Matrix3f;
typedef Matrix<EKFVIO_SHIM_SCALAR, Dynamic, 1> VectorXf;
typedef Matrix<EKFVIO_SHIM_SCALAR, Dynamic, Dynamic> MatrixXf;

// ---------------------------------------------------------------------------------------------------------------------
template <class T>
class Quaternion {
    T w_, x_, y_, z_;

public:
    Quaternion() : w_(0), x_(0), y_(0), z_(0) {}
    template <class A, class B, class C_, class D>
    Quaternion(A w, B x, C_ y, D z) : w_(T(w)), x_(T(x)), y_(T(y)), z_(T(z)) {}
    static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
    T w() const { return w_; }
    T x() const { return x_; }
    T y() const { return y_; }
    T z() const { return z_; }
    T squaredNorm() const { return x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; }
    T norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const T n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const {
        const T n2 = squaredNorm();
        if (n2 > T(0)) return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
        return Quaternion(T(0), T(0), T(0), T(0));
    }
    // Hamilton product (Eigen's quat_product, generic path)
    Quaternion operator*(const Quaternion& b) const {
        const Quaternion& a = *this;
        return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_,
                          a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                          a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
                          a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    // QuaternionBase::_transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv — assumes a unit quaternion, normalises nothing
    Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {
        T ux = y_ * v(2) - z_ * v(1), uy = z_ * v(0) - x_ * v(2), uz = x_ * v(1) - y_ * v(0);
        ux += ux; uy += uy; uz += uz;
        return Matrix<T, 3, 1>(v(0) + w_ * ux + (y_ * uz - z_ * uy), v(1) + w_ * uy + (z_ * ux - x_ * uz), v(2) + w_ * uz + (x_ * uy - y_ * ux));
    }
};
typedef Quaternion<EKFVIO_SHIM_SCALAR> Quaternionf;

// ---------------------------------------------------------------------------------------------------------------------
template <class T>
class SparseMatrix {
    Index rows_ = 0, cols_ = 0;
    std::vector<T> v_;                                    // column-major, dense backing (see the header comment)

    static bool much_smaller(T value, T reference, T epsilon) { return std::abs(value) <= std::abs(reference) * epsilon; }

public:
    typedef T Scalar;
    SparseMatrix() {}
    SparseMatrix(Index r, Index c) { resize(r, c); }
    void resize(Index r, Index c) { rows_ = r; cols_ = c; v_.assign((std::size_t)(r * c), T(0)); }
    void conservativeResize(Index r, Index c) {
        std::vector<T> nv((std::size_t)(r * c), T(0));
        for (Index j = 0; j < (c < cols_ ? c : cols_); ++j) for (Index i = 0; i < (r < rows_ ? r : rows_); ++i) nv[(std::size_t)(i + j * r)] = v_[(std::size_t)(i + j * rows_)];
        v_.swap(nv); rows_ = r; cols_ = c;
    }
    template <class I> void reserve(I) {}
    Index rows() const { return rows_; }
    Index cols() const { return cols_; }
    Index nonZeros() const { Index c = 0; for (T x : v_) c += (x != T(0)); return c; }
    T& insert(Index i, Index j) { return v_[(std::size_t)(i + j * rows_)]; }
    T& coeffRef(Index i, Index j) { return v_[(std::size_t)(i + j * rows_)]; }
    T coeff(Index i, Index j) const { return v_[(std::size_t)(i + j * rows_)]; }
    const T* data() const { return v_.data(); }
    T* data() { return v_.data(); }
    void setIdentity() { for (auto& x : v_) x = T(0); for (Index i = 0; i < (rows_ < cols_ ? rows_ : cols_); ++i) v_[(std::size_t)(i + i * rows_)] = T(1); }
    void finalize() {}

    // keep a coefficient iff !(|v| <= |reference| * epsilon)   (SparseMatrix::prune / internal::isMuchSmallerThan)
    void prune(const T& reference, const T& epsilon = T(1e-5)) { for (auto& x : v_) if (much_smaller(x, reference, epsilon)) x = T(0); }

    SparseMatrix transpose() const {
        SparseMatrix t(cols_, rows_);
        for (Index j = 0; j < cols_; ++j) for (Index i = 0; i < rows_; ++i) t.v_[(std::size_t)(j + i * cols_)] = v_[(std::size_t)(i + j * rows_)];
        return t;
    }
    Matrix<T, Dynamic, Dynamic> toDense() const {
        Matrix<T, Dynamic, Dynamic> d; d.resize(rows_, cols_);
        for (std::size_t k = 0; k < v_.size(); ++k) d.data()[k] = v_[k];
        return d;
    }
    Matrix<T, Dynamic, Dynamic> block(Index r0, Index c0, Index nr, Index nc) const {
        Matrix<T, Dynamic, Dynamic> d; d.resize(nr, nc);
        for (Index j = 0; j < nc; ++j) for (Index i = 0; i < nr; ++i) d(i, j) = coeff(r0 + i, c0 + j);
        return d;
    }

    SparseMatrix& operator+=(const SparseMatrix& o) { for (std::size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k]; return *this; }
    SparseMatrix& operator-=(const SparseMatrix& o) { for (std::size_t k = 0; k < v_.size(); ++k) v_[k] -= o.v_[k]; return *this; }

    // sparse * sparse, column by column, each column accumulated over ascending inner index with zero coefficients of the
    // right factor skipped (the order of Eigen's conservative sparse product for sorted operands)
    SparseMatrix operator*(const SparseMatrix& b) const {
        if (cols_ != b.rows_) std::abort();
        SparseMatrix r(rows_, b.cols_);
        for (Index j = 0; j < b.cols_; ++j) {
            T* rc = &r.v_[(std::size_t)(j * rows_)];
            for (Index k = 0; k < cols_; ++k) {
                const T y = b.v_[(std::size_t)(k + j * b.rows_)];
                if (y == T(0)) continue;
                const T* ac = &v_[(std::size_t)(k * rows_)];
                for (Index i = 0; i < rows_; ++i) rc[i] += ac[i] * y;
            }
        }
        return r;
    }
    // sparse * dense vector
    Matrix<T, Dynamic, 1> operator*(const Matrix<T, Dynamic, 1>& x) const {
        if (cols_ != x.rows()) std::abort();
        Matrix<T, Dynamic, 1> r(rows_);
        for (Index k = 0; k < cols_; ++k) {
            const T y = x(k);
            const T* ac = &v_[(std::size_t)(k * rows_)];
            for (Index i = 0; i < rows_; ++i) if (ac[i] != T(0)) r(i) += ac[i] * y;
        }
        return r;
    }
};

template <class T, int R, int C>
SparseMatrix<T> Matrix<T, R, C>::sparseView(const T& reference, const T& epsilon) const {
    SparseMatrix<T> s(rows_, cols_);
    for (Index j = 0; j < cols_; ++j) for (Index i = 0; i < rows_; ++i) {
        const T x = (*this)(i, j);
        if (!(std::abs(x) <= std::abs(reference) * epsilon)) s.insert(i, j) = x;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------
template <class MatrixType>
class SimplicialLDLT {
    typedef typename MatrixType::Scalar T;
    Index n_ = 0;
    std::vector<T> L_, D_;                                 // L unit lower, column-major n x n; D diagonal
    ComputationInfo info_ = Success;

public:
    SimplicialLDLT() {}
    // reads only the LOWER triangle of a (UpLo = Lower); up-looking: row k of L from the rows above it
    void compute(const MatrixType& a) {
        n_ = a.rows(); info_ = Success;
        L_.assign((std::size_t)(n_ * n_), T(0)); D_.assign((std::size_t)n_, T(0));
        std::vector<T> y((std::size_t)n_);
        for (Index k = 0; k < n_; ++k) {
            for (Index i = 0; i < k; ++i) y[(std::size_t)i] = a.coeff(k, i);
            T d = a.coeff(k, k);
            for (Index i = 0; i < k; ++i) {
                const T yi = y[(std::size_t)i];
                y[(std::size_t)i] = T(0);
                for (Index j = i + 1; j < k; ++j) y[(std::size_t)j] -= L_[(std::size_t)(j + i * n_)] * yi;
                const T lki = yi / D_[(std::size_t)i];
                d -= lki * yi;
                L_[(std::size_t)(k + i * n_)] = lki;
            }
            D_[(std::size_t)k] = d;
            L_[(std::size_t)(k + k * n_)] = T(1);
            if (d == T(0)) { info_ = NumericalIssue; return; }
        }
    }
    ComputationInfo info() const { return info_; }
    // A^{-1} B for a dense right-hand side: forward substitution, D^{-1} as a product with the reciprocal, backward substitution
    Matrix<T, Dynamic, Dynamic> solve(const Matrix<T, Dynamic, Dynamic>& b) const {
        Matrix<T, Dynamic, Dynamic> x(b);
        if (b.rows() != n_) std::abort();
        for (Index c = 0; c < x.cols(); ++c) {
            T* xc = &x(0, c);
            for (Index i = 0; i < n_; ++i) {
                const T xi = xc[i];
                if (xi == T(0)) continue;
                for (Index j = i + 1; j < n_; ++j) xc[j] -= L_[(std::size_t)(j + i * n_)] * xi;
            }
            for (Index i = 0; i < n_; ++i) xc[i] *= T(1) / D_[(std::size_t)i];
            for (Index i = n_ - 1; i >= 0; --i) {
                T acc = xc[i];
                for (Index j = i + 1; j < n_; ++j) acc -= L_[(std::size_t)(j + i * n_)] * xc[j];
                xc[i] = acc;
            }
        }
        return x;
    }
};

}  // namespace Eigen
