// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of k-sheridan/ekf_vio's TightlyCoupledEKF (reference:
// include/ekf_vio/TightlyCoupledEKF.cpp, Feature.cpp) as a dense, Scalar-templated C++ class.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use it.  The reference itself cannot be compiled in this image (Eigen, ROS and the OpenCV C++
// headers are absent), so this file *is* the executable definition of the reference algorithm:
//   * Scalar = double : the parity oracle for the 1e-9 gate (north_star asks FP64).
//   * Scalar = float  : a "literal" instantiation in the reference's own precision, used only as
//                       a loose cross-check.
// Parity pinning: the reference's tests hold exactly two known answers for this path
// (test/test_ekf.cpp:27-37 resize keeps the top-left block; :51-63 the H matrix for {T,F,T});
// both are checked in tests/test_oracle_ekf.py.  Every other numeric result of the reference is
// printed, never recorded => "parity unpinned" beyond those two; see DESIGN.md.
//
// Eigen semantics that are not visible in the reference sources are written out here
// (SURVEY.md App. A): Quaternion*Vector3 = _transformVector (no normalisation), Hamilton product,
// inverse = conjugate / squaredNorm, prune/sparseView keep |v| > 1e-8*1e-5, SimplicialLDLT =
// LDL^T without pivoting on the upper triangle of S.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ekf_oracle {

constexpr int BASE = 22;                 // TightlyCoupledEKF.h:12  BASE_STATE_SIZE
constexpr double SPARSE_THRESH = 1e-8;   // TightlyCoupledEKF.h:13
constexpr double SPARSE_EPS = 1e-5;      // TightlyCoupledEKF.h:14
constexpr double DELTA_SHIFT = 1e-3;     // TightlyCoupledEKF.cpp:182
constexpr int QZ_INDEX = 6;              // TightlyCoupledEKF.cpp:180
constexpr int AZ_INDEX = 15;             // TightlyCoupledEKF.cpp:181

struct Params {                          // Params.h:83-86 defaults
    double default_point_depth = 0.5;
    double default_point_depth_variance = 100;
    double default_point_homogenous_variance = 0.00001;
};

template <class S> struct Quat { S w, x, y, z; };
template <class S> struct Vec3 { S x, y, z; };

template <class S> inline Vec3<S> cross(const Vec3<S>& a, const Vec3<S>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen::QuaternionBase::_transformVector — assumes unit q, does NOT normalise.
template <class S> inline Vec3<S> rotate(const Quat<S>& q, const Vec3<S>& v) {
    Vec3<S> qv{q.x, q.y, q.z};
    Vec3<S> uv = cross(qv, v);
    uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
    Vec3<S> c = cross(qv, uv);
    return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
// Hamilton product a*b (Eigen quaternion operator*).
template <class S> inline Quat<S> qmul(const Quat<S>& a, const Quat<S>& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
// Eigen inverse(): conjugate / squaredNorm, zero quaternion if the norm is zero.
template <class S> inline Quat<S> qinv(const Quat<S>& q) {
    S n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    if (n2 > S(0)) return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
    return {S(0), S(0), S(0), S(0)};
}
template <class S> inline Quat<S> qnormalized(const Quat<S>& q) {
    S n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}

template <class S>
class Filter {
public:
    using Base = std::array<S, BASE>;
    using F3 = std::array<S, 3>;

    Params prm;
    Base base_mu;
    std::vector<F3> feat;                       // Feature::mu = [u, v, 1/depth]   (Feature.h:41)
    std::vector<std::array<S, 2>> klt_last;     // Feature::last_result_from_klt_tracker
    std::vector<uint8_t> delete_flag;           // Feature::delete_flag
    std::vector<S> Sigma;                       // dense N x N, row-major
    int status = 0;                             // bit0: LDLT hit a zero pivot (TightlyCoupledEKF.cpp:579); bit3: a negative pivot was seen (diagnostic)
    // E2: the function-static cache of convolveFeature (TightlyCoupledEKF.cpp:400-403), made one
    // private cache per filter with the reference's initial values.
    S last_omega[3] = {S(0), S(0), S(0)};
    Quat<S> cache_dq_inv{S(1), S(0), S(0), S(0)};

    Filter() { reset(); }
    int dim() const { return BASE + 3 * (int)feat.size(); }
    S& P(int i, int j) { return Sigma[(size_t)i * dim() + j]; }

    // TightlyCoupledEKF.cpp:10-56 (constructor + initializeBaseState)
    void reset() {
        feat.clear(); klt_last.clear(); delete_flag.clear();
        base_mu.fill(S(0));
        base_mu[3] = S(1.0);
        Sigma.assign((size_t)BASE * BASE, S(0));
        for (int i = 7; i <= 15; ++i) Sigma[(size_t)i * BASE + i] = S(30);
        for (int i = 16; i <= 21; ++i) Sigma[(size_t)i * BASE + i] = S(0.5);
        status = 0;
    }

    // TightlyCoupledEKF.cpp:58-94 + Feature.cpp:14-20
    void addNewFeatures(const S* uv, int k) {
        if (!k) return;
        int old_n = dim(), new_n = old_n + 3 * k;
        std::vector<S> ns((size_t)new_n * new_n, S(0));       // conservativeResize keeps old block
        for (int i = 0; i < old_n; ++i)
            std::memcpy(&ns[(size_t)i * new_n], &Sigma[(size_t)i * old_n], sizeof(S) * old_n);
        S depth = S(prm.default_point_depth);                  // float average_scene_depth
        int idx = old_n;
        for (int f = 0; f < k; ++f) {
            F3 mu{uv[2 * f], uv[2 * f + 1], S(1.0 / (double)depth)};   // mu(2) = 1.0/depth
            feat.push_back(mu);
            klt_last.push_back({uv[2 * f], uv[2 * f + 1]});
            delete_flag.push_back(0);
            ns[(size_t)idx * new_n + idx] = S(prm.default_point_homogenous_variance); ++idx;
            ns[(size_t)idx * new_n + idx] = S(prm.default_point_homogenous_variance); ++idx;
            ns[(size_t)idx * new_n + idx] = S(prm.default_point_depth_variance); ++idx;
        }
        Sigma.swap(ns);
    }

    // TightlyCoupledEKF.cpp:328-395
    Base convolveBaseState(const Base& last, S dt) const {
        Vec3<S> pos{last[0], last[1], last[2]};
        Quat<S> quat{last[3], last[4], last[5], last[6]};
        Vec3<S> vel{last[7], last[8], last[9]};
        Vec3<S> omega{last[10], last[11], last[12]};
        Vec3<S> accel{last[13], last[14], last[15]};

        S hdt2 = S(0.5 * (double)dt * (double)dt);             // 0.5*dt*dt is a double expression
        Vec3<S> tr{dt * vel.x + hdt2 * accel.x, dt * vel.y + hdt2 * accel.y, dt * vel.z + hdt2 * accel.z};
        Vec3<S> r = rotate(quat, tr);
        pos = {pos.x + r.x, pos.y + r.y, pos.z + r.z};

        S omega_norm = std::sqrt(omega.x * omega.x + omega.y * omega.y + omega.z * omega.z);
        Quat<S> dq;
        if (omega_norm < S(1e-10)) {
            dq = qnormalized(Quat<S>{S(1.0), omega.x * dt, omega.y * dt, omega.z * dt});
        } else {
            S theta = dt * omega_norm;
            Vec3<S> oh{omega.x / omega_norm, omega.y / omega_norm, omega.z / omega_norm};
            S st2 = std::sin(theta / 2);
            dq = {std::cos(theta / 2), oh.x * st2, oh.y * st2, oh.z * st2};
        }
        Quat<S> dq_inv = qinv(dq);
        Vec3<S> va{vel.x + dt * accel.x, vel.y + dt * accel.y, vel.z + dt * accel.z};
        vel = rotate(dq_inv, va);
        accel = rotate(dq_inv, accel);
        quat = qmul(quat, dq);

        Base o = last;                                         // omega and biases unchanged
        o[0] = pos.x; o[1] = pos.y; o[2] = pos.z;
        o[3] = quat.w; o[4] = quat.x; o[5] = quat.y; o[6] = quat.z;
        o[7] = vel.x; o[8] = vel.y; o[9] = vel.z;
        o[13] = accel.x; o[14] = accel.y; o[15] = accel.z;
        return o;
    }

    // TightlyCoupledEKF.cpp:397-460 — including the omega-keyed cache (E2).
    F3 convolveFeature(const Base& bs, const F3& fs, S dt) {
        Vec3<S> vel{bs[7], bs[8], bs[9]};
        Vec3<S> accel{bs[13], bs[14], bs[15]};
        Vec3<S> fp;
        fp.z = S(1.0 / (double)fs[2]);                         // 1.0/feature_pos(2): double division
        fp.x = fs[0] * fp.z;
        fp.y = fs[1] * fp.z;
        S hdt2 = S(0.5 * (double)dt * (double)dt);
        Vec3<S> tr{dt * vel.x + hdt2 * accel.x, dt * vel.y + hdt2 * accel.y, dt * vel.z + hdt2 * accel.z};
        if (last_omega[0] != bs[10] || last_omega[1] != bs[11] || last_omega[2] != bs[12]) {
            Vec3<S> omega{bs[10], bs[11], bs[12]};
            S omega_norm = std::sqrt(omega.x * omega.x + omega.y * omega.y + omega.z * omega.z);
            if (omega_norm < S(1e-10)) {
                cache_dq_inv = qnormalized(Quat<S>{S(1.0), -omega.x * dt, -omega.y * dt, -omega.z * dt});
            } else {
                S theta = dt * omega_norm;
                Vec3<S> oh{omega.x / omega_norm, omega.y / omega_norm, omega.z / omega_norm};
                S st2 = std::sin(theta / 2);
                cache_dq_inv = {std::cos(theta / 2), -oh.x * st2, -oh.y * st2, -oh.z * st2};
            }
            last_omega[0] = bs[10]; last_omega[1] = bs[11]; last_omega[2] = bs[12];
        }
        Vec3<S> a = rotate(cache_dq_inv, fp);
        Vec3<S> b = rotate(cache_dq_inv, tr);
        fp = {a.x - b.x, a.y - b.y, a.z - b.z};
        fp.x /= fp.z;
        fp.y /= fp.z;
        fp.z = S(1.0 / (double)fp.z);
        return {fp.x, fp.y, fp.z};
    }

    // TightlyCoupledEKF.cpp:176-325 — dense N x N row-major F.
    std::vector<S> numericallyLinearizeProcess(S dt) {
        const int N = dim(), n = (int)feat.size();
        std::vector<S> F((size_t)N * N, S(0));
        // x += DELTA_SHIFT / x -= 2*DELTA_SHIFT are double expressions narrowed to Scalar; keep the
        // exact (x+d) then (-2d) sequence of TightlyCoupledEKF.cpp:193-198.
        const S two_d = S(2 * DELTA_SHIFT);
        auto up = [](S x) { return S((double)x + DELTA_SHIFT); };
        auto dn2 = [](S x) { return S((double)x - 2 * DELTA_SHIFT); };
        Base test_mu = base_mu;
        for (int j = 0; j < BASE; ++j) {
            if (j <= AZ_INDEX) {
                test_mu[j] = up(test_mu[j]);
                Base hi = convolveBaseState(test_mu, dt);
                test_mu[j] = dn2(test_mu[j]);
                Base lo = convolveBaseState(test_mu, dt);
                test_mu[j] = base_mu[j];
                for (int i = 0; i < BASE; ++i) F[(size_t)i * N + j] = (hi[i] - lo[i]) / two_d;
                if (j > QZ_INDEX) {
                    std::vector<S> fd((size_t)3 * n);
                    test_mu[j] = up(test_mu[j]);
                    for (int f = 0; f < n; ++f) {
                        F3 v = convolveFeature(test_mu, feat[f], dt);
                        fd[3 * f] = v[0]; fd[3 * f + 1] = v[1]; fd[3 * f + 2] = v[2];
                    }
                    test_mu[j] = dn2(test_mu[j]);
                    for (int f = 0; f < n; ++f) {
                        F3 v = convolveFeature(test_mu, feat[f], dt);
                        fd[3 * f] -= v[0]; fd[3 * f + 1] -= v[1]; fd[3 * f + 2] -= v[2];
                    }
                    test_mu[j] = base_mu[j];
                    for (int i = 0; i < 3 * n; ++i) F[(size_t)(BASE + i) * N + j] = fd[i] / two_d;
                }
            } else {
                F[(size_t)j * N + j] = S(1);
            }
        }
        int col = BASE;
        for (int f = 0; f < n; ++f) {
            F3 t = feat[f];
            int row = col;
            for (int c = 0; c < 3; ++c) {
                t[c] = up(t[c]);
                F3 hi = convolveFeature(base_mu, t, dt);
                t[c] = dn2(t[c]);
                F3 lo = convolveFeature(base_mu, t, dt);
                t[c] = feat[f][c];
                for (int r = 0; r < 3; ++r) F[(size_t)(row + r) * N + col] = (hi[r] - lo[r]) / two_d;
                ++col;
            }
        }
        return F;
    }

    // TightlyCoupledEKF.cpp:123-174 — diagonal of Q (float x = 0.0001*dt: double product narrowed).
    std::vector<S> generateProcessNoise(S dt) const {
        const int N = dim();
        std::vector<S> q(N);
        S low = S(0.0001 * (double)dt), pos = S(0.0001 * (double)dt), vel = S(0.01 * (double)dt);
        S om = S(5 * dt), acc = S(5 * dt), bias = S(0.001 * (double)dt);
        for (int i = 0; i <= 6; ++i) q[i] = pos;
        for (int i = 7; i <= 9; ++i) q[i] = vel;
        for (int i = 10; i <= 12; ++i) q[i] = om;
        for (int i = 13; i <= 15; ++i) q[i] = acc;
        for (int i = 16; i <= 21; ++i) q[i] = bias;
        for (int i = BASE; i < N; ++i) q[i] = low;
        return q;
    }

    static void prune(std::vector<S>& M) {       // SparseMatrix::prune(1e-8, 1e-5): keep |v| > 1e-13
        const S lim = S(SPARSE_THRESH) * S(SPARSE_EPS);
        for (auto& v : M) if (!(std::abs(v) > lim)) v = S(0);
    }

    // C(NxN) = A * B skipping zero entries of A (what a sparse product sums over).
    static std::vector<S> mul_skip(const std::vector<S>& A, const std::vector<S>& B, int N) {
        std::vector<S> C((size_t)N * N, S(0));
        for (int i = 0; i < N; ++i) {
            S* c = &C[(size_t)i * N];
            for (int k = 0; k < N; ++k) {
                S a = A[(size_t)i * N + k];
                if (a == S(0)) continue;
                const S* b = &B[(size_t)k * N];
                for (int j = 0; j < N; ++j) c[j] += a * b[j];
            }
        }
        return C;
    }
    // C = A * M^T skipping zero entries of M.
    static std::vector<S> mul_skip_t(const std::vector<S>& A, const std::vector<S>& M, int N) {
        std::vector<S> C((size_t)N * N, S(0));
        std::vector<int> nz; std::vector<S> val;
        for (int j = 0; j < N; ++j) {
            nz.clear(); val.clear();
            for (int k = 0; k < N; ++k) if (M[(size_t)j * N + k] != S(0)) { nz.push_back(k); val.push_back(M[(size_t)j * N + k]); }
            for (int i = 0; i < N; ++i) {
                const S* a = &A[(size_t)i * N];
                S s = S(0);
                for (size_t t = 0; t < nz.size(); ++t) s += a[nz[t]] * val[t];
                C[(size_t)i * N + j] = s;
            }
        }
        return C;
    }

    // TightlyCoupledEKF.cpp:96-121
    void process(S dt) {
        const int N = dim();
        std::vector<S> F = numericallyLinearizeProcess(dt);
        for (auto& f : feat) f = convolveFeature(base_mu, f, dt);     // with the OLD base state
        base_mu = convolveBaseState(base_mu, dt);
        std::vector<S> T = mul_skip(F, Sigma, N);
        Sigma = mul_skip_t(T, F, N);
        std::vector<S> q = generateProcessNoise(dt);
        for (int i = 0; i < N; ++i) Sigma[(size_t)i * N + i] += q[i];
        prune(Sigma);
    }

    // TightlyCoupledEKF.cpp:634-661 — H as the list of selected state columns.
    std::vector<int> formFeatureMeasurementMap(const uint8_t* measured) const {
        std::vector<int> cols;
        for (size_t i = 0; i < feat.size(); ++i)
            if (measured[i]) { cols.push_back((int)i * 3 + BASE); cols.push_back((int)i * 3 + BASE + 1); }
        return cols;
    }

    // TightlyCoupledEKF.cpp:475-628.  z: n x 2, R: n x 4 (row-major 2x2), pass: n.
    void updateWithFeaturePositions(const S* zin, const S* Rin, const uint8_t* pass) {
        const int N = dim(), n = (int)feat.size();
        std::vector<int> idx = formFeatureMeasurementMap(pass);
        const int m = (int)idx.size();
        std::vector<S> R((size_t)m * m, S(0)), z(m), mu(N);
        for (int i = 0; i < BASE; ++i) mu[i] = base_mu[i];
        int j = 0;
        for (int i = 0; i < n; ++i) {
            if (pass[i]) {
                klt_last[i] = {zin[2 * i], zin[2 * i + 1]};
                z[j] = zin[2 * i];
                R[(size_t)j * m + j] = Rin[4 * i + 0];
                ++j;
                z[j] = zin[2 * i + 1];
                R[(size_t)j * m + j] = Rin[4 * i + 3];
                R[(size_t)(j - 1) * m + j] = Rin[4 * i + 1];
                R[(size_t)j * m + (j - 1)] = Rin[4 * i + 2];
                ++j;
            } else {
                delete_flag[i] = 1;
            }
            mu[BASE + 3 * i] = feat[i][0]; mu[BASE + 3 * i + 1] = feat[i][1]; mu[BASE + 3 * i + 2] = feat[i][2];
        }
        if (m == 0) {
            // H has no rows: K is N x 0, I_KH = I, Sigma = I*Sigma*I + 0, mu unchanged, quaternion renormalised.
            finish_update(mu);
            prune(Sigma);
            return;
        }
        std::vector<S> y(m);
        for (int a = 0; a < m; ++a) y[a] = z[a] - mu[idx[a]];
        // S = H Sigma H' + R
        std::vector<S> Sm((size_t)m * m);
        for (int a = 0; a < m; ++a)
            for (int b = 0; b < m; ++b) Sm[(size_t)a * m + b] = Sigma[(size_t)idx[a] * N + idx[b]] + R[(size_t)a * m + b];
        // SimplicialLDLT::compute(S^T) reads the lower triangle of S^T == the upper triangle of S.
        // LDL^T without pivoting, natural ordering (Eigen's AMD permutation only changes rounding).
        std::vector<S> L((size_t)m * m, S(0)), D(m);
        bool ok = true;
        for (int c = 0; c < m && ok; ++c) {
            S d = Sm[(size_t)c * m + c];
            for (int k = 0; k < c; ++k) d -= L[(size_t)c * m + k] * L[(size_t)c * m + k] * D[k];
            D[c] = d;
            if (d < S(0)) status |= 8;                                 // diagnostic only: S is not positive definite (the reference carries on)
            if (d == S(0)) { ok = false; break; }
            for (int r = c + 1; r < m; ++r) {
                S v = Sm[(size_t)c * m + r];                           // upper(S)(c,r) mirrored to (r,c)
                for (int k = 0; k < c; ++k) v -= L[(size_t)r * m + k] * L[(size_t)c * m + k] * D[k];
                L[(size_t)r * m + c] = v / d;
            }
        }
        if (!ok) status |= 1;
        // K' = solve((Sigma H')')  -> for each state row i: K(i,:) = S^-1 * Sigma(i, idx)
        std::vector<S> K((size_t)N * m);
        const S lim = S(SPARSE_THRESH) * S(SPARSE_EPS);
        std::vector<S> w(m);
        for (int i = 0; i < N; ++i) {
            for (int a = 0; a < m; ++a) w[a] = Sigma[(size_t)i * N + idx[a]];
            for (int a = 0; a < m; ++a) { S v = w[a]; for (int k = 0; k < a; ++k) v -= L[(size_t)a * m + k] * w[k]; w[a] = v; }
            for (int a = 0; a < m; ++a) w[a] /= D[a];
            for (int a = m - 1; a >= 0; --a) { S v = w[a]; for (int k = a + 1; k < m; ++k) v -= L[(size_t)k * m + a] * w[k]; w[a] = v; }
            for (int a = 0; a < m; ++a) K[(size_t)i * m + a] = (std::abs(w[a]) > lim) ? w[a] : S(0);   // sparseView
        }
        // I_KH = I - K H, pruned
        std::vector<S> IKH((size_t)N * N, S(0));
        for (int i = 0; i < N; ++i) IKH[(size_t)i * N + i] = S(1);
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < m; ++a) IKH[(size_t)i * N + idx[a]] -= K[(size_t)i * m + a];
        prune(IKH);
        // Sigma = I_KH Sigma I_KH' + K R K'
        std::vector<S> T = mul_skip(IKH, Sigma, N);
        std::vector<S> NS = mul_skip_t(T, IKH, N);
        std::vector<S> KR((size_t)N * m, S(0));
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < m; ++a) {
                S k = K[(size_t)i * m + a];
                if (k == S(0)) continue;
                for (int b = 0; b < m; ++b) { S r = R[(size_t)a * m + b]; if (r != S(0)) KR[(size_t)i * m + b] += k * r; }
            }
        for (int i = 0; i < N; ++i)
            for (int jj = 0; jj < N; ++jj) {
                S s = S(0);
                for (int a = 0; a < m; ++a) s += KR[(size_t)i * m + a] * K[(size_t)jj * m + a];
                NS[(size_t)i * N + jj] += s;
            }
        Sigma.swap(NS);
        for (int i = 0; i < N; ++i) {
            S s = S(0);
            for (int a = 0; a < m; ++a) s += K[(size_t)i * m + a] * y[a];
            mu[i] += s;
        }
        finish_update(mu);
        prune(Sigma);
    }

    // TightlyCoupledEKF.cpp:699-714 — returns #negative diagonal entries and max |Sij - Sji|.
    void checkSigma(int* neg_diag, double* max_asym) const {
        const int N = BASE + 3 * (int)feat.size();
        int neg = 0; double mx = 0;
        for (int i = 0; i < N; ++i) {
            if (Sigma[(size_t)i * N + i] < S(0)) ++neg;
            for (int j = i + 1; j < N; ++j) {
                double a = std::fabs((double)(Sigma[(size_t)i * N + j] - Sigma[(size_t)j * N + i]));
                if (a > mx) mx = a;
            }
        }
        *neg_diag = neg; *max_asym = mx;
    }

private:
    // TightlyCoupledEKF.cpp:604-620
    void finish_update(std::vector<S>& mu) {
        S qn = std::sqrt(mu[3] * mu[3] + mu[4] * mu[4] + mu[5] * mu[5] + mu[6] * mu[6]);
        mu[3] /= qn; mu[4] /= qn; mu[5] /= qn; mu[6] /= qn;
        for (int i = 0; i < BASE; ++i) base_mu[i] = mu[i];
        for (size_t i = 0; i < feat.size(); ++i) {
            feat[i][0] = mu[BASE + 3 * i]; feat[i][1] = mu[BASE + 3 * i + 1]; feat[i][2] = mu[BASE + 3 * i + 2];
        }
    }
};

// ---------------------------------------------------------------------------------------------
// cv::RNG restatement (OpenCV core, used only by test/analyzeEKFSimulation.cpp:11-21).
// MWC generator + Marsaglia-Tsang ziggurat (randn_0_1_32f).  Checked bit-exact against
// cv2.setRNGSeed(0); cv2.randn in tests/test_oracle_ekf.py.
struct CvRNG {
    uint64_t state;
    explicit CvRNG(uint64_t s = 0xffffffff) : state(s ? s : 0xffffffff) {}
    static uint64_t adv(uint64_t x) { return (uint64_t)(unsigned)x * 4164903690U + (unsigned)(x >> 32); }
    unsigned next() { state = adv(state); return (unsigned)state; }
    double uniform(double a, double b) {
        unsigned t = next();
        double u = (double)(((uint64_t)t << 32) | next()) * 5.4210108624275221700372640043497e-20;
        return u * (b - a) + a;
    }
    double gaussian(double sigma) { return (double)randn32f() * sigma; }
    float randn32f() {
        static unsigned kn[128]; static float wn[128], fn[128]; static bool init = false;
        const float r = 3.442620f, rng_flt = 2.3283064365386962890625e-10f;
        if (!init) {
            const double m1 = 2147483648.0;
            double dn = 3.442619855899, tn = dn, vn = 9.91256303526217e-3;
            double q = vn / std::exp(-.5 * dn * dn);
            kn[0] = (unsigned)((dn / q) * m1); kn[1] = 0;
            wn[0] = (float)(q / m1); wn[127] = (float)(dn / m1);
            fn[0] = 1.f; fn[127] = (float)std::exp(-.5 * dn * dn);
            for (int i = 126; i >= 1; i--) {
                dn = std::sqrt(-2. * std::log(vn / dn + std::exp(-.5 * dn * dn)));
                kn[i + 1] = (unsigned)((dn / tn) * m1);
                tn = dn;
                fn[i] = (float)std::exp(-.5 * dn * dn);
                wn[i] = (float)(dn / m1);
            }
            init = true;
        }
        uint64_t temp = state;
        float x, y;
        for (;;) {
            int hz = (int)temp;
            temp = adv(temp);
            int iz = hz & 127;
            x = hz * wn[iz];
            if ((unsigned)std::abs(hz) < kn[iz]) break;
            if (iz == 0) {
                do {
                    x = (unsigned)temp * rng_flt; temp = adv(temp);
                    y = (unsigned)temp * rng_flt; temp = adv(temp);
                    x = (float)(-std::log(x + 1.175494351e-38F) * 0.2904764);
                    y = (float)-std::log(y + 1.175494351e-38F);
                } while (y + y < x * x);
                x = hz > 0 ? r + x : -r - x;
                break;
            }
            y = (unsigned)temp * rng_flt; temp = adv(temp);
            if (fn[iz] + y * (fn[iz - 1] - fn[iz]) < std::exp(-.5 * x * x)) break;
        }
        state = temp;
        return x;
    }
};

// ---------------------------------------------------------------------------------------------
// Scenario generator of test/analyzeEKFSimulation.cpp:10-125 (simulateAndVisualizeEKF +
// generateFakeMeasurementsAndUpdateEKF): landmarks from cv::RNG(0), ground truth propagated in
// float exactly as the reference does, measurements = exact projections in float.
struct SimScenario {
    int n = 0, steps = 0;
    float dt = 0;
    std::vector<float> init_uv;          // n x 2
    std::vector<float> meas;             // steps x n x 2
};

inline SimScenario make_scenario(int feature_count, float depth_sigma, float depth_mu, const float b_vel[3],
                                 const float b_accel[3], const float omega_in[3], float dt, float tf, uint64_t seed = 0) {
    SimScenario sc; sc.n = feature_count; sc.dt = dt;
    CvRNG rng(seed);
    std::vector<Vec3<float>> gt(feature_count);
    for (int i = 0; i < feature_count; ++i) {
        Vec3<float> p;
        p.z = (float)(depth_mu + rng.gaussian(depth_sigma));
        p.x = (float)(rng.uniform(-1.5, 1.5) * p.z);
        p.y = (float)(rng.uniform(-1.5, 1.5) * p.z);
        gt[i] = p;
        sc.init_uv.push_back(p.x / p.z); sc.init_uv.push_back(p.y / p.z);
    }
    Vec3<float> pos{0, 0, 0}, vel{b_vel[0], b_vel[1], b_vel[2]}, accel{b_accel[0], b_accel[1], b_accel[2]};
    Vec3<float> omega{omega_in[0], omega_in[1], omega_in[2]};
    Quat<float> quat{1, 0, 0, 0};
    for (float t = dt; t <= tf; t += dt) {                   // float accumulation decides the step count
        float hdt2 = (float)(0.5 * (double)dt * (double)dt);
        Vec3<float> tr{dt * vel.x + hdt2 * accel.x, dt * vel.y + hdt2 * accel.y, dt * vel.z + hdt2 * accel.z};
        Vec3<float> r = rotate(quat, tr);
        pos = {pos.x + r.x, pos.y + r.y, pos.z + r.z};
        float on = std::sqrt(omega.x * omega.x + omega.y * omega.y + omega.z * omega.z);
        Quat<float> dq;
        if (on < 1e-10f) {
            dq = qnormalized(Quat<float>{1.0f, omega.x * dt, omega.y * dt, omega.z * dt});
        } else {
            float theta = dt * on;
            Vec3<float> oh{omega.x / on, omega.y / on, omega.z / on};
            float st2 = std::sin(theta / 2);
            dq = {std::cos(theta / 2), oh.x * st2, oh.y * st2, oh.z * st2};
        }
        Quat<float> dqi = qinv(dq);
        Vec3<float> va{vel.x + dt * accel.x, vel.y + dt * accel.y, vel.z + dt * accel.z};
        vel = rotate(dqi, va);
        accel = rotate(dqi, accel);
        quat = qmul(quat, dq);
        Quat<float> qi = qinv(quat);
        for (int i = 0; i < feature_count; ++i) {
            Vec3<float> a = rotate(qi, gt[i]), b = rotate(qi, pos);
            Vec3<float> fp{a.x - b.x, a.y - b.y, a.z - b.z};
            sc.meas.push_back(fp.x / fp.z); sc.meas.push_back(fp.y / fp.z);
        }
        ++sc.steps;
    }
    return sc;
}

}  // namespace ekf_oracle
