// TEST INFRASTRUCTURE ONLY — C entry points around the reference's own TightlyCoupledEKF, compiled from the UNMODIFIED sources
// where they lie (/root/reference/include/ekf_vio/{TightlyCoupledEKF,Feature,Params}.cpp, found through -I) against the
// stand-in headers of oracle/_shim (Eigen, ROS and OpenCV are not installed in this image).  Built by `make -C oracle ref` into
// oracle/_ref/libekf_ref_f32.so (the reference as written: float) and, with -DEKFVIO_REF_F64, libekf_ref_f64.so, in which
// the token `float` of the reference sources is mapped to `double` so the very same statements run in FP64 — the instance
// the FP64 oracle (oracle/ekf_oracle.hpp) and through it the GPU path are pinned to at 1e-9.
// No reference source is copied: they are #included from the read-only tree at build time.
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <string>
#include <string.h>
#include <vector>

#ifdef EKFVIO_REF_F64
#define EKFVIO_SHIM_SCALAR double
#else
#define EKFVIO_SHIM_SCALAR float
#endif
// the stand-ins first, parsed with the real `float` keyword (their include guards keep the reference's own #includes inert)
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include <ros/ros.h>
#include <sensor_msgs/CameraInfo.h>

#ifdef EKFVIO_REF_F64
#define float double
#endif
#include "Params.cpp"
#include "Feature.cpp"
#include "TightlyCoupledEKF.cpp"
#ifdef EKFVIO_REF_F64
#undef float
#endif

typedef EKFVIO_SHIM_SCALAR S;
typedef Eigen::Matrix<S, BASE_STATE_SIZE, 1> BaseVec;

static TightlyCoupledEKF* F_(void* h) { return static_cast<TightlyCoupledEKF*>(h); }

extern "C" {

int ekfref_scalar_bytes() { return (int)sizeof(S); }

// Params.h globals are process-wide in the reference (EKFVIO.cpp:20-67 reads them from ROS params; the tests set them by hand,
// test_ekf.cpp:22-24)
void* ekfref_create(double depth, double depth_var, double uv_var) {
    DEFAULT_POINT_DEPTH = depth; DEFAULT_POINT_DEPTH_VARIANCE = depth_var; DEFAULT_POINT_HOMOGENOUS_VARIANCE = uv_var;
    NUM_FEATURES = D_NUM_FEATURES;
    return new TightlyCoupledEKF();
}
void ekfref_destroy(void* h) { delete F_(h); }
void* ekfref_clone(void* h) { return new TightlyCoupledEKF(*F_(h)); }      // the reference's filters are copy-assignable (test_ekf.cpp:97)

// convolveFeature's function-static dq_inv cache (TightlyCoupledEKF.cpp:400-403) is shared by every filter of the process (E2):
// a call with omega = 0 puts it back to its initial content (last omega 0, identity rotation).
void ekfref_reset_static_cache() {
    TightlyCoupledEKF f;
    BaseVec b; b.setZero(); b(3) = 1; b(10) = 1;           // first a non-zero omega, so that the zero omega below is seen as a change
    Eigen::Vector3f x(0, 0, 1);
    f.convolveFeature(b, x, 0);
    b(10) = 0;
    f.convolveFeature(b, x, 0);
}

int ekfref_num_features(void* h) { return (int)F_(h)->features.size(); }

void ekfref_add_features(void* h, const double* uv, int k) {
    std::vector<Eigen::Vector2f> v;
    for (int i = 0; i < k; ++i) v.push_back(Eigen::Vector2f((S)uv[2 * i], (S)uv[2 * i + 1]));
    F_(h)->addNewFeatures(v);
}

void ekfref_process(void* h, double dt) { F_(h)->process((S)dt); }

void ekfref_update(void* h, const double* z, const double* R, const uint8_t* pass) {
    const int n = ekfref_num_features(h);
    std::vector<Eigen::Vector2f> zs; std::vector<Eigen::Matrix2f> Rs; std::vector<bool> ps;
    for (int i = 0; i < n; ++i) {
        zs.push_back(Eigen::Vector2f((S)z[2 * i], (S)z[2 * i + 1]));
        Eigen::Matrix2f r; r(0, 0) = (S)R[4 * i]; r(0, 1) = (S)R[4 * i + 1]; r(1, 0) = (S)R[4 * i + 2]; r(1, 1) = (S)R[4 * i + 3];
        Rs.push_back(r); ps.push_back(pass[i] != 0);
    }
    F_(h)->updateWithFeaturePositions(zs, Rs, ps);
}

void ekfref_get_state(void* h, double* mu, double* feat, double* P, uint8_t* flags, double* klt_last) {
    TightlyCoupledEKF* f = F_(h);
    const int N = (int)f->Sigma.rows();
    if (mu) for (int i = 0; i < BASE_STATE_SIZE; ++i) mu[i] = f->base_mu(i);
    int j = 0;
    for (auto& e : f->features) {
        if (feat) for (int k = 0; k < 3; ++k) feat[3 * j + k] = e.getMu()(k);
        if (flags) flags[j] = e.flaggedForDeletion() ? 1 : 0;
        if (klt_last) { Eigen::Vector2f l = e.getLastResultFromKLTTracker(); klt_last[2 * j] = l(0); klt_last[2 * j + 1] = l(1); }
        ++j;
    }
    if (P) for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) P[(size_t)r * N + c] = f->Sigma.coeff(r, c);
}

// callers of the reference write base_mu and feature means directly (test_ekf.cpp:170-199, jacobian_test.cpp:39-43)
void ekfref_set_mean(void* h, const double* mu, const double* feat) {
    TightlyCoupledEKF* f = F_(h);
    if (mu) for (int i = 0; i < BASE_STATE_SIZE; ++i) f->base_mu(i) = (S)mu[i];
    if (feat) { int j = 0; for (auto& e : f->features) { e.setMu(Eigen::Vector3f((S)feat[3 * j], (S)feat[3 * j + 1], (S)feat[3 * j + 2])); ++j; } }
}
void ekfref_set_sigma(void* h, const double* P) {
    TightlyCoupledEKF* f = F_(h);
    const int N = (int)f->Sigma.rows();
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) f->Sigma.coeffRef(r, c) = (S)P[(size_t)r * N + c];
}

void ekfref_linearize(void* h, double dt, double* out) {
    TightlyCoupledEKF* f = F_(h);
    Eigen::SparseMatrix<S> J = f->numericallyLinearizeProcess(f->base_mu, f->features, (S)dt);
    const int N = (int)J.rows();
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) out[(size_t)r * N + c] = J.coeff(r, c);
}
void ekfref_convolve_base(void* h, const double* mu, double dt, double* out) {
    BaseVec b; for (int i = 0; i < BASE_STATE_SIZE; ++i) b(i) = (S)mu[i];
    BaseVec r = F_(h)->convolveBaseState(b, (S)dt);
    for (int i = 0; i < BASE_STATE_SIZE; ++i) out[i] = r(i);
}
void ekfref_convolve_feature(void* h, const double* mu, const double* f3, double dt, double* out) {
    BaseVec b; for (int i = 0; i < BASE_STATE_SIZE; ++i) b(i) = (S)mu[i];
    Eigen::Vector3f x((S)f3[0], (S)f3[1], (S)f3[2]);
    Eigen::Vector3f r = F_(h)->convolveFeature(b, x, (S)dt);
    for (int i = 0; i < 3; ++i) out[i] = r(i);
}
void ekfref_process_noise(void* h, double dt, double* qdiag) {
    Eigen::SparseMatrix<S> Q = F_(h)->generateProcessNoise((S)dt);
    for (int i = 0; i < (int)Q.rows(); ++i) qdiag[i] = Q.coeff(i, i);
}
int ekfref_measurement_map(void* h, const uint8_t* measured, double* H) {
    std::vector<bool> m; const int n = ekfref_num_features(h);
    for (int i = 0; i < n; ++i) m.push_back(measured[i] != 0);
    Eigen::SparseMatrix<S> Hm = F_(h)->formFeatureMeasurementMap(m);
    for (int r = 0; r < (int)Hm.rows(); ++r) for (int c = 0; c < (int)Hm.cols(); ++c) H[(size_t)r * Hm.cols() + c] = Hm.coeff(r, c);
    return (int)Hm.rows();
}
// number of ROS_FATAL lines checkSigma raises (the reference's pass criterion: none)
long ekfref_check_sigma(void* h) {
    const long before = ekfvio_shim::fatal_count();
    F_(h)->checkSigma();
    return ekfvio_shim::fatal_count() - before;
}
long ekfref_error_count() { return ekfvio_shim::error_count(); }
double ekfref_feature_depth_variance(void* h, int i) { return F_(h)->getFeatureDepthVariance(i); }
void ekfref_feature_homogenous_covariance(void* h, int i, double* c4) {
    Eigen::Matrix2f c = F_(h)->getFeatureHomogenousCovariance(i);
    c4[0] = c(0, 0); c4[1] = c(0, 1); c4[2] = c(1, 0); c4[3] = c(1, 1);
}
void ekfref_set_feature_homogenous_covariance(void* h, int i, const double* c4) {
    Eigen::Matrix2f c; c(0, 0) = (S)c4[0]; c(0, 1) = (S)c4[1]; c(1, 0) = (S)c4[2]; c(1, 1) = (S)c4[3];
    F_(h)->setFeatureHomogenousCovariance(i, c);
}
void ekfref_pixel_maps(void* h, const double* K9_rowmajor, double* m2p_diag, double* p2m_diag) {
    Eigen::Matrix3f K; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) K(r, c) = (S)K9_rowmajor[3 * r + c];
    Eigen::SparseMatrix<S> a = F_(h)->getMetric2PixelMap(K), b = F_(h)->getPixel2MetricMap(K);
    m2p_diag[0] = a.coeff(0, 0); m2p_diag[1] = a.coeff(1, 1); p2m_diag[0] = b.coeff(0, 0); p2m_diag[1] = b.coeff(1, 1);
}

}  // extern "C"
