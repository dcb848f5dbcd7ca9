// Host-side C++ facade: the reference's TightlyCoupledEKF interface over the C ABI (ekfvio_c.h).
// All filter arithmetic runs in the CUDA library; this file only moves state across the boundary
// and keeps the reference's public members coherent.  Contract violations that the
// reference turns into ROS_ASSERT aborts (TightlyCoupledEKF.cpp:478,636) abort here too.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/ekf_vio/TightlyCoupledEKF.h"
#include "../../../include/ekfvio_c.h"

double INVERSE_IMAGE_SCALE = D_INVERSE_IMAGE_SCALE;
int KILL_PAD = D_KILL_PAD;
double KLT_MIN_EIGEN = D_KLT_MIN_EIGEN;
int NUM_FEATURES = D_NUM_FEATURES;
double DEFAULT_POINT_DEPTH = D_DEFAULT_POINT_DEPTH;
double DEFAULT_POINT_DEPTH_VARIANCE = D_DEFAULT_POINT_DEPTH_VARIANCE;
double DEFAULT_POINT_HOMOGENOUS_VARIANCE = D_DEFAULT_POINT_HOMOGENOUS_VARIANCE;
int WINDOW_SIZE = D_WINDOW_SIZE;
int MAX_PYRAMID_LEVEL = D_MAX_PYRAMID_LEVEL;

#define EKF_ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ASSERTION FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::abort(); } } while (0)
#define EKF_CALL(x) do { if ((x) != 0) { std::fprintf(stderr, "ekfvio: %s failed: %s\n", #x, ekfvio_last_error()); std::abort(); } } while (0)

namespace {
ekfvio_params current_params() {
    ekfvio_params p;
    ekfvio_default_params(&p);
    p.default_point_depth = DEFAULT_POINT_DEPTH;
    p.default_point_depth_variance = DEFAULT_POINT_DEPTH_VARIANCE;
    p.default_point_homogenous_variance = DEFAULT_POINT_HOMOGENOUS_VARIANCE;
    return p;
}
}  // namespace

TightlyCoupledEKF::TightlyCoupledEKF() {
    this->t = ros::Time(0);
    recreate(NUM_FEATURES > 8 ? NUM_FEATURES : 8);
    pull();
}

TightlyCoupledEKF::~TightlyCoupledEKF() { if (dev_) ekfvio_batch_destroy(dev_); }

TightlyCoupledEKF::TightlyCoupledEKF(const TightlyCoupledEKF& o) { *this = o; }

TightlyCoupledEKF& TightlyCoupledEKF::operator=(const TightlyCoupledEKF& o) {
    if (this == &o) return *this;
    // value semantics (tests do `tc_ekf = TightlyCoupledEKF();`, test_ekf.cpp:97): deep copy of the device state
    const_cast<TightlyCoupledEKF&>(o).pushIfEdited();
    if (dev_) { ekfvio_batch_destroy(dev_); dev_ = nullptr; }
    recreate(o.capacity_);
    const int nm = capacity_ > 0 ? capacity_ : 1, Nm = BASE_STATE_SIZE + 3 * capacity_;
    std::vector<double> mu(BASE_STATE_SIZE), feat((size_t)nm * 3), P((size_t)Nm * Nm), cache(7), klt((size_t)nm * 2);
    std::vector<uint8_t> flags(nm);
    int nfeat = 0, status = 0;
    EKF_CALL(ekfvio_batch_get_state(o.dev_, mu.data(), feat.data(), P.data(), &nfeat, cache.data(), flags.data(), klt.data(), &status));
    EKF_CALL(ekfvio_batch_set_state(dev_, mu.data(), feat.data(), nullptr, &nfeat, cache.data(), flags.data(), klt.data()));
    EKF_CALL(ekfvio_batch_set_state(dev_, nullptr, nullptr, P.data(), nullptr, nullptr, nullptr, nullptr));
    this->t = o.t;
    pull();
    return *this;
}

void TightlyCoupledEKF::recreate(int capacity) {
    ekfvio_params p = current_params();
    capacity_ = capacity;
    EKF_CALL(ekfvio_batch_create(&dev_, 0, 1, capacity_, &p));
}

// Grow the single-filter batch when more features are added than it has room for.
void TightlyCoupledEKF::ensureCapacity(int n_features) {
    if (n_features <= capacity_) return;
    int newcap = capacity_ * 2 > n_features ? capacity_ * 2 : n_features;
    const int nm0 = capacity_ > 0 ? capacity_ : 1, Nm0 = BASE_STATE_SIZE + 3 * capacity_;
    std::vector<double> mu(BASE_STATE_SIZE), feat((size_t)nm0 * 3), P((size_t)Nm0 * Nm0), cache(7), klt((size_t)nm0 * 2);
    std::vector<uint8_t> flags(nm0);
    int nfeat = 0, status = 0;
    EKF_CALL(ekfvio_batch_get_state(dev_, mu.data(), feat.data(), P.data(), &nfeat, cache.data(), flags.data(), klt.data(), &status));
    ekfvio_batch_destroy(dev_); dev_ = nullptr;
    recreate(newcap);
    const int Nm1 = BASE_STATE_SIZE + 3 * newcap;
    std::vector<double> feat1((size_t)newcap * 3, 0.0), P1((size_t)Nm1 * Nm1, 0.0), klt1((size_t)newcap * 2, 0.0);
    std::vector<uint8_t> flags1(newcap, 0);
    std::memcpy(feat1.data(), feat.data(), sizeof(double) * 3 * nfeat);
    std::memcpy(klt1.data(), klt.data(), sizeof(double) * 2 * nfeat);
    std::memcpy(flags1.data(), flags.data(), nfeat);
    const int N = BASE_STATE_SIZE + 3 * nfeat;
    for (int i = 0; i < N; ++i) std::memcpy(&P1[(size_t)i * Nm1], &P[(size_t)i * Nm0], sizeof(double) * N);
    EKF_CALL(ekfvio_batch_set_state(dev_, mu.data(), feat1.data(), nullptr, &nfeat, cache.data(), flags1.data(), klt1.data()));
    EKF_CALL(ekfvio_batch_set_state(dev_, nullptr, nullptr, P1.data(), nullptr, nullptr, nullptr, nullptr));
}

// Refresh the public members (float) from the device state (double) and remember what was shown.
void TightlyCoupledEKF::pull() {
    const int nm = capacity_ > 0 ? capacity_ : 1, Nm = BASE_STATE_SIZE + 3 * capacity_;
    dmu_.assign(BASE_STATE_SIZE, 0.0); dfeat_.assign((size_t)nm * 3, 0.0); dP_.assign((size_t)Nm * Nm, 0.0);
    std::vector<double> klt((size_t)nm * 2);
    std::vector<uint8_t> flags(nm);
    int nfeat = 0;
    EKF_CALL(ekfvio_batch_get_state(dev_, dmu_.data(), dfeat_.data(), dP_.data(), &nfeat, nullptr, flags.data(), klt.data(), nullptr));
    for (int i = 0; i < BASE_STATE_SIZE; ++i) base_mu(i) = (float)dmu_[i];
    // keep the list nodes (callers hold references: features.front().getMu(), test_ekf.cpp:198)
    while ((int)features.size() < nfeat) features.push_back(Feature());
    while ((int)features.size() > nfeat) features.pop_back();
    int i = 0;
    for (auto& e : features) {
        e.setMu(Eigen::Vector3f((float)dfeat_[3 * i], (float)dfeat_[3 * i + 1], (float)dfeat_[3 * i + 2]));
        e.setLastResultFromKLTTracker(Eigen::Vector2f((float)klt[2 * i], (float)klt[2 * i + 1]));
        e.setDeleteFlag(flags[i] != 0);
        ++i;
    }
    const int N = BASE_STATE_SIZE + 3 * nfeat;
    Sigma.resize(N, N);
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) Sigma(r, c) = (float)dP_[(size_t)r * Nm + c];
    snap_mu_.resize(BASE_STATE_SIZE); for (int k = 0; k < BASE_STATE_SIZE; ++k) snap_mu_[k] = base_mu(k);
    snap_feat_.resize((size_t)3 * nfeat); i = 0;
    for (auto& e : features) { for (int c = 0; c < 3; ++c) snap_feat_[3 * i + c] = e.getMu()(c); ++i; }
    snap_sigma_.assign(Sigma.data(), Sigma.data() + (size_t)N * N);
}

// If the caller changed base_mu / a feature's mu / Sigma since the last refresh, upload the edit
// (the edited entries in float precision, untouched entries keep their FP64 device values).
void TightlyCoupledEKF::pushIfEdited() {
    const int nm = capacity_ > 0 ? capacity_ : 1, Nm = BASE_STATE_SIZE + 3 * capacity_;
    bool mu_edit = false, feat_edit = false, sig_edit = false;
    for (int k = 0; k < BASE_STATE_SIZE; ++k) if (base_mu(k) != snap_mu_[k]) { dmu_[k] = base_mu(k); mu_edit = true; }
    int i = 0;
    const int nfeat = (int)snap_feat_.size() / 3;
    EKF_ASSERT((int)features.size() == nfeat);   // features are added through addNewFeatures only
    for (auto& e : features) {
        for (int c = 0; c < 3; ++c) if (e.getMu()(c) != snap_feat_[3 * i + c]) { dfeat_[3 * i + c] = e.getMu()(c); feat_edit = true; }
        ++i;
    }
    const int N = BASE_STATE_SIZE + 3 * nfeat;
    if (Sigma.rows() == N && Sigma.cols() == N) {
        const float* s = Sigma.data();
        for (int c = 0; c < N; ++c) for (int r = 0; r < N; ++r) if (s[(size_t)c * N + r] != snap_sigma_[(size_t)c * N + r]) { dP_[(size_t)r * Nm + c] = s[(size_t)c * N + r]; sig_edit = true; }
    }
    (void)nm;
    if (mu_edit || feat_edit) EKF_CALL(ekfvio_batch_set_state(dev_, mu_edit ? dmu_.data() : nullptr, feat_edit ? dfeat_.data() : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    if (sig_edit) EKF_CALL(ekfvio_batch_set_state(dev_, nullptr, nullptr, dP_.data(), nullptr, nullptr, nullptr, nullptr));
}

void TightlyCoupledEKF::initializeBaseState() {   // TightlyCoupledEKF.cpp:23-56
    EKF_CALL(ekfvio_batch_reset(dev_, nullptr));
    pull();
}

void TightlyCoupledEKF::addNewFeatures(std::vector<Eigen::Vector2f> f) {   // :58-94
    if (!f.size()) return;
    pushIfEdited();
    ensureCapacity((int)features.size() + (int)f.size());
    std::vector<double> uv(f.size() * 2);
    for (size_t i = 0; i < f.size(); ++i) { uv[2 * i] = f[i].x(); uv[2 * i + 1] = f[i].y(); }
    int k = (int)f.size();
    EKF_CALL(ekfvio_batch_add_features_h(dev_, &k, uv.data(), k, nullptr));
    pull();
}

std::vector<Eigen::Vector2f> TightlyCoupledEKF::previousFeaturePositionVector() {   // :462-470
    std::vector<Eigen::Vector2f> out;
    for (auto& e : features) out.push_back(e.getLastResultFromKLTTracker());
    return out;
}

void TightlyCoupledEKF::process(float dt) {   // :96-121
    pushIfEdited();
    EKF_CALL(ekfvio_batch_process_dt(dev_, (double)dt, nullptr));
    pull();
}

void TightlyCoupledEKF::updateWithFeaturePositions(std::vector<Eigen::Vector2f> z, std::vector<Eigen::Matrix2f> R, std::vector<bool> pass) {   // :475-628
    EKF_ASSERT(z.size() == R.size() && pass.size() == features.size() && R.size() == pass.size());   // :478
    if (!pass.size()) std::fprintf(stderr, "no measurements to update state with!\n");                // :482-484
    pushIfEdited();
    const int nm = capacity_ > 0 ? capacity_ : 1;
    std::vector<double> hz((size_t)nm * 2, 0.0), hR((size_t)nm * 4, 0.0);
    std::vector<uint8_t> hp(nm, 0);
    for (size_t i = 0; i < z.size(); ++i) {
        hz[2 * i] = z[i].x(); hz[2 * i + 1] = z[i].y();
        hR[4 * i] = R[i](0, 0); hR[4 * i + 1] = R[i](0, 1); hR[4 * i + 2] = R[i](1, 0); hR[4 * i + 3] = R[i](1, 1);
        hp[i] = pass[i] ? 1 : 0;
    }
    EKF_CALL(ekfvio_batch_update_h(dev_, hz.data(), hR.data(), hp.data(), nullptr));
    pull();
    if (deviceStatus() & 1) std::fprintf(stderr, "there was a problem decomposing S... maybe it was not positive semi definite\n");   // :579
}

Eigen::SparseMatrix<float> TightlyCoupledEKF::numericallyLinearizeProcess(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& mu, std::list<Feature>& feats, float dt) {   // :176-325
    // evaluated on the device for the state passed in (normally this->base_mu / this->features).  The reference's function does
    // not touch the filter: when a foreign state is passed in, the filter's own mean is put back afterwards (only the dq_inv
    // cache moves, as the reference's function-static cache does).
    EKF_ASSERT(feats.size() == features.size());
    const bool foreign = (&mu != &base_mu) || (&feats != &features);
    Eigen::Matrix<float, BASE_STATE_SIZE, 1> saved_mu = base_mu;
    std::vector<Eigen::Vector3f> saved_feat;
    if (foreign) {
        pushIfEdited();                                    // caller edits of the filter's own members first
        for (auto& e : features) saved_feat.push_back(e.getMu());
        for (int k = 0; k < BASE_STATE_SIZE; ++k) base_mu(k) = mu(k);
        if (&feats != &features) { auto it = features.begin(); for (auto& e : feats) { it->setMu(e.getMu()); ++it; } }
    }
    pushIfEdited();
    const int Nm = BASE_STATE_SIZE + 3 * capacity_, N = BASE_STATE_SIZE + 3 * (int)features.size();
    std::vector<double> F((size_t)Nm * Nm);
    EKF_CALL(ekfvio_batch_linearize_h(dev_, (double)dt, F.data()));
    Eigen::SparseMatrix<float> out(N, N);
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) out(r, c) = (float)F[(size_t)r * Nm + c];
    if (foreign) {
        pull();                                            // snapshot = the foreign state now on the device ...
        base_mu = saved_mu;
        size_t i = 0;
        for (auto& e : features) e.setMu(saved_feat[i++]);
        pushIfEdited();                                    // ... so that putting the filter's own mean back registers as an edit
    }
    pull();
    return out;
}

// convolveBaseState / convolveFeature (:328-460): single evaluations on the device.  convolveFeature goes through this
// filter's dq_inv cache with the reference's rule — keyed on omega only, so a call with an unchanged omega and a different dt
// reuses the stale rotation (E2), and a miss leaves the cache at (omega, this dt) for the next process().
Eigen::Matrix<float, BASE_STATE_SIZE, 1> TightlyCoupledEKF::convolveBaseState(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& last, float dt) {
    double mu[BASE_STATE_SIZE], out[BASE_STATE_SIZE];
    for (int k = 0; k < BASE_STATE_SIZE; ++k) mu[k] = last(k);
    EKF_CALL(ekfvio_batch_convolve_base_h(dev_, 0, mu, (double)dt, out));
    Eigen::Matrix<float, BASE_STATE_SIZE, 1> r;
    for (int k = 0; k < BASE_STATE_SIZE; ++k) r(k) = (float)out[k];
    return r;
}

Eigen::Vector3f TightlyCoupledEKF::convolveFeature(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& base_state, Eigen::Vector3f& feature_state, float dt) {
    double mu[BASE_STATE_SIZE], f3[3] = {feature_state(0), feature_state(1), feature_state(2)}, out[3];
    for (int k = 0; k < BASE_STATE_SIZE; ++k) mu[k] = base_state(k);
    EKF_CALL(ekfvio_batch_convolve_feature_h(dev_, 0, mu, f3, (double)dt, out));
    return Eigen::Vector3f((float)out[0], (float)out[1], (float)out[2]);
}

Eigen::SparseMatrix<float> TightlyCoupledEKF::generateProcessNoise(float dt) {   // :123-174 (a constant diagonal)
    int dim = BASE_STATE_SIZE + (int)features.size() * 3;
    Eigen::SparseMatrix<float> Q(dim, dim);
    float low_noise = 0.0001 * dt, pos_noise = 0.0001 * dt, velocity_noise = 0.01 * dt, omega_noise = 5 * dt, accel_noise = 5 * dt, bias_noise = 0.001 * dt;
    for (int i = 0; i <= 6; ++i) Q(i, i) = pos_noise;
    for (int i = 7; i <= 9; ++i) Q(i, i) = velocity_noise;
    for (int i = 10; i <= 12; ++i) Q(i, i) = omega_noise;
    for (int i = 13; i <= 15; ++i) Q(i, i) = accel_noise;
    for (int i = 16; i <= 21; ++i) Q(i, i) = bias_noise;
    for (int i = BASE_STATE_SIZE; i < dim; ++i) Q(i, i) = low_noise;
    return Q;
}

Eigen::SparseMatrix<float> TightlyCoupledEKF::formFeatureMeasurementMap(std::vector<bool> measured) {   // :634-661
    EKF_ASSERT(measured.size() == features.size());   // :636
    std::vector<int> indexes;
    for (size_t i = 0; i < measured.size(); i++) if (measured.at(i)) indexes.push_back((int)i * 3 + BASE_STATE_SIZE);
    Eigen::SparseMatrix<float> H((int)indexes.size() * 2, BASE_STATE_SIZE + (int)features.size() * 3);
    int row = 0;
    for (auto e : indexes) { H(row, e) = 1.0f; ++row; H(row, e + 1) = 1.0f; ++row; }
    return H;
}

Eigen::Matrix2f TightlyCoupledEKF::getFeatureHomogenousCovariance(int index) {   // :663-666
    int start = BASE_STATE_SIZE + index * 3;
    Eigen::Matrix2f m;
    m(0, 0) = Sigma(start, start); m(0, 1) = Sigma(start, start + 1); m(1, 0) = Sigma(start + 1, start); m(1, 1) = Sigma(start + 1, start + 1);
    return m;
}

void TightlyCoupledEKF::setFeatureHomogenousCovariance(int index, Eigen::Matrix2f cov) {   // :668-676
    int start = BASE_STATE_SIZE + index * 3;
    std::fprintf(stderr, "tried to set to sparse matrix\n");
    Sigma(start, start) = cov(0, 0); Sigma(start + 1, start) = cov(1, 0); Sigma(start, start + 1) = cov(0, 1); Sigma(start + 1, start + 1) = cov(1, 1);
    pushIfEdited();
    pull();
}

float TightlyCoupledEKF::getFeatureDepthVariance(int index) {   // :678-681
    int start = BASE_STATE_SIZE + index * 3 + 2;
    return Sigma(start, start);
}

Eigen::SparseMatrix<float> TightlyCoupledEKF::getMetric2PixelMap(Eigen::Matrix3f& K) {   // :683-689
    Eigen::SparseMatrix<float> J(2, 2);
    J(0, 0) = K(0, 0); J(1, 1) = K(1, 1);
    return J;
}
Eigen::SparseMatrix<float> TightlyCoupledEKF::getPixel2MetricMap(Eigen::Matrix3f& K) {   // :691-697
    Eigen::SparseMatrix<float> J(2, 2);
    J(0, 0) = 1.0f / K(0, 0); J(1, 1) = 1.0f / K(1, 1);
    return J;
}

void TightlyCoupledEKF::checkSigma() {   // :699-714, evaluated on the device
    pushIfEdited();
    EKF_CALL(ekfvio_batch_check_sigma_h(dev_, &last_check_negative_diagonals, &last_check_max_asymmetry));
    if (last_check_negative_diagonals) std::fprintf(stderr, "variance is negative for %d indices\n", last_check_negative_diagonals);
    if (last_check_max_asymmetry > 0.001) std::fprintf(stderr, "correlation is not symmetric: %g\n", last_check_max_asymmetry);
}

void TightlyCoupledEKF::fixSigma() {}   // :716-718 (a no-op in the reference)

int TightlyCoupledEKF::deviceStatus() {
    int s = 0;
    EKF_CALL(ekfvio_batch_get_state(dev_, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &s));
    return s;
}
