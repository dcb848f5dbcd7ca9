// ORACLE — TEST INFRASTRUCTURE ONLY (see ekf_oracle.hpp).  C entry points for ctypes:
// a handle around Filter<double> / Filter<float>, the cv::RNG + scenario generator of
// test/analyzeEKFSimulation.cpp, and an OpenMP batch runner used as the CPU baseline
// (bench.py cpu_baseline / --impl reference).  All array arguments are double at this boundary.
#include "ekf_oracle.hpp"

#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ekf_oracle;

namespace {
struct Handle {
    bool is_float;
    Filter<double> d;
    Filter<float> f;
};
}  // namespace

#define DISPATCH(h, expr_d, expr_f) do { if ((h)->is_float) { auto& flt = (h)->f; (void)flt; expr_f; } else { auto& flt = (h)->d; (void)flt; expr_d; } } while (0)

template <class S> static void t_add(Filter<S>& flt, const double* uv, int k) {
    std::vector<S> v(2 * (size_t)k);
    for (int i = 0; i < 2 * k; ++i) v[i] = (S)uv[i];
    flt.addNewFeatures(v.data(), k);
}
template <class S> static void t_update(Filter<S>& flt, const double* z, const double* R, const uint8_t* pass) {
    int n = (int)flt.feat.size();
    std::vector<S> zz(2 * (size_t)n), rr(4 * (size_t)n);
    for (int i = 0; i < 2 * n; ++i) zz[i] = (S)z[i];
    for (int i = 0; i < 4 * n; ++i) rr[i] = (S)R[i];
    flt.updateWithFeaturePositions(zz.data(), rr.data(), pass);
}
template <class S> static void t_get(Filter<S>& flt, double* base_mu, double* feat, double* P, double* cache7,
                                     uint8_t* del, double* klt_last, int* status) {
    int n = (int)flt.feat.size(), N = flt.dim();
    if (base_mu) for (int i = 0; i < BASE; ++i) base_mu[i] = flt.base_mu[i];
    if (feat) for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) feat[3 * i + c] = flt.feat[i][c];
    if (P) for (size_t i = 0; i < (size_t)N * N; ++i) P[i] = flt.Sigma[i];
    if (cache7) {
        cache7[0] = flt.last_omega[0]; cache7[1] = flt.last_omega[1]; cache7[2] = flt.last_omega[2];
        cache7[3] = flt.cache_dq_inv.w; cache7[4] = flt.cache_dq_inv.x; cache7[5] = flt.cache_dq_inv.y; cache7[6] = flt.cache_dq_inv.z;
    }
    if (del) for (int i = 0; i < n; ++i) del[i] = flt.delete_flag[i];
    if (klt_last) for (int i = 0; i < n; ++i) { klt_last[2 * i] = flt.klt_last[i][0]; klt_last[2 * i + 1] = flt.klt_last[i][1]; }
    if (status) *status = flt.status;
}
template <class S> static void t_set(Filter<S>& flt, const double* base_mu, const double* feat, int n, const double* P,
                                     const double* cache7, const uint8_t* del, const double* klt_last) {
    if (base_mu) for (int i = 0; i < BASE; ++i) flt.base_mu[i] = (S)base_mu[i];
    if (feat) {
        flt.feat.resize(n); flt.klt_last.resize(n, {S(0), S(0)}); flt.delete_flag.resize(n, 0);
        for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) flt.feat[i][c] = (S)feat[3 * i + c];
        flt.Sigma.resize((size_t)flt.dim() * flt.dim(), S(0));
    }
    int N = flt.dim();
    if (P) for (size_t i = 0; i < (size_t)N * N; ++i) flt.Sigma[i] = (S)P[i];
    if (cache7) {
        flt.last_omega[0] = (S)cache7[0]; flt.last_omega[1] = (S)cache7[1]; flt.last_omega[2] = (S)cache7[2];
        flt.cache_dq_inv = {(S)cache7[3], (S)cache7[4], (S)cache7[5], (S)cache7[6]};
    }
    if (del) for (int i = 0; i < (int)flt.feat.size(); ++i) flt.delete_flag[i] = del[i];
    if (klt_last) for (int i = 0; i < (int)flt.feat.size(); ++i) flt.klt_last[i] = {(S)klt_last[2 * i], (S)klt_last[2 * i + 1]};
}
template <class S> static void t_lin(Filter<S>& flt, double dt, double* F) {
    std::vector<S> f = flt.numericallyLinearizeProcess((S)dt);
    for (size_t i = 0; i < f.size(); ++i) F[i] = f[i];
}
template <class S> static void t_cbase(Filter<S>& flt, const double* in, double dt, double* out) {
    typename Filter<S>::Base b;
    for (int i = 0; i < BASE; ++i) b[i] = (S)in[i];
    auto o = flt.convolveBaseState(b, (S)dt);
    for (int i = 0; i < BASE; ++i) out[i] = o[i];
}
template <class S> static void t_cfeat(Filter<S>& flt, const double* base, const double* f3, double dt, double* out) {
    typename Filter<S>::Base b;
    for (int i = 0; i < BASE; ++i) b[i] = (S)base[i];
    typename Filter<S>::F3 f{(S)f3[0], (S)f3[1], (S)f3[2]};
    auto o = flt.convolveFeature(b, f, (S)dt);
    out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}
template <class S> static void t_noise(Filter<S>& flt, double dt, double* q) {
    auto v = flt.generateProcessNoise((S)dt);
    for (size_t i = 0; i < v.size(); ++i) q[i] = v[i];
}

extern "C" {

void* ekfo_create(int use_float, double depth, double depth_var, double uv_var) {
    Handle* h = new Handle();
    h->is_float = use_float != 0;
    Params p; p.default_point_depth = depth; p.default_point_depth_variance = depth_var; p.default_point_homogenous_variance = uv_var;
    h->d.prm = p; h->f.prm = p;
    return h;
}
void ekfo_destroy(void* hh) { delete (Handle*)hh; }
void ekfo_reset(void* hh) { Handle* h = (Handle*)hh; DISPATCH(h, flt.reset(), flt.reset()); }
int ekfo_num_features(void* hh) { Handle* h = (Handle*)hh; return h->is_float ? (int)h->f.feat.size() : (int)h->d.feat.size(); }
void ekfo_add_features(void* hh, const double* uv, int k) { Handle* h = (Handle*)hh; DISPATCH(h, t_add(flt, uv, k), t_add(flt, uv, k)); }
void ekfo_process(void* hh, double dt) { Handle* h = (Handle*)hh; DISPATCH(h, flt.process(dt), flt.process((float)dt)); }
void ekfo_update(void* hh, const double* z, const double* R, const uint8_t* pass) {
    Handle* h = (Handle*)hh; DISPATCH(h, t_update(flt, z, R, pass), t_update(flt, z, R, pass));
}
void ekfo_linearize(void* hh, double dt, double* F) { Handle* h = (Handle*)hh; DISPATCH(h, t_lin(flt, dt, F), t_lin(flt, dt, F)); }
void ekfo_get_state(void* hh, double* base_mu, double* feat, double* P, double* cache7, uint8_t* del, double* klt_last, int* status) {
    Handle* h = (Handle*)hh;
    DISPATCH(h, t_get(flt, base_mu, feat, P, cache7, del, klt_last, status), t_get(flt, base_mu, feat, P, cache7, del, klt_last, status));
}
void ekfo_set_state(void* hh, const double* base_mu, const double* feat, int n, const double* P, const double* cache7,
                    const uint8_t* del, const double* klt_last) {
    Handle* h = (Handle*)hh;
    DISPATCH(h, t_set(flt, base_mu, feat, n, P, cache7, del, klt_last), t_set(flt, base_mu, feat, n, P, cache7, del, klt_last));
}
void ekfo_convolve_base(void* hh, const double* in, double dt, double* out) { Handle* h = (Handle*)hh; DISPATCH(h, t_cbase(flt, in, dt, out), t_cbase(flt, in, dt, out)); }
void ekfo_convolve_feature(void* hh, const double* base, const double* f3, double dt, double* out) {
    Handle* h = (Handle*)hh; DISPATCH(h, t_cfeat(flt, base, f3, dt, out), t_cfeat(flt, base, f3, dt, out));
}
void ekfo_process_noise(void* hh, double dt, double* q) { Handle* h = (Handle*)hh; DISPATCH(h, t_noise(flt, dt, q), t_noise(flt, dt, q)); }
void ekfo_check_sigma(void* hh, int* neg, double* asym) { Handle* h = (Handle*)hh; DISPATCH(h, flt.checkSigma(neg, asym), flt.checkSigma(neg, asym)); }
// formFeatureMeasurementMap as a dense m x N matrix (test_ekf.cpp:51-63 known answer); returns m.
int ekfo_measurement_map(void* hh, const uint8_t* measured, double* H_out) {
    Handle* h = (Handle*)hh;
    std::vector<int> cols = h->is_float ? h->f.formFeatureMeasurementMap(measured) : h->d.formFeatureMeasurementMap(measured);
    int N = h->is_float ? h->f.dim() : h->d.dim();
    if (H_out) {
        std::fill(H_out, H_out + cols.size() * (size_t)N, 0.0);
        for (size_t r = 0; r < cols.size(); ++r) H_out[r * N + cols[r]] = 1.0;
    }
    return (int)cols.size();
}

// cv::RNG streams
void ekfo_rng_gaussian(uint64_t seed, int n, float* out) { CvRNG r(seed); for (int i = 0; i < n; ++i) out[i] = r.randn32f(); }
void ekfo_rng_uniform(uint64_t seed, int n, double a, double b, double* out) { CvRNG r(seed); for (int i = 0; i < n; ++i) out[i] = r.uniform(a, b); }

// analyzeEKFSimulation scenario; returns the number of steps (meas may be NULL to query).
int ekfo_scenario(int n, float depth_sigma, float depth_mu, const float* vel, const float* acc, const float* omega,
                  float dt, float tf, uint64_t seed, float* init_uv, float* meas, int max_steps) {
    SimScenario sc = make_scenario(n, depth_sigma, depth_mu, vel, acc, omega, dt, tf, seed);
    if (init_uv) std::copy(sc.init_uv.begin(), sc.init_uv.end(), init_uv);
    if (meas) {
        int st = std::min(sc.steps, max_steps);
        std::copy(sc.meas.begin(), sc.meas.begin() + (size_t)st * n * 2, meas);
    }
    return sc.steps;
}

int ekfo_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// CPU baseline: F independent filters, each n features, `steps` x (process(dt) + update(all
// measured)).  init_uv: F x n x 2; init_base: F x 22 (NULL = reference initial state);
// z: steps x F x n x 2; R diag value r.  Returns wall seconds; optionally writes final states.
double ekfo_batch_run(int F, int n, int steps, double dt, const double* init_uv, const double* init_base, const double* z,
                      double r, int threads, double* out_base /*F x 22 or NULL*/, double* out_P00 /*F or NULL*/) {
    std::vector<Filter<double>> flt((size_t)F);
    for (int i = 0; i < F; ++i) {
        flt[i].addNewFeatures(init_uv + (size_t)i * n * 2, n);
        if (init_base) for (int k = 0; k < BASE; ++k) flt[i].base_mu[k] = init_base[(size_t)i * BASE + k];
    }
    std::vector<double> R((size_t)n * 4, 0.0);
    for (int i = 0; i < n; ++i) { R[4 * i] = r; R[4 * i + 3] = r; }
    std::vector<uint8_t> pass((size_t)n, 1);
    auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int i = 0; i < F; ++i) {
        for (int s = 0; s < steps; ++s) {
            flt[i].process(dt);
            flt[i].updateWithFeaturePositions(z + ((size_t)s * F + i) * n * 2, R.data(), pass.data());
        }
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < F; ++i) {
        if (out_base) for (int k = 0; k < BASE; ++k) out_base[(size_t)i * BASE + k] = flt[i].base_mu[k];
        if (out_P00) out_P00[i] = flt[i].Sigma[0];
    }
    return sec;
}

// One step (process(dt) + update) of the restated algorithm evaluated in 80-bit extended precision (long double, eps 5.4e-20) from a
// state given in doubles — the yardstick for the FP64 evaluations: |FP64 oracle - this| is the rounding error the reference's step
// itself carries in FP64, which on ill-conditioned steps (cond(S) ~ 1e12) is far above 1e-9 (tests/test_gpu_ekf.py).
void ekfo_step_extended(int n, const double* base_mu, const double* feat, const double* P, const double* cache7, double dt, const double* z,
                        const double* R, const uint8_t* pass, double* out_mu, double* out_feat, double* out_P) {
    Filter<long double> flt;
    t_set(flt, base_mu, feat, n, P, cache7, nullptr, nullptr);
    flt.process((long double)dt);
    t_update(flt, z, R, pass);
    t_get(flt, out_mu, out_feat, out_P, nullptr, nullptr, nullptr, nullptr);
}

// Diagnostic companion of ekfo_batch_run: per filter the final status word (bit0 zero pivot, bit3 a negative LDL^T pivot was seen),
// the step at which bit3 was first set (-1: never), min diagonal and max |entry| of the final Sigma.
void ekfo_batch_run_diag(int F, int n, int steps, double dt, const double* init_uv, const double* z, double r, int threads,
                         int* out_status, int* out_first_neg, double* out_mindiag, double* out_maxabs) {
    std::vector<double> R((size_t)n * 4, 0.0);
    for (int i = 0; i < n; ++i) { R[4 * i] = r; R[4 * i + 3] = r; }
    std::vector<uint8_t> pass((size_t)n, 1);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int i = 0; i < F; ++i) {
        Filter<double> flt;
        flt.addNewFeatures(init_uv + (size_t)i * n * 2, n);
        int first = -1;
        for (int s = 0; s < steps; ++s) {
            flt.process(dt);
            flt.updateWithFeaturePositions(z + ((size_t)s * F + i) * n * 2, R.data(), pass.data());
            if (first < 0 && (flt.status & 8)) first = s;
        }
        const int N = BASE + 3 * n;
        double md = 1e300, ma = 0;
        for (int a = 0; a < N; ++a) { md = std::min(md, (double)flt.Sigma[(size_t)a * N + a]); for (int b = 0; b < N; ++b) ma = std::max(ma, std::fabs((double)flt.Sigma[(size_t)a * N + b])); }
        out_status[i] = flt.status; out_first_neg[i] = first; out_mindiag[i] = md; out_maxabs[i] = ma;
    }
}

}  // extern "C"
