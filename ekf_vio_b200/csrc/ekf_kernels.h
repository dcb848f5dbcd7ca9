// Internal interface between the C-ABI layer (ekf_api.cu) and the EKF kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "timing.h"

namespace ekfvio {

// Function attributes (opt-in shared memory sizes) are per device: launchers keep their one-time set-up per device.
// (Host-side state, like the reference not meant for concurrent use from several threads.)
inline int current_device_slot() { int dev = 0; return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : 0; }

// Per-filter, per-update routing decided by ekf_chol_tiled:
//   ROUTE_SYM         Sigma and R symmetric, S positive definite and well conditioned: Sigma - Z Z'
//                     (ekf_fwd_tiled + ekf_joseph_sym in its one-phase form)
//   ROUTE_JOSEPH_SYM  Sigma and R symmetric, but the reduced form is not safe — a pivot ratio of S beyond ILLCOND_RATIO, where
//                     Sigma - Z Z' would lose the small posterior variances to cancellation against a huge prior, or an S that
//                     is not positive definite, which the signed factor S = L J L' carries through as the reference's unpivoted
//                     LDL^T does (TightlyCoupledEKF.cpp:577-580) — or EKFVIO_FLAG_LITERAL_JOSEPH asks for it: the Joseph form
//                     term by term (ekf_solve_tiled + ekf_joseph_sym in its two-phase form; lower triangle + mirror, so Sigma
//                     stays exactly symmetric)
//   ROUTE_JOSEPH_FULL asymmetric R or Sigma: the Joseph form with no symmetry assumption (ekf_solve_tiled + ekf_joseph_tiled)
// (EKFVIO_FLAG_FORCE_GENERAL_PATH bypasses all of this: the general kernels, LDL^T exactly as the oracle's.)
constexpr int ROUTE_SYM = 0, ROUTE_JOSEPH_SYM = 1, ROUTE_JOSEPH_FULL = 2;
//   ROUTE_DONE        (fused mode only) ekf_update_fused has already carried out the whole update of this filter — the block-sequential
//                     form of ROUTE_SYM with Sigma resident in registers (ekf_fused.cu); every other update kernel skips it.  Filters
//                     the fused kernel leaves alone (asymmetric Sigma or R, S not positive definite or beyond ILLCOND_RATIO) are
//                     marked -1 there and routed by ekf_chol_tiled as above.
constexpr int ROUTE_DONE = 3;
// max pivot / min pivot of the factor of S.  Healthy updates sit at 1e5 (first update: cond(S) ~ 9e5).  Measured on the config-3
// streams against the extended-precision oracle (tools/step_error_probe.py, profiles/r02_illcond_sweep.log): up to 1e9 every
// reduced update stays within the FP64 oracle's own rounding error of the step; at 1e11 the first ones do not.
constexpr double ILLCOND_RATIO = 1e9;

// Plain device-pointer bundle passed by value to every EKF kernel.
struct EkfPtrs {
    double* mu; double* feat; int* nfeat; double* cache; uint8_t* dflags; double* klt_last; int* status;
    int* idx; double* y; int* m; double* K; double* W; double* L; int* asym;
    int* route;        // per update: ROUTE_SYM / ROUTE_JOSEPH_SYM / ROUTE_JOSEPH_FULL
    int F, nmax, Nmax, ldP, ldK, mmax;
    uint32_t flags;
    int fused;         // ekf_update_fused ran first: filters with route == ROUTE_DONE are finished
    int* fb;           // fused mode: [0] number of filters ekf_update_fused left alone, [1..] their indices; nullptr otherwise
    int sigma_lower;   // the input Sigma of symmetric filters is valid only up to the diagonal block of each feature row (after a lower-mode process)
    double depth, depth_var, uv_var;
    double illcond;    // pivot-ratio threshold of the reduced update (ILLCOND_RATIO; EKFVIO_ILLCOND in the environment overrides it for experiments)
    size_t gain_smem_doubles;
};

// general path (ekf_general.cu)
// lower != 0: symmetric filters get Sigma' only up to the diagonal block of each feature row (fused path; see ekf_general.cu)
cudaError_t launch_process_general(const EkfPtrs& p, const double* Pin, double* Pout, const double* dts, int mode, double* F_out, cudaStream_t st,
                                   long long* launches, int lower);
bool process_lower_capable(const EkfPtrs& p);
// ekf_process_tiles.cu: covariance pass of process() on DMMA tiles (symmetric filters, lower mode)
bool process_tiles_capable(const EkfPtrs& p);
cudaError_t launch_process_cov_tiles(const EkfPtrs& p, const double* Pin, double* Pout, const double* dts, cudaStream_t st);
cudaError_t launch_mirror_lower(const EkfPtrs& p, double* P0, cudaStream_t st);
// only_route < 0: every filter; otherwise only the filters the Cholesky kernel routed there (p.route[f] == only_route)
cudaError_t launch_gain_general(const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, double* Sg,
                                cudaStream_t st, int only_route = -1);
cudaError_t launch_joseph_general(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st, int only_route = -1);
size_t gain_general_smem_doubles(int mmax);
cudaError_t launch_convolve_single(double* d_io, double dt, int which, int fresh, cudaStream_t st);
cudaError_t launch_reset(const EkfPtrs& p, double* P0, cudaStream_t st);
cudaError_t launch_add_features(const EkfPtrs& p, double* P0, const int* ks, const double* uv, int kmax, cudaStream_t st);
cudaError_t launch_remove_features(const EkfPtrs& p, const double* Pin, double* Pout, const uint8_t* remove, cudaStream_t st);
cudaError_t launch_check_sigma(const EkfPtrs& p, const double* P0, int* neg, double* asym, cudaStream_t st);
cudaError_t launch_fill_dt(double* dts, double dt, int F, cudaStream_t st);
cudaError_t launch_pack_P(const double* P0, double* dense, int ld, int Nmax, int F, int to_dense, cudaStream_t st);
cudaError_t launch_accumulate_errors(const EkfPtrs& p, const double* truth, double* acc, cudaStream_t st);

// tiled DMMA fast path (ekf_tiled.cu)
bool joseph_tiled_supported(const EkfPtrs& p);
bool gain_tiled_supported(const EkfPtrs& p);
size_t gain_tiled_scratch_doubles(int mmax);
cudaError_t launch_gain_tiled(int which, const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, cudaStream_t st);
cudaError_t launch_joseph_tiled(const EkfPtrs& p, const double* Pin, double* Pout, int only_asym, cudaStream_t st);
bool joseph_sym_supported(const EkfPtrs& p);
cudaError_t launch_joseph_sym(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st);

// fused sequential update, Sigma in registers (ekf_fused.cu)
bool update_fused_supported(const EkfPtrs& p);
cudaError_t launch_update_fused(const EkfPtrs& p, const double* Pin, double* Pout, const double* z, const double* R, const uint8_t* pass,
                                cudaStream_t st);

// large-state blocked path (ekf_large.cu)
struct LargePtrs { double* S; double* L; double* T; int mp; int nblk; int nrt_max; };
size_t large_scratch_doubles_S(int mmax);
size_t large_scratch_doubles_T(int mmax);
cudaError_t launch_update_large(const EkfPtrs& p, const LargePtrs& lp, const double* Pin, double* Pout, const double* z, const double* R,
                                const uint8_t* pass, cudaStream_t st, long long* launches, KernelTimer* timer);

// FP64 peak probe (fp64_peak.cu)
cudaError_t measure_fp64_peak(double* dmma_tflops, double* dfma_tflops);

}  // namespace ekfvio
