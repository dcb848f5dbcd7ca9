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

// Plain device-pointer bundle passed by value to every EKF kernel.
struct EkfPtrs {
    double* mu; double* feat; int* nfeat; double* cache; uint8_t* dflags; double* klt_last; int* status;
    int* idx; double* y; int* m; double* K; double* W; double* L; int* asym;
    int F, nmax, Nmax, ldP, ldK, mmax;
    uint32_t flags;
    int sigma_lower;   // the input Sigma of symmetric filters is valid only up to the diagonal block of each feature row (after a lower-mode process)
    double depth, depth_var, uv_var;
    size_t gain_smem_doubles;
};

// general path (ekf_general.cu)
// lower != 0: symmetric filters get Sigma' only up to the diagonal block of each feature row (fused path; see ekf_general.cu)
cudaError_t launch_process_general(const EkfPtrs& p, const double* Pin, double* Pout, const double* dts, int mode, double* F_out, cudaStream_t st,
                                   long long* launches, int lower);
bool process_lower_capable(const EkfPtrs& p);
cudaError_t launch_mirror_lower(const EkfPtrs& p, double* P0, cudaStream_t st);
cudaError_t launch_gain_general(const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, double* Sg,
                                cudaStream_t st);
cudaError_t launch_joseph_general(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st);
size_t gain_general_smem_doubles(int mmax);
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

// large-state blocked path (ekf_large.cu)
struct LargePtrs { double* S; double* L; double* T; int mp; int nblk; int nrt_max; };
size_t large_scratch_doubles_S(int mmax);
size_t large_scratch_doubles_T(int mmax);
cudaError_t launch_update_large(const EkfPtrs& p, const LargePtrs& lp, const double* Pin, double* Pout, const double* z, const double* R,
                                const uint8_t* pass, cudaStream_t st, long long* launches, KernelTimer* timer);

// FP64 peak probe (fp64_peak.cu)
cudaError_t measure_fp64_peak(double* dmma_tflops, double* dfma_tflops);

}  // namespace ekfvio
