// Internal interface between klt_api.cu and klt_kernels.cu.
#pragma once
#include "klt_common.cuh"

namespace kltdev {
cudaError_t launch_level(const LevelJob& j0, const LevelJob& j1, int w, int h, int cpitch, size_t cstride, int dpitch, size_t dstride,
                         int npitch, size_t nstride, cudaStream_t st);
size_t track_smem_bytes(int win);
cudaError_t launch_track(const Pyr& pyr, const uint8_t* prev_slot, const uint8_t* next_slot, const float* prev_pts, float* next_pts,
                         uint8_t* status, float* err, const int* npts, int max_points, int first_image, int batch, const ekfvio_klt_params& prm,
                         cudaStream_t st);
cudaError_t launch_postprocess(const float* next_pts, const uint8_t* status, const int* npts, const float* K9, int max_points, int batch,
                               int cols, int rows, int kill_pad, float* measured, float* cov, uint8_t* passed, cudaStream_t st);
}  // namespace kltdev
