// Internal interface between klt_api.cu and klt_kernels.cu.
#pragma once
#include <cuda.h>

#include "klt_common.cuh"

namespace kltdev {
cudaError_t launch_level(const LevelJob& j0, const LevelJob& j1, int w, int h, int cpitch, size_t cstride, int dpitch, size_t dstride,
                         int npitch, size_t nstride, cudaStream_t st);
// TMA-staged pyramid kernels (klt_kernels.cu): level 0 in 256 x 64 tiles, levels 1..3 of an image fused in one CTA
struct TmaJob {
    uint8_t* copy_dst; short2* deriv; uint8_t* down;
    int batch, first;                              // images [first, first + batch) of the tensor map's third dimension
};
struct FusedLevel { int w, h, pitch, dpitch, spitch, srows, soff; size_t img_off, der_off, img_stride, der_stride; };   // spitch / srows / soff: staged copy in shared memory
struct FusedJob { uint8_t* slot; int want_deriv; int batch, first; };
cudaError_t launch_level0_tma(const CUtensorMap& m0, const CUtensorMap& m1, const TmaJob& j0, const TmaJob& j1, int w, int h, int cpitch, size_t cstride,
                              int dpitch, size_t dstride, int npitch, size_t nstride, cudaStream_t st);
cudaError_t launch_levels_fused(const CUtensorMap& m0, const CUtensorMap& m1, const FusedJob& j0, const FusedJob& j1, const FusedLevel* lv, int nl,
                                size_t smem, cudaStream_t st);
cudaError_t launch_sample_uncertainty(const uint8_t* ref_imgs, const uint8_t* cur_imgs, int w, int h, int pitch, size_t stride, int batch,
                                      const float* ref_pts, const float* pts, const int* npts, int max_points, float* cov, cudaStream_t st);
size_t track_smem_bytes(int win);
cudaError_t launch_track(const Pyr& pyr, const uint8_t* prev_slot, const uint8_t* next_slot, const float* prev_pts, float* next_pts,
                         uint8_t* status, float* err, const int* npts, int max_points, int first_image, int batch, const ekfvio_klt_params& prm,
                         cudaStream_t st, ExtLevel0 prev_ext = ExtLevel0{nullptr, 0, 0}, ExtLevel0 next_ext = ExtLevel0{nullptr, 0, 0});
cudaError_t launch_postprocess(const float* next_pts, const uint8_t* status, const int* npts, const float* K9, int max_points, int batch,
                               int cols, int rows, int kill_pad, float* measured, float* cov, uint8_t* passed, cudaStream_t st);
}  // namespace kltdev
