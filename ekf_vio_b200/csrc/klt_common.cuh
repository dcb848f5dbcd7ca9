// Shared definitions for the KLT kernels (sm_100a).  Arithmetic follows OpenCV's
// calcOpticalFlowPyrLK operation by operation, which is
// what KLTTracker::findNewFeaturePositionsOpenCV calls (reference KLTTracker.cpp:61-64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ekfvio_c.h"
#include "timing.h"

#define KLT_MAX_LEVELS 8

namespace kltdev {

struct Level {
    int w, h;
    int pitch;        // bytes per row of the u8 image (multiple of 16)
    int dpitch;       // short2 elements per row of the derivative image (multiple of 4)
    size_t img_off;   // byte offset of image 0 of this level inside the slot
    size_t der_off;   // byte offset of derivative image 0
    size_t img_stride;  // bytes between consecutive images of the batch
    size_t der_stride;
};

// One slot's share of a pyramid-level launch.
struct LevelJob {
    const uint8_t* src; int spitch; size_t sstride;
    uint8_t* copy_dst; short2* deriv; uint8_t* down;
    int batch;
};

// Level 0 of a slot may live in the caller's image batch instead of the slot (ekfvio_klt_build_pyramid_pair_ref): img == nullptr
// means "in the slot".
struct ExtLevel0 { const uint8_t* img; int pitch; size_t stride; };

struct Pyr {
    int levels;  // number of levels (effective max level + 1)
    Level lv[KLT_MAX_LEVELS];
};

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
__host__ __device__ __forceinline__ int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

}  // namespace kltdev

struct ekfvio_klt {
    int device = 0;
    int width = 0, height = 0, max_batch = 0, max_points = 0, num_slots = 0;
    ekfvio_klt_params prm{};
    kltdev::Pyr pyr{};
    size_t slot_bytes = 0;
    uint8_t* d_slots = nullptr;       // num_slots * slot_bytes
    bool* slot_has_derivs = nullptr;  // host
    kltdev::ExtLevel0* slot_ext = nullptr;   // host, [num_slots]: level 0 by reference (see ExtLevel0)
    int* slot_batch = nullptr;        // host
    // staging for the *_h entry point
    float* d_prev_pts = nullptr; float* d_next_pts = nullptr; uint8_t* d_status = nullptr; float* d_err = nullptr; int* d_npts = nullptr;
    uint8_t* h_img = nullptr;         // pinned: 2 * max_batch * height * level0 pitch
    float* h_pts = nullptr;           // pinned: max_batch*max_points*(2+2+1) floats + status bytes
    long long launches = 0;
    KernelTimer timer;
    int seq_prev = -1;                // slot that holds the latest frame of a sequence (ekfvio_klt_track_pair_h / _track_next_h)
    // TMA descriptors (CUtensorMap, 128 bytes each, kept as raw words so that this header needs no <cuda.h>): per slot the
    // level-0 and level-1 images as 3-D tensors of 32-bit words (x/4, y, image); see klt_api.cu make_tensor_map
    unsigned long long (*tmap_l0)[16] = nullptr;   // [num_slots]
    unsigned long long (*tmap_l1)[16] = nullptr;   // [num_slots]
    bool tma_ok = false;                           // driver entry point found and the slot levels are TMA-eligible
    int fused_levels = 0;                          // levels 1 .. fused_levels are built by klt_levels_fused_kernel (0: per-level launches)
    size_t fused_smem = 0;
    cudaStream_t copy_st = nullptr;   // uploads of the host-buffer entry point, chunk by chunk
    cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};
};
