// EKFVIO::replenishFeatures (EKFVIO.cpp:224-311) for batches of frames on sm_100a:
//   cv::FAST(img, kp, FAST_THRESHOLD, true)   -> fast_score_kernel + fast_nms_mask_kernel + fast_compact_kernel
//   checkImg / cv::circle / greedy scan        -> replenish_select_kernel
// plus the C-ABI layer of the `ekfvio_fast_*` entry points (include/ekfvio_c.h).
// Integer work throughout; results are bit-identical to OpenCV's (keypoint set, order, response,
// filled-circle raster), which the tests check against cv2 golden vectors.
#include <cstring>
#include <new>
#include <string>

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ekfvio_c.h"

namespace ekfvio {
extern thread_local std::string g_last_error;
int fail(const char* what, cudaError_t e);
int fail_msg(const std::string& msg);
}  // namespace ekfvio
using ekfvio::fail_msg;

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ekfvio::fail(#x, e_); } while (0)

struct ekfvio_fast {
    int device = 0, width = 0, height = 0, max_batch = 0, max_keypoints = 0;
    int spitch = 0;                 // pitch of the score plane (multiple of 16)
    uint8_t* d_score = nullptr;     // [max_batch][height][spitch]  cornerScore + 1, 0 = not a corner
    unsigned* d_masks = nullptr;    // [max_batch][height][ceil(width/32)] keypoint bit masks
    int* d_rowcnt = nullptr;        // [max_batch][height]
    // staging for the host-buffer entry point
    uint8_t* d_img = nullptr; short* d_kp = nullptr; int* d_resp = nullptr; int* d_count = nullptr;
    float* d_exist = nullptr; int* d_nexist = nullptr; int* d_needed = nullptr; float* d_K9 = nullptr;
    short* d_new_px = nullptr; float* d_new_metric = nullptr; int* d_nnew = nullptr;
    int max_existing = 0;
    long long launches = 0;
};

namespace {

constexpr int FTW = 128, FTH = 16;           // pixels per CTA tile of the score kernel (thread = 4 x 2 pixels)
constexpr int FSW = FTW + 8, FSH = FTH + 6;  // staged tile: x0-4 .. x0+FTW+3, y0-3 .. y0+FTH+2

// The 16-pixel ring of radius 3 in OpenCV's order (fast.cpp makeOffsets, patternSize 16); constexpr so
// that the unrolled loops address the staged tile with immediate offsets.
__host__ __device__ constexpr int ring_dx(int k) {
    return k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 2 : k == 3 ? 3 : k == 4 ? 3 : k == 5 ? 3 : k == 6 ? 2 : k == 7 ? 1 : k == 8 ? 0 : k == 9 ? -1 :
           k == 10 ? -2 : k == 11 ? -3 : k == 12 ? -3 : k == 13 ? -3 : k == 14 ? -2 : -1;
}
__host__ __device__ constexpr int ring_dy(int k) {
    return k == 0 ? 3 : k == 1 ? 3 : k == 2 ? 2 : k == 3 ? 1 : k == 4 ? 0 : k == 5 ? -1 : k == 6 ? -2 : k == 7 ? -3 : k == 8 ? -3 : k == 9 ? -3 :
           k == 10 ? -2 : k == 11 ? -1 : k == 12 ? 0 : k == 13 ? 1 : k == 14 ? 2 : 3;
}

// 9 contiguous set bits in a circular 16-bit mask
__device__ __forceinline__ bool has_arc9(unsigned m) {
    unsigned M = m | (m << 16);
    unsigned c = M & (M >> 1);      // runs of 2
    c &= c >> 2;                    // runs of 4
    c &= c >> 4;                    // runs of 8
    c &= M >> 8;                    // runs of 9
    return (c & 0xffffu) != 0;
}

// FAST-9/16 corner test + cornerScore<16> (fast_score.cpp): score image = score + 1, 0 elsewhere.
// grid = (tiles_x, tiles_y, batch); 256 threads, a thread owns 4 x 2 pixels of the tile.
__global__ void __launch_bounds__(256) fast_score_kernel(const uint8_t* __restrict__ imgs, int pitch, size_t istride, int w, int h, int threshold,
                                                         uint8_t* __restrict__ score, int spitch, size_t sstride) {
    __shared__ __align__(16) uint8_t tile[FSH][FSW];
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * FTH;
    const uint8_t* img = imgs + (size_t)b * istride;
    for (int e = tid; e < FSH * (FSW / 4); e += 256) {
        const int r = e / (FSW / 4), wc = e % (FSW / 4);
        const int y = min(max(y0 - 3 + r, 0), h - 1), xs = x0 - 4 + wc * 4;
        const uint8_t* row = img + (size_t)y * pitch;
        uint32_t v;
        if (xs >= 0 && xs + 3 < w && ((pitch & 3) == 0) && ((((size_t)img) & 3) == 0)) v = *reinterpret_cast<const uint32_t*>(row + xs);
        else v = (uint32_t)row[min(max(xs, 0), w - 1)] | ((uint32_t)row[min(max(xs + 1, 0), w - 1)] << 8) |
                 ((uint32_t)row[min(max(xs + 2, 0), w - 1)] << 16) | ((uint32_t)row[min(max(xs + 3, 0), w - 1)] << 24);
        *reinterpret_cast<uint32_t*>(&tile[r][wc * 4]) = v;     // (clamped pixels only feed border outputs, which are zero)
    }
    __syncthreads();
    const int tx = (tid & 31) * 4, ty = (tid >> 5) * 2;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy) {
        const int y = y0 + ty + yy;
        if (y >= h) break;
        uint32_t out = 0;
#pragma unroll
        for (int xx = 0; xx < 4; ++xx) {
            const int x = x0 + tx + xx;
            unsigned s = 0;
            if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
                const uint8_t* c = &tile[ty + yy + 3][tx + xx + 4];
                const int v = c[0];
                // any arc of 9 contains ring pixel 0 or 8 (then 2 or 10, 4 or 12, 6 or 14): cheap rejection first
                const int p0 = c[3 * FSW], p8 = c[-3 * FSW];
                const int lo = v - threshold, hi = v + threshold;
                unsigned quick = ((p0 > hi) | (p8 > hi)) | (((p0 < lo) | (p8 < lo)) << 1);
                if (quick) {
                    const int p4 = c[3], p12 = c[-3];
                    quick &= ((p4 > hi) | (p12 > hi)) | (((p4 < lo) | (p12 < lo)) << 1);
                }
                if (quick) {
                    int d[16];
                    unsigned bright = 0, dark = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int pk = c[ring_dy(k) * FSW + ring_dx(k)];
                        d[k] = v - pk;
                        bright |= (unsigned)(pk > hi) << k;
                        dark |= (unsigned)(pk < lo) << k;
                    }
                    if (has_arc9(bright) || has_arc9(dark)) {
                        // cornerScore<16> (fast_score.cpp): the largest threshold for which the pixel is still a corner,
                        // = max(t, max over the 16 arcs of min d, max over arcs of min -d) - 1, evaluated as OpenCV does:
                        // eight ring pixels k+1..k+8 are shared by the arcs starting at k and k+1
                        int a0 = threshold;
#pragma unroll
                        for (int k = 0; k < 16; k += 2) {
                            int a = min(d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
                            for (int j = 3; j <= 8; ++j) a = min(a, d[(k + j) & 15]);
                            a0 = max(a0, min(a, d[k]));
                            a0 = max(a0, min(a, d[(k + 9) & 15]));
                        }
                        int b0 = -a0;
#pragma unroll
                        for (int k = 0; k < 16; k += 2) {
                            int bb = max(d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
                            for (int j = 3; j <= 8; ++j) bb = max(bb, d[(k + j) & 15]);
                            b0 = min(b0, max(bb, d[k]));
                            b0 = min(b0, max(bb, d[(k + 9) & 15]));
                        }
                        const int best = -b0;
                        s = (unsigned)best;         // = score + 1
                    }
                }
            }
            out |= s << (8 * xx);
        }
        const int x = x0 + tx;
        if (x < w) *reinterpret_cast<uint32_t*>(score + (size_t)b * sstride + (size_t)y * spitch + x) = out;   // spitch is a multiple of 16
    }
}

// Non-maximum suppression (a corner survives iff its score is strictly greater than its 8
// neighbours', fast.cpp) and compaction in OpenCV's order: rows top to bottom, x ascending.
// Two launches.  fast_nms_mask_kernel: a warp per row (grid over rows and images) writes a 32-bit ballot word
// per 32 pixels and the row's keypoint count.  fast_compact_kernel: one CTA per image scans the row counts and
// every warp writes its rows' keypoints at their final positions (deterministic order without atomics).
__global__ void __launch_bounds__(256) fast_nms_mask_kernel(const uint8_t* __restrict__ score, int spitch, size_t sstride, int w, int h, int nonmax,
                                                            unsigned* __restrict__ masks, int* __restrict__ row_cnt) {
    const int wpr = (w + 31) >> 5;
    const int b = blockIdx.y, lane = threadIdx.x & 31, y = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (y >= h) return;
    const uint8_t* r1 = score + (size_t)b * sstride + (size_t)y * spitch;
    unsigned* mrow = masks + ((size_t)b * h + y) * wpr;
    int cnt = 0;
    for (int wi = 0; wi < wpr; ++wi) {
        const int x = wi * 32 + lane;
        bool keep = false;
        if (x < w) {
            const int s = r1[x];
            if (s) {
                keep = true;
                if (nonmax) {              // s > 0 implies 3 <= x < w-3, 3 <= y < h-3: all neighbours exist
                    const uint8_t* r0 = r1 - spitch;
                    const uint8_t* r2 = r1 + spitch;
                    keep = s > r0[x - 1] && s > r0[x] && s > r0[x + 1] && s > r1[x - 1] && s > r1[x + 1] && s > r2[x - 1] && s > r2[x] && s > r2[x + 1];
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) mrow[wi] = m;
        cnt += __popc(m);
    }
    if (lane == 0) row_cnt[(size_t)b * h + y] = cnt;
}

__global__ void __launch_bounds__(512) fast_compact_kernel(const uint8_t* __restrict__ score, int spitch, size_t sstride, int w, int h,
                                                           const unsigned* __restrict__ masks, const int* __restrict__ row_cnt_g,
                                                           short* __restrict__ kp_xy, int* __restrict__ response, int* __restrict__ count, int max_kp) {
    extern __shared__ int row_cnt[];           // [h] -> exclusive prefix after the scan
    const int wpr = (w + 31) >> 5;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint8_t* S = score + (size_t)b * sstride;
    for (int y = threadIdx.x; y < h; y += blockDim.x) row_cnt[y] = row_cnt_g[(size_t)b * h + y];
    __syncthreads();
    if (warp == 0) {                           // exclusive scan of the row counts
        int carry = 0;
        for (int y0 = 0; y0 < h; y0 += 32) {
            const int y = y0 + lane;
            const int c = y < h ? row_cnt[y] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (y < h) row_cnt[y] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) count[b] = carry;
    }
    __syncthreads();
    short* kp = kp_xy + (size_t)b * max_kp * 2;
    int* rs = response ? response + (size_t)b * max_kp : nullptr;
    for (int y = warp; y < h; y += nw) {
        int base = row_cnt[y];
        if (base >= max_kp) continue;
        const int next = y + 1 < h ? row_cnt[y + 1] : base + 1;   // rows without keypoints need no mask reads
        if (y + 1 < h && next == base) continue;
        const unsigned* mrow = masks + ((size_t)b * h + y) * wpr;
        for (int wi = 0; wi < wpr; ++wi) {
            const unsigned m = mrow[wi];
            if (m >> lane & 1u) {
                const int idx = base + __popc(m & ((1u << lane) - 1u));
                if (idx < max_kp) {
                    const int x = wi * 32 + lane;
                    kp[idx * 2] = (short)x; kp[idx * 2 + 1] = (short)y;
                    if (rs) rs[idx] = (int)S[(size_t)y * spitch + x] - 1;
                }
            }
            base += __popc(m);
        }
    }
}

// Half-widths of the rows of cv::circle(..., radius, ..., thickness = -1): OpenCV's midpoint
// rasteriser (drawing.cpp, Circle()) emits spans (cy -+ dy, cx -+ dx) and (cy -+ dx, cx -+ dy);
// the filled circle is their union, i.e. per row offset the widest span.
__device__ void circle_half_widths(int radius, int* hw /*[2*radius+1]*/) {
    for (int i = 0; i <= 2 * radius; ++i) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        hw[radius - dy] = max(hw[radius - dy], dx); hw[radius + dy] = max(hw[radius + dy], dx);
        hw[radius - dx] = max(hw[radius - dx], dy); hw[radius + dx] = max(hw[radius + dx], dy);
        ++dy;
        err += plus; plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// The greedy scan of EKFVIO.cpp:252-305, one warp per frame.  The check image is a bit mask in shared
// memory (one word per 32 pixels); a circle is drawn with the lanes along its rows, each lane OR-ing
// the span of its row into the two or three words it touches.
__global__ void __launch_bounds__(32) replenish_select_kernel(const short* __restrict__ kp_xy, const int* __restrict__ count, int max_kp,
                                                              const float* __restrict__ existing_px, const int* __restrict__ n_existing,
                                                              int max_existing, const int* __restrict__ needed_in, int radius, int kill_pad,
                                                              const float* __restrict__ K9, int w, int h, short* __restrict__ new_px,
                                                              float* __restrict__ new_metric, int* __restrict__ n_new, int max_new) {
    extern __shared__ int sm_sel[];
    int* hw = sm_sel;                          // [2*radius+1]
    const int wpr = (w + 31) >> 5;
    unsigned* bits = reinterpret_cast<unsigned*>(sm_sel + 2 * radius + 1);   // [h][wpr]
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < h * wpr; i += 32) bits[i] = 0u;
    if (lane == 0) circle_half_widths(radius, hw);
    __syncwarp();
    auto draw = [&](int cx, int cy) {
        for (int i = lane; i <= 2 * radius; i += 32) {
            const int y = cy - radius + i, half = hw[i];
            if (half < 0 || y < 0 || y >= h) continue;
            const int xa = max(cx - half, 0), xb = min(cx + half, w - 1);
            if (xa > xb) continue;
            const int w0 = xa >> 5, w1 = xb >> 5;
            for (int ww = w0; ww <= w1; ++ww) {
                const unsigned from = ww == w0 ? (xa & 31) : 0, to = ww == w1 ? (xb & 31) : 31;
                const unsigned m = (to == 31 ? 0xffffffffu : ((1u << (to + 1)) - 1u)) & ~((1u << from) - 1u);
                bits[y * wpr + ww] |= m;
            }
        }
        __syncwarp();
    };
    const int ne = n_existing ? n_existing[b] : 0;
    for (int e = 0; e < ne; ++e) {             // Feature::getPixel -> cv::Point: cvRound, half to even
        const float ex = existing_px[((size_t)b * max_existing + e) * 2], ey = existing_px[((size_t)b * max_existing + e) * 2 + 1];
        draw(__float2int_rn(ex), __float2int_rn(ey));
    }
    const int nk = min(count[b], max_kp);
    int needed = needed_in[b], accepted = 0;
    const short* kp = kp_xy + (size_t)b * max_kp * 2;
    const float* K = K9 ? K9 + (size_t)b * 9 : nullptr;
    for (int i = 0; i < needed && i < nk; ++i) {
        const int x = kp[i * 2], y = kp[i * 2 + 1];
        if (bits[y * wpr + (x >> 5)] >> (x & 31) & 1u) { ++needed; continue; }                     // too close to a feature (:282-286)
        if (x < kill_pad || y < kill_pad || w - x < kill_pad || h - y < kill_pad) { ++needed; continue; }   // Frame::isPixelInBox (:290-295)
        __syncwarp();
        draw(x, y);
        if (accepted < max_new && lane == 0) {
            new_px[((size_t)b * max_new + accepted) * 2] = (short)x; new_px[((size_t)b * max_new + accepted) * 2 + 1] = (short)y;
            if (new_metric && K) {             // Feature::pixel2Metric with the linear-index K (E1)
                new_metric[((size_t)b * max_new + accepted) * 2] = ((float)x - K[2]) / K[0];
                new_metric[((size_t)b * max_new + accepted) * 2 + 1] = ((float)y - K[5]) / K[4];
            }
        }
        ++accepted;
    }
    if (lane == 0) n_new[b] = min(accepted, max_new);
}

// Frame::Frame (Frame.cpp:15-21): cv::resize(img, scaled, Size(cols / inv_scale, rows / inv_scale)) with the
// default INTER_LINEAR on 8-bit images.  OpenCV's arithmetic: inv_scale 1 copies; 2 takes the
// "area fast" path (a+b+c+d+2)>>2; otherwise 11-bit fixed-point coefficients per output column / row
// (float source coordinate (d+0.5)*scale-0.5 computed in double, narrowed to float, floor, weights
// rounded to 1/2048), horizontal pass in int, vertical pass ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2.
__device__ __forceinline__ void resize_coeff(int d, double scale, int sn, int& s0, int& a0, int& a1) {
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);    // no FMA contraction: OpenCV rounds the product first
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    s0 = s;
    a0 = __float2int_rn(__fmul_rn(1.f - f, 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256) frame_resize_kernel(const uint8_t* __restrict__ src, int spitch, size_t sstride, int sw, int sh,
                                                           uint8_t* __restrict__ dst, int dpitch, size_t dstride, int dw, int dh, int inv_scale,
                                                           double scale_x, double scale_y) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
    if (x >= dw || y >= dh) return;
    const uint8_t* S = src + (size_t)b * sstride;
    int out;
    if (inv_scale == 1) {
        out = S[(size_t)y * spitch + x];
    } else if (inv_scale == 2) {
        const uint8_t* r0 = S + (size_t)(2 * y) * spitch + 2 * x;
        const uint8_t* r1 = r0 + spitch;
        out = (r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2;
    } else {
        int sx, ax0, ax1, sy, ay0, ay1;
        resize_coeff(x, scale_x, sw, sx, ax0, ax1);
        resize_coeff(y, scale_y, sh, sy, ay0, ay1);
        const int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);
        const uint8_t* r0 = S + (size_t)sy * spitch;
        const uint8_t* r1 = S + (size_t)sy1 * spitch;
        const int h0 = r0[sx] * ax0 + r0[sx1] * ax1, h1 = r1[sx] * ax0 + r1[sx1] * ax1;
        out = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out = min(max(out, 0), 255);
    }
    dst[(size_t)b * dstride + (size_t)y * dpitch + x] = (uint8_t)out;
}

}  // namespace

extern "C" {

int ekfvio_frame_resize(const uint8_t* d_src, int src_pitch, int src_width, int src_height, int batch, int inv_scale, uint8_t* d_dst,
                        int dst_pitch, void* stream) {
    if (!d_src || !d_dst || src_width <= 0 || src_height <= 0 || batch <= 0 || inv_scale <= 0) return fail_msg("ekfvio_frame_resize: bad arguments");
    const int dw = src_width / inv_scale, dh = src_height / inv_scale;
    if (dw <= 0 || dh <= 0 || src_pitch < src_width || dst_pitch < dw) return fail_msg("ekfvio_frame_resize: bad sizes");
    // resize.cpp: inv_scale_x = dsize.width / ssize.width; scale_x = 1. / inv_scale_x
    const double scale_x = 1.0 / ((double)dw / src_width), scale_y = 1.0 / ((double)dh / src_height);
    // the area-fast substitution needs exact 2x in both directions (resize.cpp: is_area_fast && iscale == 2)
    int mode = inv_scale;
    if (inv_scale == 2 && (dw * 2 != src_width || dh * 2 != src_height)) mode = 3 /* any value > 2: general path */;
    dim3 grid((dw + 31) / 32, (dh + 7) / 8, batch);
    frame_resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_pitch, (size_t)src_pitch * src_height, src_width, src_height, d_dst,
                                                               dst_pitch, (size_t)dst_pitch * dh, dw, dh, mode, scale_x, scale_y);
    CU(cudaGetLastError());
    return 0;
}


int ekfvio_fast_destroy(ekfvio_fast* f) {
    if (!f) return 0;
    cudaSetDevice(f->device);
    cudaFree(f->d_score); cudaFree(f->d_masks); cudaFree(f->d_rowcnt); cudaFree(f->d_img); cudaFree(f->d_kp); cudaFree(f->d_resp); cudaFree(f->d_count);
    cudaFree(f->d_exist); cudaFree(f->d_nexist); cudaFree(f->d_needed); cudaFree(f->d_K9); cudaFree(f->d_new_px); cudaFree(f->d_new_metric);
    cudaFree(f->d_nnew);
    delete f;
    return 0;
}

int ekfvio_fast_create(ekfvio_fast** out, int device, int width, int height, int max_batch, int max_keypoints) {
    if (!out || width < 7 || height < 7 || max_batch <= 0 || max_keypoints <= 0) return fail_msg("ekfvio_fast_create: bad arguments");
    if (width > 32767 || height > 32767) return fail_msg("ekfvio_fast_create: image larger than 32767 pixels on a side");
    const size_t nms_smem = (size_t)height * sizeof(int);
    if (nms_smem > 200 * 1024) return fail_msg("ekfvio_fast_create: image too tall for the row-count scan");
    if ((size_t)(height * ((width + 31) / 32) + 2 * 1024 + 1) * sizeof(int) > 200 * 1024) return fail_msg("ekfvio_fast_create: image too large for the shared-memory check image of the greedy scan (height * width / 8 bytes must be <= 190 KB)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("ekfvio_fast_create: no CUDA device (this library has no CPU path)");
    CU(cudaSetDevice(device));
    ekfvio_fast* f = new (std::nothrow) ekfvio_fast;
    if (!f) return fail_msg("out of host memory");
    f->device = device; f->width = width; f->height = height; f->max_batch = max_batch; f->max_keypoints = max_keypoints;
    f->spitch = (width + 15) / 16 * 16;
    f->max_existing = 512;
    const size_t plane = (size_t)f->spitch * height * max_batch;
    cudaError_t e = cudaMalloc((void**)&f->d_score, plane);
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_img, plane);
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_masks, (size_t)max_batch * height * ((width + 31) / 32) * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_rowcnt, (size_t)max_batch * height * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_kp, (size_t)max_batch * max_keypoints * 2 * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_resp, (size_t)max_batch * max_keypoints * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_count, (size_t)max_batch * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_exist, (size_t)max_batch * f->max_existing * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_nexist, (size_t)max_batch * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_needed, (size_t)max_batch * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_K9, (size_t)max_batch * 9 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_new_px, (size_t)max_batch * f->max_existing * 2 * sizeof(short));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_new_metric, (size_t)max_batch * f->max_existing * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_nnew, (size_t)max_batch * sizeof(int));
    if (e == cudaSuccess && nms_smem > 48 * 1024) e = cudaFuncSetAttribute(fast_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem);
    if (e != cudaSuccess) { ekfvio_fast_destroy(f); return ekfvio::fail("ekfvio_fast_create", e); }
    *out = f;
    return 0;
}

int ekfvio_fast_detect(ekfvio_fast* f, const uint8_t* d_imgs, int pitch, int batch, int threshold, int nonmax, short* d_kp_xy, int* d_response,
                       int* d_count, void* stream) {
    if (!f || !d_imgs || !d_kp_xy || !d_count) return fail_msg("ekfvio_fast_detect: null argument");
    if (batch <= 0 || batch > f->max_batch || pitch < f->width) return fail_msg("ekfvio_fast_detect: bad batch or pitch");
    if (threshold < 0 || threshold > 254) return fail_msg("ekfvio_fast_detect: threshold must be within 0..254");
    CU(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sstride = (size_t)f->spitch * f->height;
    dim3 grid((f->width + FTW - 1) / FTW, (f->height + FTH - 1) / FTH, batch);
    fast_score_kernel<<<grid, 256, 0, st>>>(d_imgs, pitch, (size_t)pitch * f->height, f->width, f->height, threshold, f->d_score, f->spitch, sstride);
    CU(cudaGetLastError());
    fast_nms_mask_kernel<<<dim3((f->height + 7) / 8, batch), 256, 0, st>>>(f->d_score, f->spitch, sstride, f->width, f->height, nonmax, f->d_masks,
                                                                           f->d_rowcnt);
    CU(cudaGetLastError());
    fast_compact_kernel<<<batch, 512, (size_t)f->height * sizeof(int), st>>>(f->d_score, f->spitch, sstride, f->width, f->height, f->d_masks, f->d_rowcnt,
                                                                             d_kp_xy, d_response, d_count, f->max_keypoints);
    CU(cudaGetLastError());
    f->launches += 3;
    return 0;
}

int ekfvio_fast_select(ekfvio_fast* f, const short* d_kp_xy, const int* d_count, const float* d_existing_px, const int* d_n_existing,
                       int max_existing, const int* d_needed, int min_dist, int kill_pad, const float* d_K9, short* d_new_px,
                       float* d_new_metric, int* d_n_new, int max_new, int batch, void* stream) {
    if (!f || !d_kp_xy || !d_count || !d_needed || !d_new_px || !d_n_new) return fail_msg("ekfvio_fast_select: null argument");
    if (batch <= 0 || batch > f->max_batch || min_dist < 0 || min_dist > 1024 || max_new <= 0) return fail_msg("ekfvio_fast_select: bad arguments");
    CU(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sm = ((size_t)(2 * min_dist + 1) + (size_t)f->height * ((f->width + 31) / 32)) * sizeof(int);
    if (sm > 200 * 1024) return fail_msg("ekfvio_fast_select: check image does not fit in shared memory");
    if (sm > 48 * 1024) CU(cudaFuncSetAttribute(replenish_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    replenish_select_kernel<<<batch, 32, sm, st>>>(d_kp_xy, d_count, f->max_keypoints, d_existing_px, d_n_existing, max_existing, d_needed, min_dist,
                                                   kill_pad, d_K9, f->width, f->height, d_new_px, d_new_metric, d_n_new, max_new);
    CU(cudaGetLastError());
    f->launches += 1;
    return 0;
}

int ekfvio_fast_replenish_h(ekfvio_fast* f, const uint8_t* h_imgs, int pitch, int batch, int threshold, const float* h_existing_px,
                            const int* h_n_existing, int max_existing, const int* h_needed, int min_dist, int kill_pad, const float* h_K9,
                            short* h_new_px, float* h_new_metric, int* h_n_new, int max_new, short* h_kp_xy, int* h_count, void* stream) {
    if (!f || !h_imgs || !h_needed || !h_new_px || !h_n_new) return fail_msg("ekfvio_fast_replenish_h: null argument");
    if (batch <= 0 || batch > f->max_batch || pitch < f->width || pitch > f->spitch) return fail_msg("ekfvio_fast_replenish_h: bad batch or pitch");
    if (max_existing < 0 || max_existing > f->max_existing || max_new <= 0 || max_new > f->max_existing)
        return fail_msg("ekfvio_fast_replenish_h: max_existing / max_new above the handle's capacity (512)");
    CU(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(f->d_img, h_imgs, (size_t)pitch * f->height * batch, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(f->d_needed, h_needed, batch * sizeof(int), cudaMemcpyHostToDevice, st));
    const bool have_exist = h_existing_px && h_n_existing && max_existing > 0;
    if (have_exist) {
        CU(cudaMemcpyAsync(f->d_exist, h_existing_px, (size_t)batch * max_existing * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(f->d_nexist, h_n_existing, batch * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (h_K9) CU(cudaMemcpyAsync(f->d_K9, h_K9, (size_t)batch * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = ekfvio_fast_detect(f, f->d_img, pitch, batch, threshold, 1, f->d_kp, nullptr, f->d_count, stream);
    if (rc) return rc;
    rc = ekfvio_fast_select(f, f->d_kp, f->d_count, have_exist ? f->d_exist : nullptr, have_exist ? f->d_nexist : nullptr, max_existing, f->d_needed,
                            min_dist, kill_pad, h_K9 ? f->d_K9 : nullptr, f->d_new_px, h_new_metric && h_K9 ? f->d_new_metric : nullptr, f->d_nnew,
                            max_new, batch, stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_new_px, f->d_new_px, (size_t)batch * max_new * 2 * sizeof(short), cudaMemcpyDeviceToHost, st));
    if (h_new_metric && h_K9) CU(cudaMemcpyAsync(h_new_metric, f->d_new_metric, (size_t)batch * max_new * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_n_new, f->d_nnew, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_kp_xy) CU(cudaMemcpyAsync(h_kp_xy, f->d_kp, (size_t)batch * f->max_keypoints * 2 * sizeof(short), cudaMemcpyDeviceToHost, st));
    if (h_count) CU(cudaMemcpyAsync(h_count, f->d_count, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int ekfvio_frame_resize_h(const uint8_t* h_src, int src_pitch, int src_width, int src_height, int batch, int inv_scale, uint8_t* h_dst,
                          int dst_pitch) {
    if (!h_src || !h_dst || src_width <= 0 || src_height <= 0 || batch <= 0 || inv_scale <= 0) return fail_msg("ekfvio_frame_resize_h: bad arguments");
    const int dw = src_width / inv_scale, dh = src_height / inv_scale;
    if (dw <= 0 || dh <= 0 || src_pitch < src_width || dst_pitch < dw) return fail_msg("ekfvio_frame_resize_h: bad sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("ekfvio_frame_resize_h: no CUDA device (this library has no CPU path)");
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    const size_t sb = (size_t)src_pitch * src_height * batch, db = (size_t)dst_pitch * dh * batch;
    cudaError_t e = cudaMalloc((void**)&d_src, sb);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_dst, db);
    if (e == cudaSuccess) e = cudaMemcpy(d_src, h_src, sb, cudaMemcpyHostToDevice);
    int rc = 0;
    if (e == cudaSuccess) rc = ekfvio_frame_resize(d_src, src_pitch, src_width, src_height, batch, inv_scale, d_dst, dst_pitch, nullptr);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_dst, d_dst, db, cudaMemcpyDeviceToHost);
    cudaFree(d_src); cudaFree(d_dst);
    if (e != cudaSuccess) return ekfvio::fail("ekfvio_frame_resize_h", e);
    return rc;
}

long long ekfvio_fast_launch_count(const ekfvio_fast* f) { return f ? f->launches : 0; }

}  // extern "C"
