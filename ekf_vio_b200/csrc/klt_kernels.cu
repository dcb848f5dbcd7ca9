// KLT kernels for sm_100a: pyramid + Scharr build (HBM-bound streaming) and the per-feature
// Lucas-Kanade solve (one warp per feature, all pyramid levels in one launch).
// Arithmetic is OpenCV's calcOpticalFlowPyrLK (what the reference calls at KLTTracker.cpp:61-64):
// integer pyrDown/Scharr, Q14 bilinear weights, int16 patches, exact integer window sums,
// FP32 2x2 solve.  Compiled with --fmad=false so FP32 expressions round as OpenCV's do.
#include <float.h>
#include <stdlib.h>
#include <limits.h>

#include <cuda.h>

#include "klt_common.cuh"
#include "klt_kernels.h"

using namespace kltdev;

namespace {

__device__ __forceinline__ int dp4a_us(unsigned a_u8x4, int b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp4a_uu(unsigned a_u8x4, unsigned b_u8x4, unsigned c) {
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_u8x4), "r"(c));
    return d;
}

constexpr int TW = 64, TH = 64;            // input pixels per CTA tile
constexpr int SW = TW + 16, SH = TH + 4;   // staged tile: x0-8 .. x0+TW+7 (8-byte aligned rows), y0-2 .. y0+TH+1
constexpr int RPT = 8;                     // rows per thread
constexpr int LT = 8 * (TH / RPT);         // threads per CTA: 8 column groups (8 pixels) x TH/RPT row groups

// (rows further than one reflection away are only ever staged for threads that are outside the image; clamp them)
__device__ __forceinline__ int reflect_once(int p, int len) { p = p < 0 ? -p : (p >= len ? 2 * len - 2 - p : p); return max(0, min(p, len - 1)); }

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

// Stage one TW x TH tile (+ halo) with REFLECT_101 applied to image coordinates (one reflection
// suffices: every level is wider than the 21-pixel window).  Only pixels x0-2 .. x0+TW are ever
// used: the TW interior columns travel as 8-byte cp.async vectors, the two halo words (x0-4..,
// x0+TW..) separately; tiles that overhang the right edge or unaligned sources take the word path.
__device__ __forceinline__ void stage_tile(uint8_t (*tile)[SW], const uint8_t* __restrict__ s, int spitch, int w, int h, int x0, int y0, int tid,
                                           bool aligned8) {
    if (aligned8 && x0 + TW <= w) {
        const int v = tid & 7;
        for (int r = tid >> 3; r < SH; r += LT / 8) {
            const int sy = reflect_once(y0 - 2 + r, h);
            cp_async8(&tile[r][8 + v * 8], s + (size_t)sy * spitch + x0 + v * 8);
        }
        for (int e = tid; e < 2 * SH; e += LT) {
            const int r = e >> 1, right = e & 1;
            const int sy = reflect_once(y0 - 2 + r, h), xs = right ? x0 + TW : x0 - 4;
            const uint8_t* row = s + (size_t)sy * spitch;
            uint8_t* dst = &tile[r][right ? 8 + TW : 4];
            if (xs >= 0 && xs + 3 < w) cp_async4(dst, row + xs);
            else
                *reinterpret_cast<uint32_t*>(dst) = (uint32_t)row[reflect101(xs, w)] | ((uint32_t)row[reflect101(xs + 1, w)] << 8) |
                                                    ((uint32_t)row[reflect101(xs + 2, w)] << 16) | ((uint32_t)row[reflect101(xs + 3, w)] << 24);
        }
    } else {
        const bool aligned4 = ((spitch & 3) == 0) && ((((size_t)s) & 3) == 0);
        constexpr int WPR = TW / 4 + 2;        // words per staged row: columns 4 .. TW+11
        for (int e = tid; e < SH * WPR; e += LT) {
            const int r = e / WPR, wc = e % WPR + 1;
            const int sy = reflect_once(y0 - 2 + r, h), xs = x0 - 8 + wc * 4;
            const uint8_t* row = s + (size_t)sy * spitch;
            uint32_t v4;
            if (aligned4 && xs >= 0 && xs + 3 < w) v4 = *reinterpret_cast<const uint32_t*>(row + xs);
            else v4 = (uint32_t)row[reflect101(xs, w)] | ((uint32_t)row[reflect101(xs + 1, w)] << 8) | ((uint32_t)row[reflect101(xs + 2, w)] << 16) |
                      ((uint32_t)row[reflect101(xs + 3, w)] << 24);
            *reinterpret_cast<uint32_t*>(&tile[r][wc * 4]) = v4;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// One pyramid level: reads level l of every image once and writes (a) an optional copy into the
// slot (level 0 fed from a caller buffer), (b) the interleaved int16 Scharr derivatives,
// (c) level l+1 = pyrDown(level l).  grid = (tiles_x, y chunks, images of both jobs), LT threads.
// A CTA walks down its chunk of a tile column with the next tile's cp.async staging in flight while
// it computes the current one.  A thread owns an 8 x RPT pixel block of the tile and walks down the
// RPT + 3 staged rows it needs once: each row is three shared-memory loads; the Scharr 3x3 and the
// pyrDown 5-tap rows are u8x4 dot products (IDP4A) on those words, the Scharr windows of the last
// three rows stay in registers, and every global store is a full 8/16-byte vector.  The kernel is
// meant to be bound by HBM, not by issue.
__global__ void __launch_bounds__(LT, 16) klt_level_kernel(LevelJob j0, LevelJob j1, int w, int h, int cpitch, size_t cstride, int dpitch,
                                                       size_t dstride, int npitch, size_t nstride, int tiles_per_cta) {
    __shared__ __align__(16) uint8_t tiles[2][SH][SW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    int b = blockIdx.z;
    const bool second = b >= j0.batch;
    if (second) b -= j0.batch;
    const LevelJob& J = second ? j1 : j0;      // two slots (e.g. the previous and the next frame) share one launch
    uint8_t* __restrict__ copy_dst = J.copy_dst;
    short2* __restrict__ deriv = J.deriv;
    uint8_t* __restrict__ down = J.down;
    if (!copy_dst && !deriv && !down) return;
    const int spitch = J.spitch;
    const uint8_t* __restrict__ s = J.src + (size_t)b * J.sstride;
    const bool aligned8 = ((spitch & 7) == 0) && ((((size_t)s) & 7) == 0);
    const int ty_begin = blockIdx.y * tiles_per_cta;
    const int ty_end = min(ty_begin + tiles_per_cta, (h + TH - 1) / TH);
    if (ty_begin >= ty_end) return;

    stage_tile(tiles[0], s, spitch, w, h, x0, ty_begin * TH, tid, aligned8);
    for (int ty = ty_begin; ty < ty_end; ++ty) {
    uint8_t (*tile)[SW] = tiles[(ty - ty_begin) & 1];
    const int y0 = ty * TH;
    if (ty + 1 < ty_end) {
        stage_tile(tiles[(ty + 1 - ty_begin) & 1], s, spitch, w, h, x0, y0 + TH, tid, aligned8);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int cg = tid & 7, rg = tid >> 3;
    const int xg = cg * 8, x = x0 + xg;        // first of the thread's eight columns
    const int yb = y0 + rg * RPT;              // first of its RPT rows
    if (x < w && yb < h) {                     // (otherwise its pyrDown outputs are outside level l+1 as well)
    const bool full = x + 7 < w;
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    const int gx = x / 2, gy = yb / 2;

    unsigned win[3][8];                        // Scharr windows (x-1, x, x+1, x+2) of the last three rows
    unsigned acc[RPT / 2][4];                  // pyrDown sums of output rows gy .. gy + RPT/2 - 1
#pragma unroll
    for (int o = 0; o < RPT / 2; ++o)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[o][k] = 0u;
#pragma unroll
    for (int i = 0; i < RPT + 3; ++i) {        // staged row rg*RPT + i  <->  image row yb - 2 + i
        const uint8_t* tr = &tile[rg * RPT + i][xg];
        const unsigned w0 = *reinterpret_cast<const unsigned*>(tr + 4);     // pixels x-4 .. x-1
        const uint2 w12 = *reinterpret_cast<const uint2*>(tr + 8);          // pixels x .. x+7
        const unsigned w3 = *reinterpret_cast<const unsigned*>(tr + 16);    // pixels x+8 .. x+11
        const unsigned w1 = w12.x, w2 = w12.y;
        if (copy_dst && i >= 2 && i < RPT + 2) {   // the slot's pitch is a multiple of 16: whole vectors stay inside the row
            const int y = yb + i - 2;
            if (y < h) *reinterpret_cast<uint2*>(copy_dst + (size_t)b * cstride + (size_t)y * cpitch + x) = w12;
        }
        if (down) {
            // [1 4 6 4 1] across for the outputs centred on pixels x, x+2, x+4, x+6: two dot products
            // each on the aligned words, then the same taps down the rows (output row gy + o is
            // centred on staged row 2o + 2)
            unsigned hz[4];
            hz[0] = dp4a_uu(w1, 0x00010406u, dp4a_uu(w0, 0x04010000u, 0u));
            hz[1] = dp4a_uu(w2, 0x00000001u, dp4a_uu(w1, 0x04060401u, 0u));
            hz[2] = dp4a_uu(w2, 0x00010406u, dp4a_uu(w1, 0x04010000u, 0u));
            hz[3] = dp4a_uu(w3, 0x00000001u, dp4a_uu(w2, 0x04060401u, 0u));
#pragma unroll
            for (int o = 0; o < RPT / 2; ++o) {
                const int d = i - (2 * o + 2);
                const unsigned wv = (d == 0) ? 6u : (d == 1 || d == -1) ? 4u : (d == 2 || d == -2) ? 1u : 0u;
                if (wv) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[o][k] += wv * hz[k];
                }
                if (d == 2 && gy + o < dh) {   // row complete: round, pack, store
                    const unsigned r0 = (acc[o][0] + 128u) >> 8, r1 = (acc[o][1] + 128u) >> 8, r2 = (acc[o][2] + 128u) >> 8, r3 = (acc[o][3] + 128u) >> 8;
                    uint8_t* nd = down + (size_t)b * nstride + (size_t)(gy + o) * npitch + gx;
                    if (gx + 3 < dw) *reinterpret_cast<unsigned*>(nd) = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
                    else {
                        nd[0] = (uint8_t)r0;
                        if (gx + 1 < dw) nd[1] = (uint8_t)r1;
                        if (gx + 2 < dw) nd[2] = (uint8_t)r2;
                    }
                }
            }
        }
        if (deriv && i >= 1) {
            unsigned* W = win[i % 3];
            W[0] = __funnelshift_r(w0, w1, 24);
            W[1] = w1;
            W[2] = __funnelshift_r(w1, w2, 8);
            W[3] = __funnelshift_r(w1, w2, 16);
            W[4] = __funnelshift_r(w1, w2, 24);
            W[5] = w2;
            W[6] = __funnelshift_r(w2, w3, 8);
            W[7] = __funnelshift_r(w2, w3, 16);
            if (i >= 3) {                      // rows i-2, i-1, i are the 3x3 neighbourhood of image row yb + i - 3
                const int y = yb + i - 3;
                if (y < h) {
                    const unsigned* A = win[(i - 2) % 3];
                    const unsigned* C = win[(i - 1) % 3];
                    unsigned o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int ix = dp4a_us(A[j], 0x000300FD, 0);            // (-3, 0, 3, 0)
                        ix = dp4a_us(C[j], 0x000A00F6, ix);               // (-10, 0, 10, 0)
                        ix = dp4a_us(W[j], 0x000300FD, ix);
                        int iy = dp4a_us(A[j], 0x00FDF6FD, 0);            // (-3, -10, -3, 0)
                        iy = dp4a_us(W[j], 0x00030A03, iy);               // (3, 10, 3, 0)
                        o[j] = __byte_perm((unsigned)ix, (unsigned)iy, 0x5410);   // short2(ix, iy)
                    }
                    short2* out = deriv + (size_t)b * dstride + (size_t)y * dpitch + x;
                    if (full) {
                        reinterpret_cast<uint4*>(out)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        reinterpret_cast<uint4*>(out)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (x + j < w) reinterpret_cast<unsigned*>(out)[j] = o[j];
                    }
                }
            }
        }
    }
    }
    __syncthreads();   // the buffer is restaged two tiles ahead
    }
}

// ================================================================================================================
// TMA-staged pyramid construction (north_star: "TMA-staged image-pyramid construction").
//
// The images are described to the tensor memory accelerator as 3-D tensors of 32-bit words (x/4, y, image): one elected
// thread issues cp.async.bulk.tensor for a whole tile incl. its halo and the bytes land in shared memory behind an
// mbarrier — no per-thread address arithmetic, no cp.async bookkeeping in the other warps.  A box must start on a 16-byte
// boundary of the innermost dimension (a start at -8 bytes raises an illegal-instruction fault on the B200, tools/tma_min.cu),
// hence the 16-pixel halo to the left of every staged tile.  TMA fills out-of-image elements with zeros, OpenCV's pyramid wants BORDER_REFLECT_101: the few halo columns / rows of tiles that touch the image
// border are patched in shared memory after the tile has landed (the mirrored pixels are inside the same tile).
//
//   klt_level0_tma_kernel   level 0 (any level with a TMA-eligible source): 256 x 64 pixel tiles, a warp owns an 8-row band of a
//                           tile and its lanes 32 consecutive 8-pixel groups, so every shared-memory access of a warp is one
//                           contiguous 128/256-byte segment (the cp.async kernel above had four row groups of a warp on the same
//                           banks); two tiles in flight per CTA (double buffer, one mbarrier each).
//   klt_levels_fused_kernel levels 1 .. 3 of one image per CTA: level 1 arrives by TMA as one box and stays in shared memory
//                           together with the levels derived from it; derivatives and pyrDown results of all three levels
//                           leave in one launch instead of three (the small levels were latency / launch bound).
// Per-thread arithmetic (level_block) is the IDP4A formulation of klt_level_kernel, bit-identical to cv::pyrDown / cv::Scharr.

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tile (x in 32-bit words, y, image) -> shared memory; completion is signalled on `bar` with the byte count
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, unsigned long long* bar, int cx, int cy, int cz) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(cx), "r"(cy), "r"(cz)
                 : "memory");
}
// generic-proxy accesses to a shared-memory buffer (reads of the previous tile, border patches) are ordered before the
// async-proxy write of the next TMA into it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One thread: 8 columns x RPT rows of a level from a staged tile whose halo is materialised (REFLECT_101 already applied).
// t0 points at the staged pixel (x - 8, yb - 2).  Optional outputs: pass-through copy, interleaved Scharr derivatives, pyrDown
// rows of the next level to global memory and / or to a shared-memory image (`down_s`, pixel (0,0) of the next level).
template <int RPTT>
__device__ __forceinline__ void level_block(const uint8_t* __restrict__ t0, int tpitch, int x, int yb, int w, int h, uint8_t* __restrict__ copy_img,
                                            int cpitch, short2* __restrict__ deriv, int dpitch, uint8_t* __restrict__ down, int npitch,
                                            uint8_t* __restrict__ down_s, int nspitch) {
    const bool full = x + 7 < w;
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    const int gx = x / 2, gy = yb / 2;
    const bool want_down = down || down_s;
    unsigned win[3][8];                        // Scharr windows (x-1, x, x+1, x+2) of the last three rows
    unsigned acc[RPTT / 2][4];                 // pyrDown sums of output rows gy .. gy + RPT/2 - 1
#pragma unroll
    for (int o = 0; o < RPTT / 2; ++o)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[o][k] = 0u;
#pragma unroll
    for (int i = 0; i < RPTT + 3; ++i) {       // staged row i  <->  image row yb - 2 + i
        const uint8_t* tr = t0 + i * tpitch;
        const unsigned w0 = *reinterpret_cast<const unsigned*>(tr + 4);     // pixels x-4 .. x-1
        const uint2 w12 = *reinterpret_cast<const uint2*>(tr + 8);          // pixels x .. x+7
        const unsigned w3 = *reinterpret_cast<const unsigned*>(tr + 16);    // pixels x+8 .. x+11
        const unsigned w1 = w12.x, w2 = w12.y;
        if (copy_img && i >= 2 && i < RPTT + 2) {
            const int y = yb + i - 2;
            if (y < h) *reinterpret_cast<uint2*>(copy_img + (size_t)y * cpitch + x) = w12;
        }
        if (want_down) {
            unsigned hz[4];
            hz[0] = dp4a_uu(w1, 0x00010406u, dp4a_uu(w0, 0x04010000u, 0u));
            hz[1] = dp4a_uu(w2, 0x00000001u, dp4a_uu(w1, 0x04060401u, 0u));
            hz[2] = dp4a_uu(w2, 0x00010406u, dp4a_uu(w1, 0x04010000u, 0u));
            hz[3] = dp4a_uu(w3, 0x00000001u, dp4a_uu(w2, 0x04060401u, 0u));
#pragma unroll
            for (int o = 0; o < RPTT / 2; ++o) {
                const int d = i - (2 * o + 2);
                const unsigned wv = (d == 0) ? 6u : (d == 1 || d == -1) ? 4u : (d == 2 || d == -2) ? 1u : 0u;
                if (wv) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[o][k] += wv * hz[k];
                }
                if (d == 2 && gy + o < dh) {   // row complete: round, pack, store
                    const unsigned r0 = (acc[o][0] + 128u) >> 8, r1 = (acc[o][1] + 128u) >> 8, r2 = (acc[o][2] + 128u) >> 8, r3 = (acc[o][3] + 128u) >> 8;
                    const unsigned packed = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
                    if (down) {
                        uint8_t* nd = down + (size_t)(gy + o) * npitch + gx;
                        if (gx + 3 < dw) *reinterpret_cast<unsigned*>(nd) = packed;
                        else {
                            nd[0] = (uint8_t)r0;
                            if (gx + 1 < dw) nd[1] = (uint8_t)r1;
                            if (gx + 2 < dw) nd[2] = (uint8_t)r2;
                        }
                    }
                    if (down_s) *reinterpret_cast<unsigned*>(down_s + (gy + o) * nspitch + gx) = packed;   // (columns >= dw are patched afterwards)
                }
            }
        }
        if (deriv && i >= 1) {
            unsigned* W = win[i % 3];
            W[0] = __funnelshift_r(w0, w1, 24);
            W[1] = w1;
            W[2] = __funnelshift_r(w1, w2, 8);
            W[3] = __funnelshift_r(w1, w2, 16);
            W[4] = __funnelshift_r(w1, w2, 24);
            W[5] = w2;
            W[6] = __funnelshift_r(w2, w3, 8);
            W[7] = __funnelshift_r(w2, w3, 16);
            if (i >= 3) {
                const int y = yb + i - 3;
                if (y < h) {
                    const unsigned* A = win[(i - 2) % 3];
                    const unsigned* C = win[(i - 1) % 3];
                    unsigned o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int ix = dp4a_us(A[j], 0x000300FD, 0);
                        ix = dp4a_us(C[j], 0x000A00F6, ix);
                        ix = dp4a_us(W[j], 0x000300FD, ix);
                        int iy = dp4a_us(A[j], 0x00FDF6FD, 0);
                        iy = dp4a_us(W[j], 0x00030A03, iy);
                        o[j] = __byte_perm((unsigned)ix, (unsigned)iy, 0x5410);
                    }
                    short2* out = deriv + (size_t)y * dpitch + x;
                    if (full) {
                        reinterpret_cast<uint4*>(out)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        reinterpret_cast<uint4*>(out)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (x + j < w) reinterpret_cast<unsigned*>(out)[j] = o[j];
                    }
                }
            }
        }
    }
}

// REFLECT_101 halo of a staged image region: `img` points at staged pixel (X0, Y0) of a level of size w x h, the region holds
// rows [Y0, Y0 + rows) and columns [X0, X0 + cols).  Columns x = -2, -1, w, w + 1 of the in-image rows first, then whole rows
// y = -2, -1, h, h + 1 (callers synchronise between the two and afterwards).
__device__ __forceinline__ void patch_cols(uint8_t* img, int pitch, int X0, int Y0, int cols, int rows, int w, int h, int tid, int nthreads) {
    const bool left = X0 < 0, right = X0 + cols > w;
    if (!left && !right) return;
    for (int r = tid; r < rows; r += nthreads) {
        const int y = Y0 + r;
        if (y < 0 || y >= h) continue;
        uint8_t* row = img + r * pitch - X0;                 // row[x] = pixel (x, y)
        if (left) { row[-1] = row[1]; row[-2] = row[2]; }
        if (right) { row[w] = row[w - 2]; if (X0 + cols > w + 1) row[w + 1] = row[w - 3]; }
    }
}
__device__ __forceinline__ void patch_rows(uint8_t* img, int pitch, int X0, int Y0, int cols, int rows, int w, int h, int tid, int nthreads) {
    (void)X0; (void)w;
    const int words = cols >> 2;
    for (int k = 0; k < 4; ++k) {
        const int y = k < 2 ? k - 2 : h + (k - 2);           // -2, -1, h, h + 1
        const int r = y - Y0;
        if (r < 0 || r >= rows) continue;
        const int ys = y < 0 ? -y : 2 * (h - 1) - y, rs = ys - Y0;
        if (rs < 0 || rs >= rows) continue;
        const unsigned* src = reinterpret_cast<const unsigned*>(img + rs * pitch);
        unsigned* dst = reinterpret_cast<unsigned*>(img + r * pitch);
        for (int c = tid; c < words; c += nthreads) dst[c] = src[c];
    }
}

#ifndef KLT_L0_CTAS
#define KLT_L0_CTAS 3
#endif
#ifndef KLT_FUSED_THREADS
#define KLT_FUSED_THREADS 256
#endif
constexpr int ATW = 256, ATH = 64;                 // tile of klt_level0_tma_kernel (pixels)
constexpr int HX = 16;                             // halo columns staged left and right of a tile (TMA box start: 16-byte aligned)
constexpr int ASW = ATW + 2 * HX, ASH = ATH + 4;   // staged: x0-16 .. x0+ATW+15, y0-2 .. y0+ATH+1
constexpr int ANT = 256;                           // threads: 8 warps = 8 row bands of RPT rows, 32 lanes = 32 column groups of 8 pixels

__global__ void __launch_bounds__(ANT, KLT_L0_CTAS) klt_level0_tma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                                                                TmaJob j0, TmaJob j1, int w, int h, int cpitch, size_t cstride, int dpitch,
                                                                size_t dstride, int npitch, size_t nstride, int tiles_per_cta) {
    constexpr int TILE_STRIDE = (ASH * ASW + 127) / 128 * 128;   // TMA destinations are 128-byte aligned
    __shared__ __align__(128) uint8_t tiles[2][TILE_STRIDE];
    __shared__ __align__(8) unsigned long long bars[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * ATW;
    int b = blockIdx.z;
    const bool second = b >= j0.batch;
    if (second) b -= j0.batch;
    const TmaJob& J = second ? j1 : j0;
    if (!J.copy_dst && !J.deriv && !J.down) return;
    const int ty_begin = blockIdx.y * tiles_per_cta;
    const int ty_end = min(ty_begin + tiles_per_cta, (h + ATH - 1) / ATH);
    if (ty_begin >= ty_end) return;
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr unsigned TILE_BYTES = ASH * ASW;
    auto issue = [&](int ty, int buf) {            // one elected thread
        mbar_expect_tx(&bars[buf], TILE_BYTES);
        // (the descriptor must be addressed in parameter space: no pointer selected at run time, which would make a local copy)
        if (second) tma_load_3d(tiles[buf], &map1, &bars[buf], (x0 - HX) / 4, ty * ATH - 2, J.first + b);
        else tma_load_3d(tiles[buf], &map0, &bars[buf], (x0 - HX) / 4, ty * ATH - 2, J.first + b);
    };
    if (tid == 0) issue(ty_begin, 0);
    uint8_t* copy_img = J.copy_dst ? J.copy_dst + (size_t)b * cstride : nullptr;
    short2* deriv = J.deriv ? J.deriv + (size_t)b * dstride : nullptr;
    uint8_t* down = J.down ? J.down + (size_t)b * nstride : nullptr;
    for (int ty = ty_begin; ty < ty_end; ++ty) {
        const int it = ty - ty_begin, buf = it & 1;
        if (tid == 0 && ty + 1 < ty_end) issue(ty + 1, buf ^ 1);      // (that buffer was released by the __syncthreads ending the previous iteration)
        mbar_wait(&bars[buf], (it >> 1) & 1);
        uint8_t* tile = tiles[buf];
        const int y0 = ty * ATH;
        const bool border = x0 == 0 || x0 + ATW + HX > w || y0 == 0 || y0 + ATH + 2 > h;
        if (border) {                              // uniform per CTA
            patch_cols(tile, ASW, x0 - HX, y0 - 2, ASW, ASH, w, h, tid, ANT);
            __syncthreads();
            patch_rows(tile, ASW, x0 - HX, y0 - 2, ASW, ASH, w, h, tid, ANT);
            __syncthreads();
        }
        const int x = x0 + lane * 8, yb = y0 + warp * RPT;
        if (x < w && yb < h)
            level_block<RPT>(tile + (warp * RPT) * ASW + lane * 8 + (HX - 8), ASW, x, yb, w, h, copy_img, cpitch, deriv, dpitch, down, npitch, nullptr, 0);
        fence_proxy_async();
        __syncthreads();                           // everyone is done with this buffer: it may be refilled by the TMA issued next iteration
    }
}

// Levels 1 .. nl of one image per CTA (nl <= 3).  Shared memory: the levels with a halo of HX columns left / right and 2 rows
// above / below (pitch = level pitch + 2 HX, rows = roundup(h, 8) + 4: a thread's 8-row block may overhang the image); level 3 reuses
// the space of level 1, which is dead by then.
constexpr int FNT = KLT_FUSED_THREADS;

__global__ void __launch_bounds__(FNT, 2) klt_levels_fused_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                                                                  FusedJob j0, FusedJob j1, FusedLevel l1, FusedLevel l2, FusedLevel l3, int nl) {
    extern __shared__ __align__(128) uint8_t fsm_raw[];
    __shared__ __align__(8) unsigned long long bar;
    uint8_t* fsm = fsm_raw + ((128u - (smem_u32(fsm_raw) & 127u)) & 127u);      // TMA destinations are 128-byte aligned
    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const bool second = b >= j0.batch;
    if (second) b -= j0.batch;
    const FusedJob& J = second ? j1 : j0;
    const int img = J.first + b;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, (unsigned)(l1.spitch * l1.srows));
        if (second) tma_load_3d(fsm + l1.soff, &map1, &bar, -HX / 4, -2, img);      // the whole level incl. halo: one box
        else tma_load_3d(fsm + l1.soff, &map0, &bar, -HX / 4, -2, img);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
#pragma unroll
    for (int li = 0; li < 3; ++li) {
        if (li >= nl) break;
        const FusedLevel& L = li == 0 ? l1 : (li == 1 ? l2 : l3);
        uint8_t* simg = fsm + L.soff;                                    // staged pixel (-HX, -2)
        patch_cols(simg, L.spitch, -HX, -2, L.spitch, L.srows, L.w, L.h, tid, FNT);
        __syncthreads();
        patch_rows(simg, L.spitch, -HX, -2, L.spitch, L.srows, L.w, L.h, tid, FNT);
        __syncthreads();
        const bool last = li + 1 >= nl;
        short2* deriv = J.want_deriv ? reinterpret_cast<short2*>(J.slot + L.der_off + (size_t)img * L.der_stride) : nullptr;
        uint8_t* down = nullptr; uint8_t* down_s = nullptr; int npitch = 0, nspitch = 0;
        if (!last) {
            const FusedLevel& Nx = li == 0 ? l2 : l3;
            down = J.slot + Nx.img_off + (size_t)img * Nx.img_stride; npitch = Nx.pitch;
            down_s = fsm + Nx.soff + 2 * Nx.spitch + HX; nspitch = Nx.spitch;
        }
        if (deriv || !last) {
            const int bx = (L.w + 7) / 8, by = (L.h + RPT - 1) / RPT;
            for (int e = tid; e < bx * by; e += FNT) {
                const int cx = e % bx, cy = e / bx;
                const int x = cx * 8, yb = cy * RPT;
                level_block<RPT>(simg + yb * L.spitch + x + (HX - 8), L.spitch, x, yb, L.w, L.h, nullptr, 0, deriv, L.dpitch, down, npitch, down_s, nspitch);
            }
        }
        __syncthreads();
    }
}

// exact warp-wide sum of 32-bit partials in 64 bits: two REDUX.SUM on the 16-bit halves
__device__ __forceinline__ long long warp_sum_ll(int v) {
    const int lo = v & 0xffff, hi = v >> 16;       // v == hi * 65536 + lo
    const int slo = __reduce_add_sync(0xffffffffu, lo), shi = __reduce_add_sync(0xffffffffu, hi);
    return (long long)shi * 65536 + slo;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    w00 = __float2int_rn((1.f - a) * (1.f - b) * (float)(1 << 14));
    w01 = __float2int_rn(a * (1.f - b) * (float)(1 << 14));
    w10 = __float2int_rn((1.f - a) * b * (float)(1 << 14));
    w11 = (1 << 14) - w00 - w01 - w10;
}

constexpr int WARPS = 4;      // features per CTA
constexpr int SEG = 7;        // pixels per row segment: 8 tile bytes give 7 horizontal tap pairs

__host__ __device__ inline int track_segs_per_row(int win) { return (win + SEG - 1) / SEG; }
// bytes per staged u8 tile row: the last segment reads three aligned words from its start.  The word count is == 4 (mod 8): a
// warp's 32 segment reads span 11 rows x a 4-bank window each, and with a row pitch of 4 banks (mod 32) eight consecutive rows
// land on distinct banks — two wavefronts per load, the minimum for 11 rows (an odd pitch gave up to three: 44 % of the
// kernel's shared-memory wavefronts were bank conflicts, profiles/r01_ncu_full_summary.txt)
__host__ __device__ inline int track_tile_stride(int win, int mode = 1) {
    int words = ((SEG * (track_segs_per_row(win) - 1)) >> 2) + 3;
    if (words * 4 < win + 1) words = (win + 4) / 4;
    if (mode == 0) return (words | 1) * 4;         // (round 1's odd pitch, kept for A/B measurements: EKFVIO_KLT_TRACK_STRIDE=0)
    while ((words & 7) != 4) ++words;
    return words * 4;
}
__host__ __device__ inline size_t track_warp_bytes(int win, int mode = 1) {
    size_t nseg = (size_t)win * track_segs_per_row(win), t = ((size_t)win + 1) * (win + 1);
    size_t bytes = nseg * 8 * 4 /*Cp*/ + nseg * 8 * 4 /*dI*/ + t * 4 /*Dt*/ + ((size_t)win + 1) * track_tile_stride(win, mode) /*Jt*/;
    return (bytes + 15) & ~(size_t)15;
}

// (signed 16-bit weights: rounding can leave w11 = -1)
__device__ __forceinline__ unsigned dp2a_lo(unsigned a_u16x2, unsigned b_u8x4, unsigned c) {
    unsigned d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u16x2), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned dp2a_hi(unsigned a_u16x2, unsigned b_u8x4, unsigned c) {
    unsigned d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u16x2), "r"(b_u8x4), "r"(c));
    return d;
}

// Bilinear taps of one 7-pixel row segment: v[k] = c[k] + w00 t[k] + w01 t[k+1] + w10 b[k] + w11 b[k+1]
// where t / b are the 8 tile bytes of the segment in the upper / lower row.  wt = w00 | w01 << 16,
// wb = w10 | w11 << 16 (Q14 weights fit signed 16 bits), two IDP2A per pixel.
__device__ __forceinline__ void seg_taps(const uint8_t* Jt, int JS, int al, int r, int sg, unsigned wt, unsigned wb, const int (&c)[SEG], int (&v)[SEG]) {
    const int off = SEG * sg + al, sh = (off & 3) * 8;        // al: byte offset of the tile's first column inside the staged row (0..3)
    const unsigned* tr = reinterpret_cast<const unsigned*>(Jt + r * JS) + (off >> 2);
    const unsigned* br = tr + (JS >> 2);
    const unsigned t0 = tr[0], t1 = tr[1], t2 = tr[2], b0 = br[0], b1 = br[1], b2 = br[2];
    const unsigned tlo = __funnelshift_r(t0, t1, sh), thi = __funnelshift_r(t1, t2, sh);   // bytes 0..3, 4..7
    const unsigned blo = __funnelshift_r(b0, b1, sh), bhi = __funnelshift_r(b1, b2, sh);
    const unsigned tm = __funnelshift_r(tlo, thi, 8), tm2 = thi >> 8;                      // bytes 1..4, 5..7
    const unsigned bm = __funnelshift_r(blo, bhi, 8), bm2 = bhi >> 8;
    v[0] = (int)dp2a_lo(wb, blo, dp2a_lo(wt, tlo, (unsigned)c[0]));
    v[1] = (int)dp2a_lo(wb, bm, dp2a_lo(wt, tm, (unsigned)c[1]));
    v[2] = (int)dp2a_hi(wb, blo, dp2a_hi(wt, tlo, (unsigned)c[2]));
    v[3] = (int)dp2a_hi(wb, bm, dp2a_hi(wt, tm, (unsigned)c[3]));
    v[4] = (int)dp2a_lo(wb, bhi, dp2a_lo(wt, thi, (unsigned)c[4]));
    v[5] = (int)dp2a_lo(wb, bm2, dp2a_lo(wt, tm2, (unsigned)c[5]));
    v[6] = (int)dp2a_hi(wb, bhi, dp2a_hi(wt, thi, (unsigned)c[6]));
}

// LKTrackerInvoker for every level, one warp per point.  The win x win patch is cut into row
// segments of 7 pixels (8 slots each); a lane owns segments lane, lane + 32, ...  Per-warp shared memory:
//   Cp[2][nseg*4] int32 (slots 0-3 of every segment, then slots 4-7: consecutive lanes read consecutive 16-byte vectors): 256 - 512 * (patch intensity), the rounding constant and the patch value folded
//                     into the accumulator the IDP2A chain starts from, so diff = chain >> 9;
//   dI[nseg*8] short2 (Ix, Iy of the patch; zero in unused slots),
//   Dt[(win+1)^2] short2 (staged derivative tile), Jt[(win+1) x stride] u8 (staged I or J tile).
// Tiles are staged row by row (lanes = columns), so every bilinear tap is a shared-memory read and
// no per-pixel integer division or border test remains in the loops.
__global__ void __launch_bounds__(WARPS * 32) klt_track_kernel(Pyr pyr, const uint8_t* __restrict__ prev_slot, const uint8_t* __restrict__ next_slot,
                                                               const float* __restrict__ prev_pts, float* __restrict__ next_pts,
                                                               uint8_t* __restrict__ status, float* __restrict__ err, const int* __restrict__ npts,
                                                               int max_points, int win, int max_count, double epsilon, double min_eig,
                                                               int use_initial_flow, int first_image, int stride_mode, ExtLevel0 prev_ext, ExtLevel0 next_ext) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y + first_image;      // (a launch may cover a sub-range of the batch: chunked uploads)
    const int pt = blockIdx.x * WARPS + warp;
    if (pt >= npts[b]) return;
    const int jw1 = win + 1, tarea = jw1 * jw1;
    const int SPR = track_segs_per_row(win), nseg = win * SPR, JS = track_tile_stride(win, stride_mode);
    uint8_t* base = smem_raw + warp * track_warp_bytes(win, stride_mode);
    int* Cp = reinterpret_cast<int*>(base);
    unsigned* dI = reinterpret_cast<unsigned*>(Cp + nseg * 8);
    short2* Dt = reinterpret_cast<short2*>(dI + nseg * 8);
    uint8_t* Jt = reinterpret_cast<uint8_t*>(Dt + tarea);
    // lane's segments: s = lane, lane + 32, ...  ->  (row, segment in row) advanced without divisions
    const int r_first = lane / SPR, sg_first = lane - r_first * SPR;
    const int qstep = 32 / SPR, rstep = 32 - qstep * SPR;

    const size_t pidx = ((size_t)b * max_points + pt) * 2;
    const float prevx = prev_pts[pidx], prevy = prev_pts[pidx + 1];
    float stx = use_initial_flow ? next_pts[pidx] : prevx, sty = use_initial_flow ? next_pts[pidx + 1] : prevy;  // stored nextPts
    int st = 1;
    float errv = 0.f;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = pyr.levels - 1;

    // stage a (win+1)^2 tile of an 8-bit image at (ox, oy) with REFLECT_101 borders; returns the byte offset of the tile's first
    // column inside the staged rows.  A tile whose columns lie inside an image with word-aligned rows is fetched as aligned 32-bit words — eight
    // word slots per row, four rows per pass, all lanes loading (six passes for a 22-row tile, every load issued before the first
    // store can block) — and keeps its alignment offset; the per-column byte loop (22 dependent-latency rounds of 22 bytes each,
    // the kernel's largest stall: long_scoreboard 29 %, profiles/r02_ncu_klt_summary.txt) remains for tiles across the left / right border and odd pitches.
    const int jwords_max = (3 + jw1 + 3) >> 2;
    const bool words_fit = jwords_max <= 8 && JS >= 4 * jwords_max && JS >= 4 * (((SEG * (SPR - 1) + 3) >> 2) + 3);   // (a segment reads three words from its start)
    auto stage_u8 = [&](const uint8_t* img, int ipitch, const Level& L, int ox, int oy) -> int {
        const bool in_x = ox >= 0 && ox + win < L.w, in_y = oy >= 0 && oy + win < L.h, inside = in_x && in_y;
        int al = 0;
        if (in_x && words_fit && ((reinterpret_cast<size_t>(img) | (size_t)ipitch) & 3) == 0) {      // (rows may still reflect)
            al = ox & 3;
            const int nw = (al + jw1 + 3) >> 2, w = lane & 7;
            const uint8_t* p = img + (ox - al) + 4 * w;
            if (w < nw)
                for (int y = lane >> 3; y < jw1; y += 4) {
                    const int Y = in_y ? oy + y : reflect101(oy + y, L.h);
                    *reinterpret_cast<unsigned*>(Jt + y * JS + 4 * w) = *reinterpret_cast<const unsigned*>(p + (size_t)Y * ipitch);
                }
        } else if (lane < jw1) {
            if (inside) {
                const uint8_t* p = img + (size_t)oy * ipitch + ox + lane;
                for (int y = 0; y < jw1; ++y) Jt[y * JS + lane] = p[(size_t)y * ipitch];
            } else {
                const int cx = reflect101(ox + lane, L.w);
                for (int y = 0; y < jw1; ++y) Jt[y * JS + lane] = img[(size_t)reflect101(oy + y, L.h) * ipitch + cx];
            }
        }
        __syncwarp();
        return al;
    };

    for (int level = top; level >= 0; --level) {
        const Level L = pyr.lv[level];
        // level 0 may be the caller's own image batch (ekfvio_klt_build_pyramid_pair_ref)
        const bool iext = level == 0 && prev_ext.img, jext = level == 0 && next_ext.img;
        const uint8_t* I = iext ? prev_ext.img + (size_t)b * prev_ext.stride : prev_slot + L.img_off + (size_t)b * L.img_stride;
        const uint8_t* J = jext ? next_ext.img + (size_t)b * next_ext.stride : next_slot + L.img_off + (size_t)b * L.img_stride;
        const int ipitch = iext ? prev_ext.pitch : L.pitch, jpitch = jext ? next_ext.pitch : L.pitch;
        const short2* dIm = reinterpret_cast<const short2*>(prev_slot + L.der_off + (size_t)b * L.der_stride);
        const float scale = (float)(1. / (1 << level));
        float px = prevx * scale, py = prevy * scale;
        float nx, ny;
        if (level == top) {
            if (use_initial_flow) { nx = stx * scale; ny = sty * scale; }
            else { nx = px; ny = py; }
        } else { nx = stx * 2.f; ny = sty * 2.f; }
        stx = nx; sty = ny;
        px -= half; py -= half;
        int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) {
            if (level == 0) { st = 0; errv = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(px - ipx, py - ipy, w00, w01, w10, w11);
        // stage the I tile (into Jt) and the derivative tile (zero outside the image)
        __syncwarp();
        const int ial = stage_u8(I, ipitch, L, ipx, ipy);
        if (lane < jw1) {
            const int X = ipx + lane;
            const bool xin = (unsigned)X < (unsigned)L.w;
            for (int y = 0; y < jw1; ++y) {
                const int Y = ipy + y;
                Dt[y * jw1 + lane] = (xin && (unsigned)Y < (unsigned)L.h) ? dIm[(size_t)Y * L.dpitch + X] : make_short2(0, 0);
            }
        }
        __syncwarp();
        int sa11 = 0, sa12 = 0, sa22 = 0;     // per-lane partial sums fit 32 bits (<= 31 slots x 4080^2)
        {
            const unsigned wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16), wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);
            const int c256[SEG] = {256, 256, 256, 256, 256, 256, 256};
            for (int sI = lane, r = r_first, sg = sg_first; sI < nseg; sI += 32) {
                int v[SEG];
                seg_taps(Jt, JS, ial, r, sg, wt, wb, c256, v);
                int cw[8];
                unsigned dw[8];
#pragma unroll
                for (int k = 0; k < SEG; ++k) {
                    const int x = SEG * sg + k;
                    if (x < win) {
                        const int ival = v[k] >> 9;
                        const int o = r * jw1 + x;
                        const short2 d00 = Dt[o], d01 = Dt[o + 1], d10 = Dt[o + jw1], d11 = Dt[o + jw1 + 1];
                        const int ixval = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
                        const int iyval = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
                        cw[k] = 256 - 512 * (int)(short)ival;
                        dw[k] = ((unsigned)ixval & 0xffffu) | ((unsigned)iyval << 16);
                        const int sx = (short)ixval, sy = (short)iyval;
                        sa11 += sx * sx; sa12 += sx * sy; sa22 += sy * sy;
                    } else { cw[k] = 0; dw[k] = 0u; }
                }
                cw[7] = 0; dw[7] = 0u;
                *reinterpret_cast<int4*>(Cp + sI * 4) = make_int4(cw[0], cw[1], cw[2], cw[3]);
                *reinterpret_cast<int4*>(Cp + (nseg + sI) * 4) = make_int4(cw[4], cw[5], cw[6], cw[7]);
                *reinterpret_cast<uint4*>(dI + sI * 4) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                *reinterpret_cast<uint4*>(dI + (nseg + sI) * 4) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
                r += qstep; sg += rstep;
                if (sg >= SPR) { sg -= SPR; ++r; }
            }
        }
        const long long a11 = warp_sum_ll(sa11), a12 = warp_sum_ll(sa12), a22 = warp_sum_ll(sa22);
        float A11 = (float)a11 * FLT_SCALE, A12 = (float)a12 * FLT_SCALE, A22 = (float)a22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < min_eig || D < FLT_EPSILON) {
            if (level == 0) st = 0;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int sjx = INT_MIN, sjy = INT_MIN, jal = 0;   // cell (and alignment offset) of the J tile currently staged (none: Jt holds the I tile)
        for (int j = 0; j < max_count; ++j) {
            int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            lk_weights(nx - inx, ny - iny, w00, w01, w10, w11);
            if (inx != sjx || iny != sjy) {       // the staged J tile is reused while the integer cell does not move
                __syncwarp();
                jal = stage_u8(J, jpitch, L, inx, iny);
                sjx = inx; sjy = iny;
            }
            int sb1 = 0, sb2 = 0;                 // <= 35 slots x 8160 x 4080 fits 32 bits
            {
                const unsigned wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16), wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);
                for (int sI = lane, r = r_first, sg = sg_first; sI < nseg; sI += 32) {
                    const int4 c0 = *reinterpret_cast<const int4*>(Cp + sI * 4), c1 = *reinterpret_cast<const int4*>(Cp + (nseg + sI) * 4);
                    const uint4 d0 = *reinterpret_cast<const uint4*>(dI + sI * 4), d1 = *reinterpret_cast<const uint4*>(dI + (nseg + sI) * 4);
                    const int c[SEG] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z};
                    const unsigned dw[SEG] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z};
                    int v[SEG];
                    seg_taps(Jt, JS, jal, r, sg, wt, wb, c, v);
#pragma unroll
                    for (int k = 0; k < SEG; ++k) {   // unused slots carry zero derivatives
                        const int diff = v[k] >> 9;
                        sb1 += diff * (int)(short)(dw[k] & 0xffffu);
                        sb2 += diff * ((int)dw[k] >> 16);
                    }
                    r += qstep; sg += rstep;
                    if (sg >= SPR) { sg -= SPR; ++r; }
                }
            }
            const long long b1 = warp_sum_ll(sb1), b2 = warp_sum_ll(sb2);
            float fb1 = (float)b1 * FLT_SCALE, fb2 = (float)b2 * FLT_SCALE;
            float dx = (A12 * fb2 - A22 * fb1) * D, dy = (A12 * fb1 - A11 * fb2) * D;
            nx += dx; ny += dy;
            stx = nx + half; sty = ny + half;
            if ((double)dx * (double)dx + (double)dy * (double)dy <= epsilon) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                stx -= dx * 0.5f; sty -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && level == 0) {  // final bounds check + residual (err is always requested by the reference)
            float fx = stx - half, fy = sty - half;
            int inx = (int)floorf(fx), iny = (int)floorf(fy);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                st = 0;
            } else {
                lk_weights(fx - inx, fy - iny, w00, w01, w10, w11);
                if (inx != sjx || iny != sjy) {
                    __syncwarp();
                    jal = stage_u8(J, jpitch, L, inx, iny);
                }
                int es = 0;                       // <= 35 slots x 8160
                {
                    const unsigned wt = ((unsigned)w00 & 0xffffu) | ((unsigned)w01 << 16), wb = ((unsigned)w10 & 0xffffu) | ((unsigned)w11 << 16);
                    for (int sI = lane, r = r_first, sg = sg_first; sI < nseg; sI += 32) {
                        const int4 c0 = *reinterpret_cast<const int4*>(Cp + sI * 4), c1 = *reinterpret_cast<const int4*>(Cp + (nseg + sI) * 4);
                        const int c[SEG] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z};
                        int v[SEG];
                        seg_taps(Jt, JS, jal, r, sg, wt, wb, c, v);
#pragma unroll
                        for (int k = 0; k < SEG; ++k) {
                            const int diff = v[k] >> 9;
                            if (SEG * sg + k < win) es += diff < 0 ? -diff : diff;
                        }
                        r += qstep; sg += rstep;
                        if (sg >= SPR) { sg -= SPR; ++r; }
                    }
                }
                const long long est = warp_sum_ll(es);
                errv = (float)est * 1.f / (float)(32 * win * win);
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        next_pts[pidx] = stx; next_pts[pidx + 1] = sty;
        status[(size_t)b * max_points + pt] = (uint8_t)st;
        if (err) err[(size_t)b * max_points + pt] = errv;
    }
}

// KLTTracker.cpp:72-92 + Feature::pixel2Metric (Feature.h:60-62).  K9: column-major 3x3 per image.
__global__ void klt_postprocess_kernel(const float* __restrict__ next_pts, const uint8_t* __restrict__ status, const int* __restrict__ npts,
                                       const float* __restrict__ K9, int max_points, int cols, int rows, int kill_pad,
                                       float* __restrict__ measured, float* __restrict__ cov, uint8_t* __restrict__ passed) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts[b]) return;
    const size_t o = (size_t)b * max_points + i;
    const float* K = K9 + (size_t)b * 9;
    float x = next_pts[o * 2], y = next_pts[o * 2 + 1];
    float pad = (float)kill_pad;
    if (status[o] == 1 && !(x < pad || y < pad || (float)cols - x < pad || (float)rows - y < pad)) {
        passed[o] = 1;
        float sx = (float)pow(1.0 / (double)K[0], 2.0), sy = (float)pow(1.0 / (double)K[4], 2.0);
        cov[o * 4 + 0] = 0.00001f * sx; cov[o * 4 + 1] = 0.f * sx;
        cov[o * 4 + 3] = 0.00001f * sy; cov[o * 4 + 2] = 0.f * sy;
        measured[o * 2] = (x - K[2]) / K[0];      // K(2) == K(2,0): E1, principal point dropped
        measured[o * 2 + 1] = (y - K[5]) / K[4];  // K(5) == K(2,1)
    } else {
        passed[o] = 0;
        cov[o * 4 + 0] = cov[o * 4 + 1] = cov[o * 4 + 2] = cov[o * 4 + 3] = 0.f;
    }
}

// KLTTracker::estimateUncertaintySampleBased (KLTTracker.cpp:111-175; dead code in the reference — nothing calls it): a 5x5
// reference patch around mu_ref in the last frame (cv::getRectSubPix, CV_32F) against 5x5 patches of the current frame at the 25
// offsets (du, dv) in {-10, -5, 0, 5, 10}^2 around mu; rd = exp(-0.01 * SSD / 25) weights the second moments of the offsets.
// One warp per feature: lane s < 25 evaluates sample s, lane 0 accumulates the four sums in the reference's order (du outer,
// dv inner), so the float result does not depend on a reduction tree.  getRectSubPix as OpenCV's 8u -> 32f routine evaluates it:
// a = max(frac x, 1e-4), a running "prev" term along the row, replicated borders.
__device__ __forceinline__ void rect_subpix5(const uint8_t* __restrict__ img, int pitch, int w, int h, float cx, float cy, float (&out)[25]) {
    cx -= 2.f; cy -= 2.f;                               // (win_size - 1) * 0.5
    const int ix = (int)floorf(cx), iy = (int)floorf(cy);
    float a = cx - ix; const float b = cy - iy;
    a = fmaxf(a, 0.0001f);
    const float a12 = a * (1.f - b), a22 = a * b, b1 = 1.f - b, b2 = b, sc = (1.f - a) / a;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int y0 = min(max(iy + i, 0), h - 1), y1 = min(max(iy + i + 1, 0), h - 1);
        const uint8_t* r0 = img + (size_t)y0 * pitch;
        const uint8_t* r1 = img + (size_t)y1 * pitch;
        const int x0 = min(max(ix, 0), w - 1);
        float prev = (1.f - a) * (b1 * (float)r0[x0] + b2 * (float)r1[x0]);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int x1 = min(max(ix + j + 1, 0), w - 1);
            const float t = a12 * (float)r0[x1] + a22 * (float)r1[x1];
            out[i * 5 + j] = prev + t;
            prev = t * sc;
        }
    }
}

__global__ void klt_sample_uncertainty_kernel(const uint8_t* __restrict__ ref_imgs, const uint8_t* __restrict__ cur_imgs, int w, int h, int pitch,
                                              size_t stride, const float* __restrict__ ref_pts, const float* __restrict__ pts,
                                              const int* __restrict__ npts, int max_points, float* __restrict__ cov) {
    const int lane = threadIdx.x & 31, b = blockIdx.y, pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= npts[b]) return;
    const size_t o = (size_t)b * max_points + pt;
    const uint8_t* I = ref_imgs + (size_t)b * stride;
    const uint8_t* J = cur_imgs + (size_t)b * stride;
    float rd = 0.f;
    if (lane < 25) {
        float ref[25], smp[25];
        rect_subpix5(I, pitch, w, h, ref_pts[o * 2], ref_pts[o * 2 + 1], ref);
        const float du = (float)(lane / 5) * 5.f - 10.f, dv = (float)(lane % 5) * 5.f - 10.f;
        rect_subpix5(J, pitch, w, h, pts[o * 2] + du, pts[o * 2 + 1] + dv, smp);
        float ssd = 0.f;
#pragma unroll
        for (int k = 0; k < 25; ++k) { const float d = ref[k] - smp[k]; ssd += (float)pow((double)d, 2.0); }
        ssd /= 25.f;
        rd = (float)exp((double)(-0.01f * ssd));
    }
    float s = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
    for (int k = 0; k < 25; ++k) {                      // the reference's accumulation order
        const float r = __shfl_sync(0xffffffffu, rd, k);
        const float du = (float)(k / 5) * 5.f - 10.f, dv = (float)(k % 5) * 5.f - 10.f;
        s += r; sxx += r * du * du; syy += r * dv * dv; sxy += r * du * dv;
    }
    if (lane == 0) { cov[o * 4] = sxx / s; cov[o * 4 + 3] = syy / s; cov[o * 4 + 1] = sxy / s; cov[o * 4 + 2] = sxy / s; }
}

}  // namespace

namespace kltdev {

cudaError_t launch_sample_uncertainty(const uint8_t* ref_imgs, const uint8_t* cur_imgs, int w, int h, int pitch, size_t stride, int batch,
                                      const float* ref_pts, const float* pts, const int* npts, int max_points, float* cov, cudaStream_t st) {
    dim3 grid((max_points + 3) / 4, batch);
    klt_sample_uncertainty_kernel<<<grid, 128, 0, st>>>(ref_imgs, cur_imgs, w, h, pitch, stride, ref_pts, pts, npts, max_points, cov);
    return cudaGetLastError();
}

cudaError_t launch_level(const LevelJob& j0, const LevelJob& j1, int w, int h, int cpitch, size_t cstride, int dpitch, size_t dstride,
                         int npitch, size_t nstride, cudaStream_t st) {
    // a CTA walks down several tiles of a column (staging overlapped with compute) when the batch is
    // large enough to fill the GPU without splitting columns
    const int tx = (w + TW - 1) / TW, ty = (h + TH - 1) / TH, imgs = j0.batch + j1.batch;
    const long long want = 148LL * 32;
    int chunks = (int)((want + (long long)tx * imgs - 1) / ((long long)tx * imgs));
    chunks = chunks < 1 ? 1 : (chunks > ty ? ty : chunks);
    const int per = (ty + chunks - 1) / chunks;
    dim3 grid(tx, (ty + per - 1) / per, imgs);
    klt_level_kernel<<<grid, LT, 0, st>>>(j0, j1, w, h, cpitch, cstride, dpitch, dstride, npitch, nstride, per);
    return cudaGetLastError();
}

cudaError_t launch_level0_tma(const CUtensorMap& m0, const CUtensorMap& m1, const TmaJob& j0, const TmaJob& j1, int w, int h, int cpitch, size_t cstride,
                              int dpitch, size_t dstride, int npitch, size_t nstride, cudaStream_t st) {
    const int tx = (w + ATW - 1) / ATW, ty = (h + ATH - 1) / ATH, imgs = j0.batch + j1.batch;
    const long long want = 148LL * 6;              // two waves of three resident CTAs per SM; beyond that, CTAs walk down several tiles
    int chunks = (int)((want + (long long)tx * imgs - 1) / ((long long)tx * imgs));
    chunks = chunks < 1 ? 1 : (chunks > ty ? ty : chunks);
    const int per = (ty + chunks - 1) / chunks;
    dim3 grid(tx, (ty + per - 1) / per, imgs);
    klt_level0_tma_kernel<<<grid, ANT, 0, st>>>(m0, m1, j0, j1, w, h, cpitch, cstride, dpitch, dstride, npitch, nstride, per);
    return cudaGetLastError();
}

cudaError_t launch_levels_fused(const CUtensorMap& m0, const CUtensorMap& m1, const FusedJob& j0, const FusedJob& j1, const FusedLevel* lv, int nl,
                                size_t smem, cudaStream_t st) {
    static size_t configured_on[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_on[(dev >= 0 && dev < 64) ? dev : 0];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(klt_levels_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    FusedLevel z{};
    klt_levels_fused_kernel<<<j0.batch + j1.batch, FNT, smem, st>>>(m0, m1, j0, j1, lv[0], nl > 1 ? lv[1] : z, nl > 2 ? lv[2] : z, nl);
    return cudaGetLastError();
}

static int track_stride_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("EKFVIO_KLT_TRACK_STRIDE"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode;
}
size_t track_smem_bytes(int win) { return track_warp_bytes(win, track_stride_mode()) * WARPS; }

cudaError_t launch_track(const Pyr& pyr, const uint8_t* prev_slot, const uint8_t* next_slot, const float* prev_pts, float* next_pts,
                         uint8_t* status, float* err, const int* npts, int max_points, int first_image, int batch, const ekfvio_klt_params& prm,
                         cudaStream_t st, ExtLevel0 prev_ext, ExtLevel0 next_ext) {
    int mc = prm.max_iterations < 0 ? 0 : (prm.max_iterations > 100 ? 100 : prm.max_iterations);
    double eps = prm.epsilon < 0 ? 0 : (prm.epsilon > 10 ? 10 : prm.epsilon);
    eps *= eps;
    dim3 grid((max_points + WARPS - 1) / WARPS, batch);
    const size_t smem = track_smem_bytes(prm.window_size);
    if (smem > 48 * 1024) {                        // windows of 29 and 31 pixels need the opt-in shared-memory size
        static size_t configured_on[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& configured = configured_on[(dev >= 0 && dev < 64) ? dev : 0];
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(klt_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
    }
    klt_track_kernel<<<grid, WARPS * 32, smem, st>>>(pyr, prev_slot, next_slot, prev_pts, next_pts, status, err, npts,
                                                                                 max_points, prm.window_size, mc, eps, prm.min_eigen,
                                                                                 prm.use_initial_flow, first_image, track_stride_mode(), prev_ext, next_ext);
    return cudaGetLastError();
}

cudaError_t launch_postprocess(const float* next_pts, const uint8_t* status, const int* npts, const float* K9, int max_points, int batch,
                               int cols, int rows, int kill_pad, float* measured, float* cov, uint8_t* passed, cudaStream_t st) {
    dim3 grid((max_points + 127) / 128, batch);
    klt_postprocess_kernel<<<grid, 128, 0, st>>>(next_pts, status, npts, K9, max_points, cols, rows, kill_pad, measured, cov, passed);
    return cudaGetLastError();
}

}  // namespace kltdev
