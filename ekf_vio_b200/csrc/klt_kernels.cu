// KLT kernels for sm_100a: pyramid + Scharr build (HBM-bound streaming) and the per-feature
// Lucas-Kanade solve (one warp per feature, all pyramid levels in one launch).
// Arithmetic is OpenCV's calcOpticalFlowPyrLK (what the reference calls at KLTTracker.cpp:61-64):
// integer pyrDown/Scharr, Q14 bilinear weights, int16 patches, exact integer window sums,
// FP32 2x2 solve.  Compiled with --fmad=false so FP32 expressions round as OpenCV's do.
#include <float.h>

#include "klt_common.cuh"
#include "klt_kernels.h"

using namespace kltdev;

namespace {

constexpr int TW = 64, TH = 32;          // input pixels per CTA tile
constexpr int SW = TW + 8, SH = TH + 4;  // staged tile: x0-4 .. x0+TW+3, y0-2 .. y0+TH+1

// One pyramid level: reads level l of every image once and writes (a) an optional copy into the
// slot (level 0 fed from a caller buffer), (b) the interleaved int16 Scharr derivatives,
// (c) level l+1 = pyrDown(level l).  grid = (tiles_x, tiles_y, batch), 256 threads.
__global__ void __launch_bounds__(256) klt_level_kernel(const uint8_t* __restrict__ src, int spitch, size_t sstride, int w, int h,
                                                        uint8_t* __restrict__ copy_dst, int cpitch, size_t cstride,
                                                        short2* __restrict__ deriv, int dpitch, size_t dstride,
                                                        uint8_t* __restrict__ down, int npitch, size_t nstride) {
    __shared__ __align__(16) uint8_t tile[SH][SW];
    __shared__ uint16_t hs[SH][TW / 2];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const uint8_t* s = src + (size_t)b * sstride;

    // stage the tile with REFLECT_101 applied to image coordinates
    const bool interior_x = (x0 - 4 >= 0) && (x0 + TW + 4 <= w) && ((spitch & 3) == 0) && ((((size_t)s) & 3) == 0);
    if (interior_x) {
        for (int e = tid; e < SH * (SW / 4); e += 256) {
            int r = e / (SW / 4), wc = e % (SW / 4);
            int sy = reflect101(y0 - 2 + r, h);
            uint32_t v = *reinterpret_cast<const uint32_t*>(s + (size_t)sy * spitch + (x0 - 4) + wc * 4);
            *reinterpret_cast<uint32_t*>(&tile[r][wc * 4]) = v;
        }
    } else {
        for (int e = tid; e < SH * SW; e += 256) {
            int r = e / SW, c = e % SW;
            int sy = reflect101(y0 - 2 + r, h), sx = reflect101(x0 - 4 + c, w);
            tile[r][c] = s[(size_t)sy * spitch + sx];
        }
    }
    __syncthreads();

    if (copy_dst) {
        uint8_t* cd = copy_dst + (size_t)b * cstride;
        for (int e = tid; e < TH * (TW / 4); e += 256) {
            int r = e / (TW / 4), wc = e % (TW / 4);
            int y = y0 + r, x = x0 + wc * 4;
            if (y < h && x < w)  // pitch is a multiple of 16: whole words stay inside the row
                *reinterpret_cast<uint32_t*>(cd + (size_t)y * cpitch + x) = *reinterpret_cast<const uint32_t*>(&tile[r + 2][wc * 4 + 4]);
        }
    }

    if (deriv) {
        short2* dd = deriv + (size_t)b * dstride;
        for (int g = tid; g < TH * (TW / 4); g += 256) {
            int r = g / (TW / 4), xg = (g % (TW / 4)) * 4;
            int y = y0 + r, x = x0 + xg;
            if (y >= h || x >= w) continue;
            // columns c0-1 .. c0+4 of rows r+1, r+2, r+3 (tile coordinates), c0 = 4 + xg
            int sm_[6], dm_[6];
            const uint8_t* t0 = &tile[r + 1][xg + 3];
            const uint8_t* t1 = &tile[r + 2][xg + 3];
            const uint8_t* t2 = &tile[r + 3][xg + 3];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                int a = t0[c], m = t1[c], z = t2[c];
                sm_[c] = 3 * a + 10 * m + 3 * z;
                dm_[c] = z - a;
            }
            short2 o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[q].x = (short)(sm_[q + 2] - sm_[q]);
                o[q].y = (short)(3 * dm_[q] + 10 * dm_[q + 1] + 3 * dm_[q + 2]);
            }
            short2* out = dd + (size_t)y * dpitch + x;
            if (x + 3 < w) {
                uint4 v;
                v.x = (uint32_t)(uint16_t)o[0].x | ((uint32_t)(uint16_t)o[0].y << 16);
                v.y = (uint32_t)(uint16_t)o[1].x | ((uint32_t)(uint16_t)o[1].y << 16);
                v.z = (uint32_t)(uint16_t)o[2].x | ((uint32_t)(uint16_t)o[2].y << 16);
                v.w = (uint32_t)(uint16_t)o[3].x | ((uint32_t)(uint16_t)o[3].y << 16);
                *reinterpret_cast<uint4*>(out) = v;
            } else {
                for (int q = 0; q < 4 && x + q < w; ++q) out[q] = o[q];
            }
        }
    }

    if (down) {
        const int dw = (w + 1) / 2, dh = (h + 1) / 2;
        // horizontal [1 4 6 4 1] on every staged row
        for (int e = tid; e < SH * (TW / 2); e += 256) {
            int r = e / (TW / 2), ox = e % (TW / 2);
            const uint8_t* t = &tile[r][2 + 2 * ox];
            hs[r][ox] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
        }
        __syncthreads();
        uint8_t* nd = down + (size_t)b * nstride;
        for (int e = tid; e < (TH / 2) * (TW / 2); e += 256) {
            int oy = e / (TW / 2), ox = e % (TW / 2);
            int gx = x0 / 2 + ox, gy = y0 / 2 + oy;
            if (gx < dw && gy < dh) {
                int sum = hs[2 * oy][ox] + 4 * hs[2 * oy + 1][ox] + 6 * hs[2 * oy + 2][ox] + 4 * hs[2 * oy + 3][ox] + hs[2 * oy + 4][ox];
                nd[(size_t)gy * npitch + gx] = (uint8_t)((sum + 128) >> 8);
            }
        }
    }
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    w00 = __float2int_rn((1.f - a) * (1.f - b) * (float)(1 << 14));
    w01 = __float2int_rn(a * (1.f - b) * (float)(1 << 14));
    w10 = __float2int_rn((1.f - a) * b * (float)(1 << 14));
    w11 = (1 << 14) - w00 - w01 - w10;
}

constexpr int WARPS = 4;      // features per CTA

// LKTrackerInvoker for every level, one warp per point.  Per-warp shared memory:
//   Ipatch[win*win] int16, dI[win*win] short2, Jt[(win+1)^2] u8.
__global__ void __launch_bounds__(WARPS * 32) klt_track_kernel(Pyr pyr, const uint8_t* __restrict__ prev_slot, const uint8_t* __restrict__ next_slot,
                                                               const float* __restrict__ prev_pts, float* __restrict__ next_pts,
                                                               uint8_t* __restrict__ status, float* __restrict__ err, const int* __restrict__ npts,
                                                               int max_points, int win, int max_count, double epsilon, double min_eig,
                                                               int use_initial_flow) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int pt = blockIdx.x * WARPS + warp;
    if (pt >= npts[b]) return;
    const int area = win * win, jw1 = win + 1;
    const size_t per_warp = (size_t)area * 2 + (size_t)area * 4 + (((size_t)jw1 * jw1 + 15) & ~(size_t)15);
    uint8_t* base = smem_raw + warp * ((per_warp + 15) & ~(size_t)15);
    short2* dI = reinterpret_cast<short2*>(base);
    short* Ip = reinterpret_cast<short*>(base + (size_t)area * 4);
    uint8_t* Jt = base + (size_t)area * 6;

    const size_t pidx = ((size_t)b * max_points + pt) * 2;
    const float prevx = prev_pts[pidx], prevy = prev_pts[pidx + 1];
    float stx = use_initial_flow ? next_pts[pidx] : prevx, sty = use_initial_flow ? next_pts[pidx + 1] : prevy;  // stored nextPts
    int st = 1;
    float errv = 0.f;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = pyr.levels - 1;

    for (int level = top; level >= 0; --level) {
        const Level L = pyr.lv[level];
        const uint8_t* I = prev_slot + L.img_off + (size_t)b * L.img_stride;
        const uint8_t* J = next_slot + L.img_off + (size_t)b * L.img_stride;
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
        long long a11 = 0, a12 = 0, a22 = 0;
        for (int e = lane; e < area; e += 32) {
            int y = e / win, x = e - y * win;
            int X = ipx + x, Y = ipy + y;
            int X0 = reflect101(X, L.w), X1 = reflect101(X + 1, L.w), Y0 = reflect101(Y, L.h), Y1 = reflect101(Y + 1, L.h);
            int i00 = I[(size_t)Y0 * L.pitch + X0], i01 = I[(size_t)Y0 * L.pitch + X1];
            int i10 = I[(size_t)Y1 * L.pitch + X0], i11 = I[(size_t)Y1 * L.pitch + X1];
            int ival = (i00 * w00 + i01 * w01 + i10 * w10 + i11 * w11 + (1 << 8)) >> 9;
            short2 z2 = make_short2(0, 0);
            bool xin0 = (unsigned)X < (unsigned)L.w, xin1 = (unsigned)(X + 1) < (unsigned)L.w;
            bool yin0 = (unsigned)Y < (unsigned)L.h, yin1 = (unsigned)(Y + 1) < (unsigned)L.h;
            short2 d00 = (xin0 && yin0) ? dIm[(size_t)Y * L.dpitch + X] : z2;
            short2 d01 = (xin1 && yin0) ? dIm[(size_t)Y * L.dpitch + X + 1] : z2;
            short2 d10 = (xin0 && yin1) ? dIm[(size_t)(Y + 1) * L.dpitch + X] : z2;
            short2 d11 = (xin1 && yin1) ? dIm[(size_t)(Y + 1) * L.dpitch + X + 1] : z2;
            int ixval = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
            int iyval = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
            Ip[e] = (short)ival;
            dI[e] = make_short2((short)ixval, (short)iyval);
            a11 += ixval * ixval; a12 += ixval * iyval; a22 += iyval * iyval;
        }
        a11 = warp_sum_ll(a11); a12 = warp_sum_ll(a12); a22 = warp_sum_ll(a22);
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
        __syncwarp();
        for (int j = 0; j < max_count; ++j) {
            int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            lk_weights(nx - inx, ny - iny, w00, w01, w10, w11);
            for (int e = lane; e < jw1 * jw1; e += 32) {
                int y = e / jw1, x = e - y * jw1;
                Jt[e] = J[(size_t)reflect101(iny + y, L.h) * L.pitch + reflect101(inx + x, L.w)];
            }
            __syncwarp();
            long long b1 = 0, b2 = 0;
            for (int e = lane; e < area; e += 32) {
                int y = e / win, x = e - y * win;
                const uint8_t* jp = Jt + y * jw1 + x;
                int diff = ((jp[0] * w00 + jp[1] * w01 + jp[jw1] * w10 + jp[jw1 + 1] * w11 + (1 << 8)) >> 9) - Ip[e];
                short2 d = dI[e];
                b1 += diff * d.x; b2 += diff * d.y;
            }
            b1 = warp_sum_ll(b1); b2 = warp_sum_ll(b2);
            __syncwarp();
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
                __syncwarp();
                for (int e = lane; e < jw1 * jw1; e += 32) {
                    int y = e / jw1, x = e - y * jw1;
                    Jt[e] = J[(size_t)reflect101(iny + y, L.h) * L.pitch + reflect101(inx + x, L.w)];
                }
                __syncwarp();
                long long es = 0;
                for (int e = lane; e < area; e += 32) {
                    int y = e / win, x = e - y * win;
                    const uint8_t* jp = Jt + y * jw1 + x;
                    int diff = ((jp[0] * w00 + jp[1] * w01 + jp[jw1] * w10 + jp[jw1 + 1] * w11 + (1 << 8)) >> 9) - Ip[e];
                    es += diff < 0 ? -diff : diff;
                }
                es = warp_sum_ll(es);
                errv = (float)es * 1.f / (float)(32 * win * win);
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

}  // namespace

namespace kltdev {

cudaError_t launch_level(const uint8_t* src, int spitch, size_t sstride, int w, int h, uint8_t* copy_dst, int cpitch, size_t cstride,
                         short2* deriv, int dpitch, size_t dstride, uint8_t* down, int npitch, size_t nstride, int batch, cudaStream_t st) {
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, batch);
    klt_level_kernel<<<grid, 256, 0, st>>>(src, spitch, sstride, w, h, copy_dst, cpitch, cstride, deriv, dpitch, dstride, down, npitch, nstride);
    return cudaGetLastError();
}

size_t track_smem_bytes(int win) {
    size_t area = (size_t)win * win, jw1 = (size_t)win + 1;
    size_t per_warp = area * 2 + area * 4 + ((jw1 * jw1 + 15) & ~(size_t)15);
    per_warp = (per_warp + 15) & ~(size_t)15;
    return per_warp * WARPS;
}

cudaError_t launch_track(const Pyr& pyr, const uint8_t* prev_slot, const uint8_t* next_slot, const float* prev_pts, float* next_pts,
                         uint8_t* status, float* err, const int* npts, int max_points, int batch, const ekfvio_klt_params& prm, cudaStream_t st) {
    int mc = prm.max_iterations < 0 ? 0 : (prm.max_iterations > 100 ? 100 : prm.max_iterations);
    double eps = prm.epsilon < 0 ? 0 : (prm.epsilon > 10 ? 10 : prm.epsilon);
    eps *= eps;
    dim3 grid((max_points + WARPS - 1) / WARPS, batch);
    klt_track_kernel<<<grid, WARPS * 32, track_smem_bytes(prm.window_size), st>>>(pyr, prev_slot, next_slot, prev_pts, next_pts, status, err, npts,
                                                                                 max_points, prm.window_size, mc, eps, prm.min_eigen,
                                                                                 prm.use_initial_flow);
    return cudaGetLastError();
}

cudaError_t launch_postprocess(const float* next_pts, const uint8_t* status, const int* npts, const float* K9, int max_points, int batch,
                               int cols, int rows, int kill_pad, float* measured, float* cov, uint8_t* passed, cudaStream_t st) {
    dim3 grid((max_points + 127) / 128, batch);
    klt_postprocess_kernel<<<grid, 128, 0, st>>>(next_pts, status, npts, K9, max_points, cols, rows, kill_pad, measured, cov, passed);
    return cudaGetLastError();
}

}  // namespace kltdev
