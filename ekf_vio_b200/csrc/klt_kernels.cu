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

constexpr int TW = 64, TH = 32;            // input pixels per CTA tile
constexpr int SW = TW + 16, SH = TH + 4;   // staged tile: x0-8 .. x0+TW+7 (16-byte aligned rows), y0-2 .. y0+TH+1

__device__ __forceinline__ int reflect_once(int p, int len) { return p < 0 ? -p : (p >= len ? 2 * len - 2 - p : p); }

// One pyramid level: reads level l of every image once and writes (a) an optional copy into the
// slot (level 0 fed from a caller buffer), (b) the interleaved int16 Scharr derivatives,
// (c) level l+1 = pyrDown(level l).  grid = (tiles_x, tiles_y, images of both jobs), 256 threads.
// Integer work is done with u8x4 dot products (IDP4A) on 32-bit windows of the staged tile, one
// 16-byte vector store per 4 output pixels; the kernel is meant to be bound by HBM, not by issue.
__global__ void __launch_bounds__(256) klt_level_kernel(LevelJob j0, LevelJob j1, int w, int h, int cpitch, size_t cstride, int dpitch,
                                                        size_t dstride, int npitch, size_t nstride) {
    __shared__ __align__(16) uint8_t tile[SH][SW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
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

    // stage the tile with REFLECT_101 applied to image coordinates (one reflection suffices: the
    // halo is at most 8 pixels and every level is wider than the 21-pixel window)
    const bool aligned8 = ((spitch & 7) == 0) && ((((size_t)s) & 7) == 0);
    if (aligned8 && x0 >= 8 && x0 + TW + 8 <= w) {
        for (int e = tid; e < SH * (SW / 8); e += 256) {
            const int r = e / (SW / 8), v = e % (SW / 8);
            const int sy = reflect_once(y0 - 2 + r, h);
            *reinterpret_cast<uint2*>(&tile[r][v * 8]) = *reinterpret_cast<const uint2*>(s + (size_t)sy * spitch + (x0 - 8) + v * 8);
        }
    } else {
        const bool aligned4 = ((spitch & 3) == 0) && ((((size_t)s) & 3) == 0);
        for (int e = tid; e < SH * (SW / 4); e += 256) {
            const int r = e / (SW / 4), wc = e % (SW / 4);
            const int sy = reflect_once(y0 - 2 + r, h), xs = x0 - 8 + wc * 4;
            const uint8_t* row = s + (size_t)sy * spitch;
            uint32_t v;
            if (aligned4 && xs >= 0 && xs + 3 < w) v = *reinterpret_cast<const uint32_t*>(row + xs);
            else v = (uint32_t)row[reflect101(xs, w)] | ((uint32_t)row[reflect101(xs + 1, w)] << 8) | ((uint32_t)row[reflect101(xs + 2, w)] << 16) |
                     ((uint32_t)row[reflect101(xs + 3, w)] << 24);
            *reinterpret_cast<uint32_t*>(&tile[r][wc * 4]) = v;
        }
    }
    __syncthreads();

    if (copy_dst) {   // 32 rows x 8 eight-byte vectors = one per thread
        const int r = tid >> 3, v = tid & 7;
        const int y = y0 + r, x = x0 + v * 8;
        if (y < h && x < w)   // the slot's pitch is a multiple of 16: whole vectors stay inside the row
            *reinterpret_cast<uint2*>(copy_dst + (size_t)b * cstride + (size_t)y * cpitch + x) = *reinterpret_cast<const uint2*>(&tile[r + 2][v * 8 + 8]);
    }

    if (deriv) {
        // Scharr: for pixel x the window bytes (x-1, x, x+1, x+2) of rows y-1, y, y+1 are dotted with the
        // packed 3x3 weights; one thread = 8 pixels of a row.
        const int r = tid >> 3, xg = (tid & 7) * 8;
        const int y = y0 + r, x = x0 + xg;
        if (y < h && x < w) {
            unsigned o[8];
            int ix[8], iy[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { ix[j] = 0; iy[j] = 0; }
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const unsigned* tw = reinterpret_cast<const unsigned*>(&tile[r + 1 + rr][xg + 4]);   // word 0 = pixels xg-4 .. xg-1
                const unsigned w0 = tw[0], w1 = tw[1], w2 = tw[2], w3 = tw[3];
                unsigned win[8];
                win[0] = __funnelshift_r(w0, w1, 24);
                win[1] = w1;
                win[2] = __funnelshift_r(w1, w2, 8);
                win[3] = __funnelshift_r(w1, w2, 16);
                win[4] = __funnelshift_r(w1, w2, 24);
                win[5] = w2;
                win[6] = __funnelshift_r(w2, w3, 8);
                win[7] = __funnelshift_r(w2, w3, 16);
                const int wx = (rr == 1) ? 0x000A00F6 : 0x000300FD;                   // (-10,0,10,0) / (-3,0,3,0)
                const int wy = (rr == 0) ? 0x00FDF6FD : 0x00030A03;                   // (-3,-10,-3,0) / (3,10,3,0)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    ix[j] = dp4a_us(win[j], wx, ix[j]);
                    if (rr != 1) iy[j] = dp4a_us(win[j], wy, iy[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __byte_perm((unsigned)ix[j], (unsigned)iy[j], 0x5410);   // short2(ix, iy)
            short2* out = deriv + (size_t)b * dstride + (size_t)y * dpitch + x;
            if (x + 7 < w) {
                reinterpret_cast<uint4*>(out)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4*>(out)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (x + j < w) reinterpret_cast<unsigned*>(out)[j] = o[j];
            }
        }
    }

    if (down) {
        // pyrDown: thread = two horizontally adjacent outputs; [1 4 6 4 1] across as (1,4,6,4).window + 5th
        // byte, then [1 4 6 4 1] down the five rows, all in registers.
        const int dw = (w + 1) / 2, dh = (h + 1) / 2;
        const int oy = tid >> 4, oxp = (tid & 15) * 2;
        const int gx = x0 / 2 + oxp, gy = y0 / 2 + oy;
        if (gx < dw && gy < dh) {
            int sa = 0, sb = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const unsigned* tw = reinterpret_cast<const unsigned*>(&tile[2 * oy + j][2 * oxp + 4]);   // bytes base+4 ..
                const unsigned w1 = tw[0], w2 = tw[1], w3 = tw[2];
                const unsigned ha = dp4a_uu(__funnelshift_r(w1, w2, 16), 0x04060401u, (w2 >> 16) & 0xffu);  // taps base+6 .. base+10
                const unsigned hb = dp4a_uu(w2, 0x04060401u, w3 & 0xffu);                                    // taps base+8 .. base+12
                const int wv = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
                sa += wv * (int)ha; sb += wv * (int)hb;
            }
            uint8_t* nd = down + (size_t)b * nstride + (size_t)gy * npitch + gx;
            nd[0] = (uint8_t)((sa + 128) >> 8);
            if (gx + 1 < dw) nd[1] = (uint8_t)((sb + 128) >> 8);
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

__host__ __device__ inline size_t track_warp_bytes(int win) {
    size_t area = (size_t)win * win, t = ((size_t)win + 1) * (win + 1);
    size_t bytes = area * 4 /*dI*/ + t * 4 /*Dt*/ + ((area * 2 + 3) & ~(size_t)3) /*Ip*/ + ((t + 3) & ~(size_t)3) /*Jt*/;
    return (bytes + 15) & ~(size_t)15;
}

// LKTrackerInvoker for every level, one warp per point.  Per-warp shared memory:
//   dI[win*win] short2 (Ix, Iy of the patch), Dt[(win+1)^2] short2 (staged derivative tile),
//   Ip[win*win] int16 (patch intensities, 5 fractional bits), Jt[(win+1)^2] u8 (staged I or J tile).
// Tiles are staged row by row (lanes = columns), so every bilinear tap is a shared-memory read and
// no per-pixel integer division or border test remains in the loops.
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
    const int area = win * win, jw1 = win + 1, tarea = jw1 * jw1;
    uint8_t* base = smem_raw + warp * track_warp_bytes(win);
    short2* dI = reinterpret_cast<short2*>(base);
    short2* Dt = dI + area;
    short* Ip = reinterpret_cast<short*>(Dt + tarea);
    uint8_t* Jt = reinterpret_cast<uint8_t*>(Ip) + ((area * 2 + 3) & ~3);
    // lane's pixel slots: e = lane, lane + 32, ...  ->  (y, x) advanced without divisions
    const int y_first = lane / win, x_first = lane - y_first * win;
    const int qstep = 32 / win, rstep = 32 - qstep * win;

    const size_t pidx = ((size_t)b * max_points + pt) * 2;
    const float prevx = prev_pts[pidx], prevy = prev_pts[pidx + 1];
    float stx = use_initial_flow ? next_pts[pidx] : prevx, sty = use_initial_flow ? next_pts[pidx + 1] : prevy;  // stored nextPts
    int st = 1;
    float errv = 0.f;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = pyr.levels - 1;

    // stage a (win+1)^2 tile of an 8-bit image at (ox, oy) with REFLECT_101 borders
    auto stage_u8 = [&](const uint8_t* img, const Level& L, int ox, int oy) {
        if (lane < jw1) {
            const bool inside = ox >= 0 && oy >= 0 && ox + win < L.w && oy + win < L.h;
            if (inside) {
                const uint8_t* p = img + (size_t)oy * L.pitch + ox + lane;
                for (int y = 0; y < jw1; ++y) Jt[y * jw1 + lane] = p[(size_t)y * L.pitch];
            } else {
                const int cx = reflect101(ox + lane, L.w);
                for (int y = 0; y < jw1; ++y) Jt[y * jw1 + lane] = img[(size_t)reflect101(oy + y, L.h) * L.pitch + cx];
            }
        }
        __syncwarp();
    };

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
        // stage the I tile (into Jt) and the derivative tile (zero outside the image)
        __syncwarp();
        stage_u8(I, L, ipx, ipy);
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
        for (int e = lane, y = y_first, x = x_first; e < area; e += 32) {
            const int o = y * jw1 + x;
            const int ival = (Jt[o] * w00 + Jt[o + 1] * w01 + Jt[o + jw1] * w10 + Jt[o + jw1 + 1] * w11 + (1 << 8)) >> 9;
            const short2 d00 = Dt[o], d01 = Dt[o + 1], d10 = Dt[o + jw1], d11 = Dt[o + jw1 + 1];
            const int ixval = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
            const int iyval = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
            Ip[e] = (short)ival;
            dI[e] = make_short2((short)ixval, (short)iyval);
            sa11 += ixval * ixval; sa12 += ixval * iyval; sa22 += iyval * iyval;
            y += qstep; x += rstep;
            if (x >= win) { x -= win; ++y; }
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
        for (int j = 0; j < max_count; ++j) {
            int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            lk_weights(nx - inx, ny - iny, w00, w01, w10, w11);
            __syncwarp();
            stage_u8(J, L, inx, iny);
            int sb1 = 0, sb2 = 0;                 // <= 31 slots x 8160 x 4080 fits 32 bits
            for (int e = lane, y = y_first, x = x_first; e < area; e += 32) {
                const uint8_t* jp = Jt + y * jw1 + x;
                const int diff = ((jp[0] * w00 + jp[1] * w01 + jp[jw1] * w10 + jp[jw1 + 1] * w11 + (1 << 8)) >> 9) - Ip[e];
                const short2 d = dI[e];
                sb1 += diff * d.x; sb2 += diff * d.y;
                y += qstep; x += rstep;
                if (x >= win) { x -= win; ++y; }
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
                __syncwarp();
                stage_u8(J, L, inx, iny);
                long long es = 0;
                for (int e = lane, y = y_first, x = x_first; e < area; e += 32) {
                    const uint8_t* jp = Jt + y * jw1 + x;
                    const int diff = ((jp[0] * w00 + jp[1] * w01 + jp[jw1] * w10 + jp[jw1 + 1] * w11 + (1 << 8)) >> 9) - Ip[e];
                    es += diff < 0 ? -diff : diff;
                    y += qstep; x += rstep;
                    if (x >= win) { x -= win; ++y; }
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

cudaError_t launch_level(const LevelJob& j0, const LevelJob& j1, int w, int h, int cpitch, size_t cstride, int dpitch, size_t dstride,
                         int npitch, size_t nstride, cudaStream_t st) {
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, j0.batch + j1.batch);
    klt_level_kernel<<<grid, 256, 0, st>>>(j0, j1, w, h, cpitch, cstride, dpitch, dstride, npitch, nstride);
    return cudaGetLastError();
}

size_t track_smem_bytes(int win) { return track_warp_bytes(win) * WARPS; }

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
