// ekf_update_fused: the whole measurement update of a symmetric filter in ONE kernel with Sigma resident in registers
// (updateWithFeaturePositions, TightlyCoupledEKF.cpp:475-628, for the filters the reduced form Sigma - Z Z' is valid for —
// ekf_kernels.h ROUTE_SYM).
//
// R is block diagonal (one 2x2 block per feature, :538-557) and H is a 0/1 selection (:634-661), so the batch update with all m
// measurements equals the measurement blocks applied one after the other.  With the state re-ordered so that the measured rows
// come first (a permutation applied while Sigma is loaded and undone while it is stored) this is a right-looking elimination on
// 8x8 tiles: for block j
//      S_j  = Sigma(j,j) + R_j = L_j L_j'                 (8x8; its pivots ARE the pivots of the Cholesky factor of the full S)
//      Z_j  = Sigma(:,j) inv(L_j)'                         (N x 8 panel)
//      v_j  = inv(L_j) (y_j - dmu(j)),  dmu += Z_j v_j     (K y of :600, block by block)
//      Sigma -= Z_j Z_j'                                   (rank-8 update of all 253 lower-triangular tiles: 2 DMMA per tile)
// The forward substitution of the batch form (N m^2 FLOPs) and the separate factorisation of S (m^3/3) disappear: the columns
// Sigma(:,idx) are part of the matrix the rank-8 updates act on anyway.  Sigma is read once (its lower triangle) and written once;
// the lower triangle lives in the accumulator registers of 16 warps (<= 18 tiles = 72 registers per thread) through all 13 steps.
//
// Warp roles (tile units; 22 tile rows cover N <= 176):
//   warps 0..9   one full 4x4-tile block (R0, C0) below the diagonal of the 5x5 grid of super blocks that covers rows 0..19
//   warps 10..14 the lower triangle of diagonal super block d (10 tiles) + the 2x4 tiles of rows 20, 21 under it
//   warp 15      tiles (20,20) (21,20) (21,21), and the serial work: it keeps private copies of the measurement diagonal tiles in
//                shared memory, applies the rank-8 updates to them itself and factors tile j+1 WHILE the other warps run the
//                rank-8 update of step j (look-ahead), so the 8x8 factorisation is off the critical path.
// A filter whose S turns out not positive definite or beyond the pivot ratio of ILLCOND_RATIO is abandoned here (nothing has
// been written) and served by the kernels of ekf_tiled.cu, which decide its route as before.
#include "ekf_common.cuh"
#include "ekf_kernels.h"
#include "ekf_tiles.cuh"

using namespace ekfvio;

namespace {

constexpr int NTR = 22;          // tile rows
constexpr int NBM = 13;          // measurement blocks (m <= 104)
constexpr int FW = 16;           // warps

#ifdef EKFVIO_PROFILE_CLOCKS
__device__ unsigned long long g_fclk[16];
#define FCLK(i) do { if (lane == 0 && (warp == 0 || warp == 15)) { long long t_ = clock64(); atomicAdd(&g_fclk[(i) + (warp == 15 ? 8 : 0)], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define FCLK(i) do {} while (0)
#endif

__device__ __forceinline__ int tri(int ri, int cj) { return ri * (ri + 1) / 2 + cj; }

// Staging buffer of the result (dynamic shared memory): Sigma' in its own (un-permuted) order, rows 0..21 complete, feature row a
// up to column a + 2 or a + 3 (covers the 3x3 diagonal block of its feature), rows 16-byte aligned.
__device__ __forceinline__ int orow_len(int a) { return a < BASE ? NTR * 8 : ((a + 4) & ~1); }
constexpr int OB_DOUBLES = BASE * NTR * 8 + 15708 + 8;     // 22 full rows + sum of orow_len(22..175) (+ slack)

// element pair (pr, pc), (pr, pc+1) of the permuted matrix, read through the lower triangle of the stored one.  Permuted row N
// (the first spare row) carries the innovation: y over the measured columns, zero elsewhere — the rank-8 updates then turn it
// into y - dmu over the measured and -dmu over the other columns, i.e. K y comes out of the elimination itself.
__device__ __forceinline__ double2 load_pair(const double* __restrict__ Pi, int ld, int N, int m, const int* perm, const double* sy, int pr, int pc) {
    double2 v = make_double2(0.0, 0.0);
    if (pr < N) {
        const int a = perm[pr];
        if (pc < N) { const int b = perm[pc]; v.x = Pi[(size_t)max(a, b) * ld + min(a, b)]; }
        if (pc + 1 < N) { const int b = perm[pc + 1]; v.y = Pi[(size_t)max(a, b) * ld + min(a, b)]; }
    } else if (pr == N) {
        if (pc < m) v.x = sy[pc];
        if (pc + 1 < m) v.y = sy[pc + 1];
    }
    return v;
}

// Element (pr, pc) of the permuted result into the staging buffer: its place in the lower triangle of the stored matrix, plus the
// mirror image where the row of the smaller index extends that far (rows 0..21 complete, the 3x3 diagonal block of a feature).
// Row N hands dmu to s_dl.
__device__ __forceinline__ void stage_elem(double* Ob, const int* ooff, double* s_dl, const double* sy, int N, int m, const int* perm,
                                           int pr, int pc, double v) {
    if (pc > pr || pc >= N) return;
    if (pr >= N) {
        if (pr == N) s_dl[pc] = ((pc < m) ? sy[pc] : 0.0) - v;
        return;
    }
    const int a = perm[pr], b = perm[pc];
    const int hi = max(a, b), lo = min(a, b);
    v = prune(v);
    Ob[ooff[hi] + lo] = v;
    if (hi != lo && hi < orow_len(lo)) Ob[ooff[lo] + hi] = v;
}

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

// private copy of a measurement diagonal tile: T -= Z_I Z_I'
__device__ __forceinline__ void diag_update(double* T, const double* Zt, int lane) {
    const int r = lane >> 2, q = lane & 3;
    double2 c = *reinterpret_cast<double2*>(&T[tsw(r, 2 * q)]);
    const double z0 = Zt[tsw(r, q)], z1 = Zt[tsw(r, 4 + q)];
    dmma884(c.x, c.y, -z0, z0);
    dmma884(c.x, c.y, -z1, z1);
    *reinterpret_cast<double2*>(&T[tsw(r, 2 * q)]) = c;
}

// Z tile of a panel from a register tile of column j (direct) or row j (transposed): Z = T inv(L)' resp. T' inv(L)'.
__device__ __forceinline__ void z_tile(double* Zt, double c0, double c1, bool transposed, int lane, int mrem, double lb0, double lb1) {
    const int r = lane >> 2, q = lane & 3;
    double a0, a1;
    if (!transposed) {
        cfrag_to_afrag(c0, c1, lane, a0, a1);
    } else {          // A[r][k] = T[k][r]: held by lane (k, r/2), component r&1
        const int s0 = q * 4 + (r >> 1), s1 = (q + 4) * 4 + (r >> 1);
        double v0 = __shfl_sync(0xffffffffu, c0, s0), v1 = __shfl_sync(0xffffffffu, c1, s0);
        a0 = (r & 1) ? v1 : v0;
        v0 = __shfl_sync(0xffffffffu, c0, s1); v1 = __shfl_sync(0xffffffffu, c1, s1);
        a1 = (r & 1) ? v1 : v0;
    }
    if (q >= mrem) a0 = 0.0;             // columns of the block beyond m are not measurements
    if (q + 4 >= mrem) a1 = 0.0;
    double t0 = 0.0, t1 = 0.0;
    dmma884(t0, t1, a0, lb0);
    dmma884(t0, t1, a1, lb1);
    *reinterpret_cast<double2*>(&Zt[tsw(r, 2 * q)]) = make_double2(t0, t1);
}

// S_j = Dt + R_j (lower triangle), factored S_j = L L' by warp 15; inv(L) to Li (tsw layout).  Rows of the block beyond m are
// replaced by identity rows.  Returns false (on every lane) for a pivot that is not positive or breaks the pivot ratio.
// Elimination in LDL^T form (the reference's unpivoted SimplicialLDLT order, :577-580): one reciprocal per column on the
// critical path; the square roots only scale the finished columns.
__device__ __forceinline__ bool factor_block(const double* Dt, double* Li, const double* sR, int j, int m,
                                             double illcond, double& dmin, double& dmax, int lane) {
    const int rr = lane & 7;
    const bool meas = 8 * j + rr < m;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        a[c] = 0.0;
        if (c <= rr) {
            if (meas) {
                a[c] = Dt[tsw(rr, c)];
                if ((c >> 1) == (rr >> 1)) a[c] += sR[4 * (4 * j + (rr >> 1)) + (c & 1) * 2 + (rr & 1)];      // upper(S)(c, rr), :559-561, :578
            } else if (c == rr) a[c] = 1.0;
        }
    }
    bool ok = true;
    double rs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);            // pivot
        double l[8];
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) l[c2] = __shfl_sync(0xffffffffu, a[c], c2);
        if (!(d > 0.0)) ok = false;
        if (8 * j + c < m) { dmax = fmax(dmax, d); dmin = fmin(dmin, d); }
        const double t = a[c] * __drcp_rn(d);
        rs[c] = rsqrt(d);
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2)
            if (rr >= c2) a[c2] -= t * l[c2];
    }
    if (!(dmax <= illcond * dmin)) ok = false;
    // L = L_ldl |D|^1/2 (strictly lower part; the diagonal enters through rs = 1 / L(c,c))
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (c < rr) ? a[c] * rs[c] : 0.0;
    // column rr of inv(L) by forward substitution, rows fetched by shuffle
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double sacc = (i == rr) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) sacc -= __shfl_sync(0xffffffffu, a[k], i) * x[k];
        x[i] = sacc * rs[i];
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Li[tsw(i, rr)] = (i >= rr) ? x[i] : 0.0;
    }
    return ok;
}

__global__ void __launch_bounds__(FW * 32, 1) ekf_update_fused(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout,
                                                               const double* __restrict__ z, const double* __restrict__ Rin,
                                                               const uint8_t* __restrict__ pass) {
    extern __shared__ __align__(16) double Ob[];      // staging buffer of the result (OB_DOUBLES)
    __shared__ __align__(16) double Zs[NTR * 64];     // the panel Z_j, one swizzled 8x8 tile per tile row
    __shared__ __align__(16) double Dg[NBM * 64];     // private copies of the measurement diagonal tiles (look-ahead factorisation)
    __shared__ __align__(16) double Li[64];           // inv(L_j)
    __shared__ double s_y[NBM * 8], s_dl[NTR * 8], s_x[NTR * 8], s_R[NBM * 16];
    __shared__ int s_perm[NTR * 8], s_ooff[NTR * 8];
    __shared__ int s_m, s_elig, s_abort;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int ld = p.ldP, nmax = p.nmax;

    // role of this warp
    int R0 = 0, C0 = 0;          // type F: first tile row / column of the 4x4 block; type D: R0 = C0 = 4 d
    if (warp < 10) {
        int sr = 1, w = warp;
        while (w >= sr) { w -= sr; ++sr; }
        R0 = 4 * sr; C0 = 4 * w;
    } else if (warp < 15) {
        R0 = C0 = 4 * (warp - 10);
    }
    const bool typeF = warp < 10;
    if (warp == 0) {             // row offsets of the staging buffer
        int o = 0;
        for (int a0 = 0; a0 < NTR * 8; a0 += 32) {
            const int a = a0 + lane;
            int len = orow_len(a), inc = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            s_ooff[a] = o + inc - len;
            o += __shfl_sync(0xffffffffu, inc, 31);
        }
    }

    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const int n = p.nfeat[f], N = BASE + 3 * n;
        const double* Pi = Pin + (size_t)f * ld * ld;
        double* Po = Pout + (size_t)f * ld * ld;
        double* mu_g = p.mu + (size_t)f * BASE;
        double* feat_g = p.feat + (size_t)f * nmax * 3;

        __syncthreads();                                   // shared memory of the previous filter is free
#ifdef EKFVIO_PROFILE_CLOCKS
        long long t_prev = clock64();
#endif
        if (warp == 0) {
            // formFeatureMeasurementMap (:634-661) + bookkeeping of :506-529 for features lane and lane + 32, all loads issued
            // at once; then the permutation: measured rows in measurement order, the other state rows behind them in their own order
            const double* zf = z + (size_t)f * nmax * 2;
            const double* Rf = Rin + (size_t)f * nmax * 4;
            const uint8_t* pf = pass + (size_t)f * nmax;
            bool pr[2]; double zx[2], zy[2], fx[2], fy[2], rv[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                pr[h] = false; zx[h] = zy[h] = fx[h] = fy[h] = 0.0; rv[h][0] = rv[h][1] = rv[h][2] = rv[h][3] = 0.0;
                if (i < n) {
                    pr[h] = pf[i] != 0;
                    zx[h] = zf[2 * i]; zy[h] = zf[2 * i + 1];
                    fx[h] = feat_g[3 * i]; fy[h] = feat_g[3 * i + 1];
#pragma unroll
                    for (int k = 0; k < 4; ++k) rv[h][k] = Rf[4 * i + k];
                }
            }
            const unsigned mk0 = __ballot_sync(0xffffffffu, pr[0]), mk1 = __ballot_sync(0xffffffffu, pr[1]);
            const unsigned lt = (1u << lane) - 1u;
            const int m = 2 * (__popc(mk0) + __popc(mk1));
            bool elig = p.asym[f] == 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                const int pm = h == 0 ? __popc(mk0 & lt) : __popc(mk0) + __popc(mk1 & lt);   // measured features before i
                const int rest = m + BASE + 3 * (i - pm) + pm;
                if (pr[h]) {
                    const int pos = 2 * pm;
                    s_perm[pos] = BASE + 3 * i; s_perm[pos + 1] = BASE + 3 * i + 1;
                    s_y[pos] = zx[h] - fx[h];
                    s_y[pos + 1] = zy[h] - fy[h];
#pragma unroll
                    for (int k = 0; k < 4; ++k) s_R[4 * pm + k] = rv[h][k];
                    p.klt_last[((size_t)f * nmax + i) * 2] = zx[h];
                    p.klt_last[((size_t)f * nmax + i) * 2 + 1] = zy[h];
                    if (rv[h][1] != rv[h][2]) elig = false;            // asymmetric R: the kernels without symmetry assumption
                    s_perm[rest] = BASE + 3 * i + 2;
                } else if (i < n) {
                    p.dflags[(size_t)f * nmax + i] = 1;
                    s_perm[rest] = BASE + 3 * i; s_perm[rest + 1] = BASE + 3 * i + 1; s_perm[rest + 2] = BASE + 3 * i + 2;
                }
            }
            if (lane < BASE) s_perm[m + lane] = lane;
            elig = __all_sync(0xffffffffu, elig);
            for (int a = N + lane; a < NTR * 8; a += 32) s_perm[a] = 0;
            for (int a = m + lane; a < NBM * 8; a += 32) s_y[a] = 0.0;
            if (lane == 0) { s_m = m; s_elig = elig ? 1 : 0; s_abort = 0; p.m[f] = m; }
        } else {
            // the mean, in its own order
            for (int a = tid - 32; a < N; a += (FW - 1) * 32) s_x[a] = a < BASE ? mu_g[a] : feat_g[a - BASE];
        }
        __syncthreads();
        if (!s_elig) {
            if (tid == 0) p.route[f] = -1;                // pending: ekf_chol_tiled decides
            continue;
        }
        const int m = s_m, nb = (m + 7) >> 3;
        FCLK(0);

        // The two roles run separate loops with the same barrier sequence (bar.sync 0 counts arrivals, not code addresses), so
        // that the accumulator tiles of the tile warps are not live — and not spilled — in the factorisation code of warp 15.
        if (warp != 15) {
            // ---- Sigma (lower triangle, permuted) into the accumulator registers ----
            double c0[18], c1[18];
#pragma unroll
            for (int s = 0; s < 18; ++s) { c0[s] = 0.0; c1[s] = 0.0; }
            if (typeF) {
#pragma unroll
                for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                    for (int cj = 0; cj < 4; ++cj) {
                        const double2 v = load_pair(Pi, ld, N, m, s_perm, s_y, 8 * (R0 + ri) + r, 8 * (C0 + cj) + 2 * q);
                        c0[ri * 4 + cj] = v.x; c1[ri * 4 + cj] = v.y;
                    }
            } else {
#pragma unroll
                for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                    for (int cj = 0; cj < 4; ++cj) {
                        if (cj <= ri) {
                            const double2 v = load_pair(Pi, ld, N, m, s_perm, s_y, 8 * (R0 + ri) + r, 8 * (C0 + cj) + 2 * q);
                            c0[tri(ri, cj)] = v.x; c1[tri(ri, cj)] = v.y;
                        }
                    }
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int cj = 0; cj < 4; ++cj) {
                        const double2 v = load_pair(Pi, ld, N, m, s_perm, s_y, 8 * (20 + e) + r, 8 * (C0 + cj) + 2 * q);
                        c0[10 + e * 4 + cj] = v.x; c1[10 + e * 4 + cj] = v.y;
                    }
            }
            {   // the next filter of this CTA: its Sigma rows towards L2 while this one is being worked on
                const int fn = f + gridDim.x;
                if (fn < p.F && tid < NTR * 8) {
                    const int Nn = BASE + 3 * p.nfeat[fn];
                    if (tid < Nn) {
                        const double* row = Pin + (size_t)fn * ld * ld + (size_t)tid * ld;
                        const unsigned bytes = (unsigned)((tid < BASE ? Nn : tid + 1) * 8 + 15) & ~15u;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(bytes) : "memory");
                    }
                }
            }
            FCLK(1);
            for (int j = 0; j < nb; ++j) {
                FCLK(3);
                cta_sync();                                    // inv(L_j) is there; the rank-8 update of step j-1 is complete
                FCLK(4);
                if (s_abort) break;
                const int mrem = m - 8 * j;                    // (>= 8 except in the last block)
                const int sj = j >> 2, cq = j & 3;
                // ---- the panel Z_j = Sigma(:,j) inv(L_j)' from the register tiles of column j and (transposed) row j ----
                {
                    const double lb0 = Li[tsw(r, q)], lb1 = Li[tsw(r, 4 + q)];       // B[k][col] = inv(L)[col][k]
                    if (typeF) {
                        if (C0 == 4 * sj) {
#pragma unroll
                            for (int cj = 0; cj < 4; ++cj)
                                if (cj == cq) {
#pragma unroll
                                    for (int ri = 0; ri < 4; ++ri) z_tile(Zs + (R0 + ri) * 64, c0[ri * 4 + cj], c1[ri * 4 + cj], false, lane, mrem, lb0, lb1);
                                }
                        }
                        if (R0 == 4 * sj) {
#pragma unroll
                            for (int ri = 0; ri < 4; ++ri)
                                if (ri == cq) {
#pragma unroll
                                    for (int cj = 0; cj < 4; ++cj) z_tile(Zs + (C0 + cj) * 64, c0[ri * 4 + cj], c1[ri * 4 + cj], true, lane, mrem, lb0, lb1);
                                }
                        }
                    } else if (R0 == 4 * sj) {
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj)
                            if (cj == cq) {
#pragma unroll
                                for (int ri = 0; ri < 4; ++ri)
                                    if (ri >= cj) z_tile(Zs + (R0 + ri) * 64, c0[tri(ri, cj)], c1[tri(ri, cj)], false, lane, mrem, lb0, lb1);
#pragma unroll
                                for (int e = 0; e < 2; ++e) z_tile(Zs + (20 + e) * 64, c0[10 + e * 4 + cj], c1[10 + e * 4 + cj], false, lane, mrem, lb0, lb1);
#pragma unroll
                                for (int c2 = 0; c2 < 4; ++c2)
                                    if (c2 < cj) z_tile(Zs + (C0 + c2) * 64, c0[tri(cj, c2)], c1[tri(cj, c2)], true, lane, mrem, lb0, lb1);
                            }
                    }
                }
                FCLK(2);
                cta_sync();                                    // Z_j is there
                FCLK(4);
                // ---- Sigma -= Z_j Z_j' on every tile (row N: the innovation) ----
                if (typeF) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        double za[4], zb[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            za[i] = -Zs[(R0 + i) * 64 + tsw(r, q + 4 * kk)];
                            zb[i] = Zs[(C0 + i) * 64 + tsw(r, q + 4 * kk)];
                        }
#pragma unroll
                        for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                            for (int cj = 0; cj < 4; ++cj) dmma884(c0[ri * 4 + cj], c1[ri * 4 + cj], za[ri], zb[cj]);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        double zr[4], ze[2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) zr[i] = Zs[(R0 + i) * 64 + tsw(r, q + 4 * kk)];
#pragma unroll
                        for (int e = 0; e < 2; ++e) ze[e] = -Zs[(20 + e) * 64 + tsw(r, q + 4 * kk)];
#pragma unroll
                        for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                            for (int cj = 0; cj < 4; ++cj)
                                if (cj <= ri) dmma884(c0[tri(ri, cj)], c1[tri(ri, cj)], -zr[ri], zr[cj]);
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int cj = 0; cj < 4; ++cj) dmma884(c0[10 + e * 4 + cj], c1[10 + e * 4 + cj], ze[e], zr[cj]);
                    }
                }
                if (j + 2 + warp < nb) diag_update(Dg + (j + 2 + warp) * 64, Zs + (j + 2 + warp) * 64, lane);   // the private diagonal tiles beyond j+1, one per warp
            }
            FCLK(3);
            cta_sync();
            FCLK(4);
            // ---- Sigma' back to its own order: into the staging buffer ----
            if (!s_abort) {
                if (typeF) {
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            const int pr = 8 * (R0 + ri) + r, pc = 8 * (C0 + cj) + 2 * q;
                            stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc, c0[ri * 4 + cj]);
                            stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc + 1, c1[ri * 4 + cj]);
                        }
                } else {
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            if (cj <= ri) {
                                const int pr = 8 * (R0 + ri) + r, pc = 8 * (C0 + cj) + 2 * q;
                                stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc, c0[tri(ri, cj)]);
                                stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc + 1, c1[tri(ri, cj)]);
                            }
                        }
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            const int pr = 8 * (20 + e) + r, pc = 8 * (C0 + cj) + 2 * q;
                            stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc, c0[10 + e * 4 + cj]);
                            stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, pr, pc + 1, c1[10 + e * 4 + cj]);
                        }
                }
            }
        } else {
            // ---- warp 15: tiles (20,20) (21,20) (21,21), and the factorisation of the measurement diagonal tiles one step ahead ----
            for (int I = 0; I < nb; ++I) {
                const double2 v = load_pair(Pi, ld, N, m, s_perm, s_y, 8 * I + r, 8 * I + 2 * q);
                *reinterpret_cast<double2*>(&Dg[I * 64 + tsw(r, 2 * q)]) = v;
            }
            double2 t0 = load_pair(Pi, ld, N, m, s_perm, s_y, 160 + r, 160 + 2 * q);
            double2 t1 = load_pair(Pi, ld, N, m, s_perm, s_y, 168 + r, 160 + 2 * q);
            double2 t2 = load_pair(Pi, ld, N, m, s_perm, s_y, 168 + r, 168 + 2 * q);
            double dmin = 1.79e308, dmax = 0.0;
            FCLK(1);
            if (nb > 0) {
                __syncwarp();
                if (!factor_block(Dg, Li, s_R, 0, m, p.illcond, dmin, dmax, lane) && lane == 0) s_abort = 1;
            }
            for (int j = 0; j < nb; ++j) {
                FCLK(3);
                cta_sync();
                FCLK(4);
                if (s_abort) break;
                FCLK(2);
                cta_sync();
                FCLK(4);
                if (j + 1 < nb) {      // look-ahead: the next diagonal tile and its factorisation, while the tile warps update Sigma
                    diag_update(Dg + (j + 1) * 64, Zs + (j + 1) * 64, lane);
                    __syncwarp();
                    if (!factor_block(Dg + (j + 1) * 64, Li, s_R, j + 1, m, p.illcond, dmin, dmax, lane) && lane == 0) s_abort = 1;
                }
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const double z20 = Zs[20 * 64 + tsw(r, q + 4 * kk)], z21 = Zs[21 * 64 + tsw(r, q + 4 * kk)];
                    dmma884(t0.x, t0.y, -z20, z20);
                    dmma884(t1.x, t1.y, -z21, z20);
                    dmma884(t2.x, t2.y, -z21, z21);
                }
            }
            FCLK(3);
            cta_sync();
            FCLK(4);
            if (!s_abort) {
                stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 160 + r, 160 + 2 * q, t0.x); stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 160 + r, 161 + 2 * q, t0.y);
                stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 168 + r, 160 + 2 * q, t1.x); stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 168 + r, 161 + 2 * q, t1.y);
                stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 168 + r, 168 + 2 * q, t2.x); stage_elem(Ob, s_ooff, s_dl, s_y, N, m, s_perm, 168 + r, 169 + 2 * q, t2.y);
            }
        }
        if (s_abort) {
            if (tid == 0) p.route[f] = -1;                // ekf_chol_tiled routes it (Joseph form / signed factor)
            continue;
        }
        __syncthreads();
        FCLK(5);
        // mu += K y (:600), in shared memory
        if (tid < N && m > 0) s_x[s_perm[tid]] += s_dl[tid];
        // rows of Sigma': a warp per row, 16-byte stores; row a is valid up to min(orow_len(a), N)
        for (int a = warp; a < N; a += FW) {
            const int len = min(orow_len(a), N);
            const double* src = Ob + s_ooff[a];
            double* dst = Po + (size_t)a * ld;
            for (int c = 2 * lane; c + 1 < len; c += 64) *reinterpret_cast<double2*>(dst + c) = *reinterpret_cast<const double2*>(src + c);
            if ((len & 1) && lane == 0) dst[len - 1] = src[len - 1];
        }
        __syncthreads();
        if (warp == 0) {  // renormalise the quaternion (:605-609) and flag non-finite states
            const double qn = sqrt(s_x[3] * s_x[3] + s_x[4] * s_x[4] + s_x[5] * s_x[5] + s_x[6] * s_x[6]);
            double v = lane < BASE ? s_x[lane] : 0.0;
            if (lane >= 3 && lane <= 6) v /= qn;
            const bool fin = __all_sync(0xffffffffu, isfinite(v));
            if (lane < BASE) mu_g[lane] = v;
            if (lane == 0) {
                if (!fin) atomicOr(&p.status[f], 2);
                p.route[f] = ROUTE_DONE;
            }
        } else if (m > 0) {
            for (int a = BASE + tid - 32; a < N; a += (FW - 1) * 32) feat_g[a - BASE] = s_x[a];
        }
        FCLK(6);
    }
}

}  // namespace

namespace ekfvio {

bool update_fused_supported(const EkfPtrs& p) { return p.Nmax <= NTR * 8 && p.mmax <= NBM * 8; }

cudaError_t launch_update_fused(const EkfPtrs& p, const double* Pin, double* Pout, const double* z, const double* R, const uint8_t* pass,
                                cudaStream_t st) {
    static int sms_count_on[64] = {0};
    int& sms_count = sms_count_on[current_device_slot()];
    if (!sms_count) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const size_t sm = (size_t)OB_DOUBLES * sizeof(double);
    static bool configured_on[64] = {false};
    bool& configured = configured_on[current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_update_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int grid = p.F < sms_count ? p.F : sms_count;    // persistent: one CTA per SM
    ekf_update_fused<<<grid, FW * 32, sm, st>>>(p, Pin, Pout, z, R, pass);
    return cudaGetLastError();
}

}  // namespace ekfvio

#ifdef EKFVIO_PROFILE_CLOCKS
// slots 0..7: warp 0 (0 map, 1 load, 2 panel, 3 update, 4 barrier wait, 5 store, 6 tail); 8..15: the same marks on warp 15
extern "C" void ekfvio_debug_fused_clocks(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_fclk, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_fclk, z, sizeof(z)); }
}
#endif
