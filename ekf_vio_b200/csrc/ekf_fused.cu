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
// Warp roles (tile units; 22 tile rows cover N <= 176).  The 5x5 grid of 4x4-tile super blocks covers tile rows 0..19:
//   12 tile warps (warp & 3 != 3): ten hold one full super block (R0, C0) below the diagonal, two the lower triangle of diagonal
//                super block 3 resp. 4 (10 tiles) plus the 2x4 tiles of rows 20, 21 under it — 16 or 18 tiles in registers
//   57 tiles in shared memory (Ts): diagonal super blocks 0..2, rows 20, 21 under them, and (20,20) (21,20) (21,21); the tile
//                warps update five of them each per step
//   warp 15      the serial chain: it factors the 8x8 diagonal tile of step j+1 (private copies Dg, updated by the tile warps)
//                WHILE the tile warps run the rank-8 update of step j (look-ahead)
//   warps 3, 7, 11 share warp 15's scheduler and therefore stay off the FP64 pipe during the update phase — DMMA and DFMA issue
//                on one pipe, and a chain of dependent DFMAs makes no headway against the queue of three DMMA-issuing warps
//                (measured: the factorisation took 4.2 k instead of 1.9 k clocks per step).  They do the panel tiles that come
//                from Ts, the L2 prefetch of the next filter, and the loading / staging of Ts.
// A filter whose S turns out not positive definite or beyond the pivot ratio of ILLCOND_RATIO is abandoned here (nothing has
// been written) and served by the kernels of ekf_tiled.cu, which decide its route as before.
#include <cstdlib>

#include "ekf_common.cuh"
#include "ekf_kernels.h"
#include "ekf_tiles.cuh"

using namespace ekfvio;

namespace {

constexpr int NTR = 22;          // tile rows
constexpr int NBM = 13;          // measurement blocks (m <= 104)
constexpr int FW = 16;           // warps

#ifdef EKFVIO_PROFILE_CLOCKS
__device__ unsigned long long g_fclk[24];
#define FSUB(i) do { if (lane == 0) { long long t_ = clock64(); atomicAdd(&g_fclk[16 + (i)], (unsigned long long)(t_ - t_sub)); t_sub = t_; } } while (0)
#define FCLK(i) do { if (lane == 0 && (warp == 0 || warp == 15)) { long long t_ = clock64(); atomicAdd(&g_fclk[(i) + (warp == 15 ? 8 : 0)], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define FCLK(i) do {} while (0)
#define FSUB(i) do {} while (0)
#endif

__device__ __forceinline__ int tri(int ri, int cj) { return ri * (ri + 1) / 2 + cj; }

// the tiles kept in shared memory: diagonal super blocks 0..2 (lower triangles), rows 20, 21 x columns 0..11, the last three
constexpr int NTS = 57;
struct TsTable { unsigned char ij[NTS][2]; };
constexpr TsTable make_ts_table() {
    TsTable t{};
    int k = 0;
    for (int d = 0; d < 3; ++d)
        for (int ri = 0; ri < 4; ++ri)
            for (int cj = 0; cj <= ri; ++cj) { t.ij[k][0] = (unsigned char)(4 * d + ri); t.ij[k][1] = (unsigned char)(4 * d + cj); ++k; }
    for (int e = 0; e < 2; ++e)
        for (int c = 0; c < 12; ++c) { t.ij[k][0] = (unsigned char)(20 + e); t.ij[k][1] = (unsigned char)c; ++k; }
    t.ij[k][0] = 20; t.ij[k][1] = 20; ++k;
    t.ij[k][0] = 21; t.ij[k][1] = 20; ++k;
    t.ij[k][0] = 21; t.ij[k][1] = 21; ++k;
    return t;
}
__constant__ TsTable c_ts = make_ts_table();
__device__ __forceinline__ int ts_index(int I, int J) {      // (I, J) must be one of the 57
    if (I < 12) return 10 * (I >> 2) + tri(I & 3, J & 3);
    if (J < 12) return 30 + 12 * (I - 20) + J;
    return 54 + (I - 20) + (J - 20);
}

// Staging buffer of the result (dynamic shared memory): the 253 lower-triangular tiles of the permuted Sigma' as the warps hold
// them (row-major 8x8, conflict-free fragment stores); the copy-out pass un-permutes while it writes rows of the stored matrix.
constexpr int NTILES = NTR * (NTR + 1) / 2;
constexpr int OB_DOUBLES = NTILES * 64;
__device__ __forceinline__ double staged(const double* Tb, int pa, int pc) {      // element of the permuted result, either triangle
    const int hi = max(pa, pc), lo = min(pa, pc);
    const int I = hi >> 3, J = lo >> 3;
    return Tb[(I * (I + 1) / 2 + J) * 64 + (hi & 7) * 8 + (lo & 7)];
}
__device__ __forceinline__ void stage_tile(double* Tb, int I, int J, int lane, double c0, double c1) {
    *reinterpret_cast<double2*>(&Tb[(I * (I + 1) / 2 + J) * 64 + (lane >> 2) * 8 + 2 * (lane & 3)]) = make_double2(c0, c1);
}

// element pair (pr, pc), (pr, pc+1) of the permuted matrix, read through the lower triangle of the stored one.  Permuted row N
// (the first spare row) carries the innovation: y over the measured columns, zero elsewhere — the rank-8 updates then turn it
// into y - dmu over the measured and -dmu over the other columns, i.e. K y comes out of the elimination itself.
// (Ib: the lower triangle of the stored matrix, row r at in_off(r), rows padded to an even length)
__device__ __forceinline__ int in_off(int r) { const int h = (r + 1) >> 1; return (r & 1) ? 2 * h * h : 2 * h * (h + 1); }
// the same with the row / column look-ups hoisted: ra = perm[pr] (-1: beyond the matrix, -2: the innovation row), oa = in_off(ra);
// b, ob likewise for the two columns
struct PermIdx { int a, off; };
__device__ __forceinline__ PermIdx perm_idx(const int* perm, int N, int pr, bool row) {
    PermIdx x; x.a = -1; x.off = 0;
    if (pr < N) { x.a = perm[pr]; x.off = in_off(x.a); }
    else if (row && pr == N) x.a = -2;
    return x;
}
__device__ __forceinline__ double load_elem(const double* Ib, int m, const double* sy, const PermIdx& ra, const PermIdx& cb, int pc) {
    if (ra.a >= 0) return cb.a >= 0 ? Ib[ra.a >= cb.a ? ra.off + cb.a : cb.off + ra.a] : 0.0;
    return (ra.a == -2 && pc < m) ? sy[pc] : 0.0;
}
__device__ __forceinline__ double2 load_pair(const double* Ib, int N, int m, const int* perm, const double* sy, int pr, int pc) {
    double2 v = make_double2(0.0, 0.0);
    if (pr < N) {
        const int a = perm[pr];
        if (pc < N) { const int b = perm[pc]; v.x = Ib[in_off(max(a, b)) + min(a, b)]; }
        if (pc + 1 < N) { const int b = perm[pc + 1]; v.y = Ib[in_off(max(a, b)) + min(a, b)]; }
    } else if (pr == N) {
        if (pc < m) v.x = sy[pc];
        if (pc + 1 < m) v.y = sy[pc + 1];
    }
    return v;
}

// 1 / d to full FP64 accuracy for the normal, positive pivots the elimination accepts (a pivot that is not is rejected by the
// caller anyway): hardware seed (2^-20) and one cubic Newton step, e = 1 - d r, r (1 + e + e^2) — three dependent FMAs, ~1 ulp
__device__ __forceinline__ double rcp_fast(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);                      // 1/d = r (1 + e + e^2 + e^3 ...): truncation 2^-60
}

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

// bulk (TMA) row copies global -> shared memory behind an mbarrier
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
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
    double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;     // two independent DMMAs: a dependent pair costs a second pipeline latency (~300 clk)
    dmma884(t0, t1, a0, lb0);
    dmma884(u0, u1, a1, lb1);
    *reinterpret_cast<double2*>(&Zt[tsw(r, 2 * q)]) = make_double2(t0 + u0, t1 + u1);
}

// the same from a tile in shared memory (tsw layout): no shuffles, the A fragment is read in either orientation
__device__ __forceinline__ void z_tile_smem(double* Zt, const double* T, bool transposed, int lane, int mrem, double lb0, double lb1) {
    const int r = lane >> 2, q = lane & 3;
    double a0 = transposed ? T[tsw(q, r)] : T[tsw(r, q)];
    double a1 = transposed ? T[tsw(4 + q, r)] : T[tsw(r, 4 + q)];
    if (q >= mrem) a0 = 0.0;
    if (q + 4 >= mrem) a1 = 0.0;
    double t0 = 0.0, t1 = 0.0, u0 = 0.0, u1 = 0.0;
    dmma884(t0, t1, a0, lb0);
    dmma884(u0, u1, a1, lb1);
    *reinterpret_cast<double2*>(&Zt[tsw(r, 2 * q)]) = make_double2(t0 + u0, t1 + u1);
}

// S_j = Dt + R_j (lower triangle), factored S_j = L L' by warp 15; inv(L) to Li (tsw layout).  Rows of the block beyond m are
// replaced by identity rows.  Returns false (on every lane) for a pivot that is not positive or breaks the pivot ratio.
// Elimination in LDL^T form (the reference's unpivoted SimplicialLDLT order, :577-580): one reciprocal per column on the
// critical path; the square roots only scale the finished columns.
// Zt != nullptr: the rank-8 update of this tile by the previous step's panel, Dt - Zt Zt', is applied on the way in — by DFMA, row
// per lane: a DMMA here would put its pipeline latency (and the queue of the tile warps' DMMAs) on the factorisation chain.
__device__ __forceinline__ bool factor_block(const double* Dt, const double* Zt, double* Li, const double* sR, int j, int m,
                                             double illcond, double& dmin, double& dmax, int lane) {
    const int rr = lane & 7;
    const bool meas = 8 * j + rr < m;
#ifdef EKFVIO_PROFILE_CLOCKS
    long long t_sub = clock64();
#endif
    // row rr of the tile, straight-line: all eight columns are loaded and updated (only c <= rr is used further down), so that
    // the independent chains interleave.  The rank-8 update is split over the four lane groups (k = 2g, 2g+1 on lanes 8g..8g+7)
    // and summed by two shuffle stages.
    const int g = lane >> 3;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c += 2) { const double2 v = *reinterpret_cast<const double2*>(&Dt[tsw(rr, c)]); a[c] = v.x; a[c + 1] = v.y; }
    if (Zt != nullptr) {
        const double2 zr = *reinterpret_cast<const double2*>(&Zt[tsw(rr, 2 * g)]);
        double part[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 zc = *reinterpret_cast<const double2*>(&Zt[tsw(c, 2 * g)]);
            part[c] = fma(zr.x, zc.x, zr.y * zc.y);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            part[c] += __shfl_xor_sync(0xffffffffu, part[c], 8);
            part[c] += __shfl_xor_sync(0xffffffffu, part[c], 16);
            a[c] -= part[c];
        }
    }
    {   // + R_j on the 2x2 diagonal blocks: upper(S)(c, rr), :559-561, :578 (R symmetric here); rows beyond m: identity
        const double* rp = sR + 4 * (4 * j + (rr >> 1)) + (rr & 1);
        const double r0 = rp[0], r1 = rp[2];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if ((c >> 1) == (rr >> 1)) a[c] += (c & 1) ? r1 : r0;
            if (!meas) a[c] = (c == rr) ? 1.0 : 0.0;
        }
    }
    FSUB(0);
    double own = 1.0;                  // this row's pivot
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);            // pivot
        double l[8];
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) l[c2] = __shfl_sync(0xffffffffu, a[c], c2);
        if (c == rr) own = d;
        const double t = a[c] * rcp_fast(d);
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2)
            if (rr >= c2) a[c2] -= t * l[c2];
    }
    FSUB(1);
    // pivot checks and 1 / sqrt(d) once per row, off the elimination chain
    bool ok = own > 0.0;
    const double myrs = rsqrt(own);
    double pmax = meas ? own : 0.0, pmin = meas ? own : 1.79e308;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o)); pmin = fmin(pmin, __shfl_xor_sync(0xffffffffu, pmin, o)); }
    dmax = fmax(dmax, pmax); dmin = fmin(dmin, pmin);
    ok = __all_sync(0xffffffffu, ok) && (dmax <= illcond * dmin);
    double rs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) rs[c] = __shfl_sync(0xffffffffu, myrs, c);
    FSUB(2);
    // L = L_ldl |D|^1/2; inv(L)(i, rr) = rs_i (delta_i,rr - sum_k L(i,k) inv(L)(k, rr)): with the rows pre-scaled by their own
    // rs_i the forward substitution is one dependent FMA per row
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (c < rr) ? a[c] * rs[c] * myrs : 0.0;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double sacc = (i == rr) ? rs[i] : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) sacc -= __shfl_sync(0xffffffffu, a[k], i) * x[k];
        x[i] = sacc;
    }
    FSUB(3);
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) Li[tsw(i, rr)] = (i >= rr) ? x[i] : 0.0;
    }
    FSUB(4);
    return ok;
}

__global__ void __launch_bounds__(FW * 32, 1) ekf_update_fused(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout,
                                                               const double* __restrict__ z, const double* __restrict__ Rin,
                                                               const uint8_t* __restrict__ pass, int variant) {
    extern __shared__ __align__(16) double Ob[];      // staging buffer of the result (OB_DOUBLES)
    __shared__ __align__(16) double Zs[NTR * 64];     // the panel Z_j, one swizzled 8x8 tile per tile row
    __shared__ __align__(16) double Dg[NBM * 64];     // private copies of the measurement diagonal tiles (look-ahead factorisation)
    __shared__ __align__(16) double Li[64];           // inv(L_j)
    double* Ts = Ob + OB_DOUBLES;                     // the tiles no warp holds in registers (tsw layout; NTS * 64)
    __shared__ double s_y[NBM * 8], s_x[NTR * 8], s_R[NBM * 16];
    __shared__ int s_perm[NTR * 8];
    __shared__ int2 s_hl[NTR * 8];                   // copy-out: per stored index, the staged offset of its permuted position as the larger (.x) / smaller (.y) index of a pair
    __shared__ unsigned long long s_bar;              // mbarrier of the bulk row loads
    __shared__ int s_m, s_elig, s_abort;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int ld = p.ldP, nmax = p.nmax;

    // role of this warp
    const bool is_factor = warp == 15, is_helper = (warp & 3) == 3 && !is_factor;
    const int slot = warp - (warp >> 2);       // tile warps: 0..11
    const int helper = warp >> 2;              // helper warps: 0..2
    int R0 = 0, C0 = 0;          // type F: first tile row / column of the 4x4 block; type D: R0 = C0 = 4 d
    if (slot < 10) {
        int sr = 1, w = slot;
        while (w >= sr) { w -= sr; ++sr; }
        R0 = 4 * sr; C0 = 4 * w;
    } else {
        R0 = C0 = 4 * (slot - 7);              // diagonal super blocks 3 and 4
    }
    const bool typeF = slot < 10;
    // Every CTA has the same work per filter, so the CTAs would run in lockstep and hit HBM together: 148 x 125 KB of loads, then
    // nothing, then 148 x 125 KB of stores.  A start-up skew of up to ~20 us spreads the memory phases of the CTAs over the
    // compute phases of the others for the rest of the launch (variant bit 1 switches it off).
    if (!(variant & 2) && gridDim.x > 8 && p.F > 2 * (int)gridDim.x) {
        const long long until = clock64() + (long long)(blockIdx.x & 7) * 5000;
        while (clock64() < until) __nanosleep(200);
    }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned bar_phase = 0;
    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const int n = p.nfeat[f], N = BASE + 3 * n;
        const double* Pi = Pin + (size_t)f * ld * ld;
        double* Po = Pout + (size_t)f * ld * ld;
        double* mu_g = p.mu + (size_t)f * BASE;
        double* feat_g = p.feat + (size_t)f * nmax * 3;

        fence_proxy_async();                               // (the copy-out's reads of Ob before the bulk loads' writes)
        __syncthreads();                                   // shared memory of the previous filter is free
#ifdef EKFVIO_PROFILE_CLOCKS
        long long t_prev = clock64();
#endif
        // the lower triangle of Sigma, one bulk copy (TMA) per row, into the buffer that stages the result later
        if (tid == 0) mbar_expect_tx(&s_bar, 8u * (unsigned)in_off(N));
        if (tid < N) bulk_load(Ob + in_off(tid), Pi + (size_t)tid * ld, 8u * (unsigned)((tid + 2) & ~1), &s_bar);
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
        mbar_wait(&s_bar, bar_phase);
        bar_phase ^= 1;
        __syncthreads();
        if (!s_elig) {
            if (tid == 0) { p.route[f] = -1; p.fb[1 + atomicAdd(p.fb, 1)] = f; }      // pending: ekf_chol_tiled decides
            continue;
        }
        const int m = s_m, nb = (m + 7) >> 3;
        FCLK(0);

        // The two roles run separate loops with the same barrier sequence (bar.sync 0 counts arrivals, not code addresses), so
        // that the accumulator tiles of the tile warps are not live — and not spilled — in the factorisation code of warp 15.
        if (!is_factor && !is_helper) {
            // ---- Sigma (lower triangle, permuted) into the accumulator registers ----
            double c0[18], c1[18];
#pragma unroll
            for (int s = 0; s < 18; ++s) { c0[s] = 0.0; c1[s] = 0.0; }
            {
                PermIdx cb[4][2];
#pragma unroll
                for (int cj = 0; cj < 4; ++cj) { cb[cj][0] = perm_idx(s_perm, N, 8 * (C0 + cj) + 2 * q, false); cb[cj][1] = perm_idx(s_perm, N, 8 * (C0 + cj) + 2 * q + 1, false); }
#pragma unroll
                for (int ri = 0; ri < 4; ++ri) {
                    const PermIdx ra = perm_idx(s_perm, N, 8 * (R0 + ri) + r, true);
#pragma unroll
                    for (int cj = 0; cj < 4; ++cj) {
                        if (typeF) {
                            c0[ri * 4 + cj] = load_elem(Ob, m, s_y, ra, cb[cj][0], 8 * (C0 + cj) + 2 * q);
                            c1[ri * 4 + cj] = load_elem(Ob, m, s_y, ra, cb[cj][1], 8 * (C0 + cj) + 2 * q + 1);
                        } else if (cj <= ri) {
                            c0[tri(ri, cj)] = load_elem(Ob, m, s_y, ra, cb[cj][0], 8 * (C0 + cj) + 2 * q);
                            c1[tri(ri, cj)] = load_elem(Ob, m, s_y, ra, cb[cj][1], 8 * (C0 + cj) + 2 * q + 1);
                        }
                    }
                }
                if (!typeF) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const PermIdx ra = perm_idx(s_perm, N, 8 * (20 + e) + r, true);
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            c0[10 + e * 4 + cj] = load_elem(Ob, m, s_y, ra, cb[cj][0], 8 * (C0 + cj) + 2 * q);
                            c1[10 + e * 4 + cj] = load_elem(Ob, m, s_y, ra, cb[cj][1], 8 * (C0 + cj) + 2 * q + 1);
                        }
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
                // up to five of the shared-memory tiles belong to this warp: three before and two after its register tiles, so that
                // their loads and (dependent) DMMA pairs overlap with independent work
                auto smem_tiles = [&](int k0, int k1) {
                    double2 tc[3]; double ta[3][2], tb[3][2];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int t = slot + 12 * (k0 + k);
                        if (k0 + k < k1 && t < NTS) {
                            const int I = c_ts.ij[t][0], J = c_ts.ij[t][1];
                            tc[k] = *reinterpret_cast<const double2*>(&Ts[t * 64 + tsw(r, 2 * q)]);
                            ta[k][0] = -Zs[I * 64 + tsw(r, q)]; ta[k][1] = -Zs[I * 64 + tsw(r, 4 + q)];
                            tb[k][0] = Zs[J * 64 + tsw(r, q)]; tb[k][1] = Zs[J * 64 + tsw(r, 4 + q)];
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (k0 + k < k1 && slot + 12 * (k0 + k) < NTS) dmma884(tc[k].x, tc[k].y, ta[k][kk], tb[k][kk]);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (k0 + k < k1 && slot + 12 * (k0 + k) < NTS) *reinterpret_cast<double2*>(&Ts[(slot + 12 * (k0 + k)) * 64 + tsw(r, 2 * q)]) = tc[k];
                };
                smem_tiles(0, 3);
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
                smem_tiles(3, 5);
                if (j + 2 + slot < nb) diag_update(Dg + (j + 2 + slot) * 64, Zs + (j + 2 + slot) * 64, lane);   // the private diagonal tiles beyond j+1, one per warp
            }
            FCLK(3);
            cta_sync();
            FCLK(4);
            // ---- the tiles into the staging buffer ----
            if (!s_abort) {
                if (typeF) {
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) stage_tile(Ob, R0 + ri, C0 + cj, lane, c0[ri * 4 + cj], c1[ri * 4 + cj]);
                } else {
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj)
                            if (cj <= ri) stage_tile(Ob, R0 + ri, C0 + cj, lane, c0[tri(ri, cj)], c1[tri(ri, cj)]);
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) stage_tile(Ob, 20 + e, C0 + cj, lane, c0[10 + e * 4 + cj], c1[10 + e * 4 + cj]);
                }
            }
        } else if (is_helper) {
            // ---- warps 3, 7, 11: the tiles in shared memory (load, panel tiles, staging) and the L2 prefetch; no FP64 work while
            //      warp 15 factors ----
            for (int t = helper; t < NTS; t += 3) {
                const int I = c_ts.ij[t][0], J = c_ts.ij[t][1];
                const double2 v = load_pair(Ob, N, m, s_perm, s_y, 8 * I + r, 8 * J + 2 * q);
                *reinterpret_cast<double2*>(&Ts[t * 64 + tsw(r, 2 * q)]) = v;
            }
            FCLK(1);
            for (int j = 0; j < nb; ++j) {
                cta_sync();
                if (s_abort) break;
                const int mrem = m - 8 * j;
                const int sj = j >> 2, cq = j & 3;
                if (sj < 3) {
                    // panel tiles whose source lives in Ts: (j..4sj+3, j) and (20, j), (21, j) direct, (j, 4sj..j-1) transposed — six tiles
                    const double lb0 = Li[tsw(r, q)], lb1 = Li[tsw(r, 4 + q)];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int t = helper + 3 * k;
                        if (t < 4 - cq) z_tile_smem(Zs + (j + t) * 64, Ts + ts_index(j + t, j) * 64, false, lane, mrem, lb0, lb1);
                        else if (t < 6 - cq) z_tile_smem(Zs + (20 + t - (4 - cq)) * 64, Ts + ts_index(20 + t - (4 - cq), j) * 64, false, lane, mrem, lb0, lb1);
                        else z_tile_smem(Zs + (4 * sj + t - (6 - cq)) * 64, Ts + ts_index(j, 4 * sj + t - (6 - cq)) * 64, true, lane, mrem, lb0, lb1);
                    }
                }
                cta_sync();
                // (no FP64 work here: measured with 18 / 36 / 57 of the shared-memory tiles updated by these warps, the factorisation on
                // warp 15 slowed down from 2.5 k to 3.5 - 4.6 k clocks per step and the launch by 20 - 45 %)
                if (j == 0) {   // the next filter of this CTA: the lower triangle of its Sigma towards L2 while this one is being worked on
                    const int fn = f + gridDim.x;
                    if (fn < p.F) {
                        const int Nn = BASE + 3 * p.nfeat[fn];
                        const char* base = reinterpret_cast<const char*>(Pin + (size_t)fn * ld * ld);
                        for (int a = helper; a < Nn; a += 3) {
                            if (lane * 128 < (a + 1) * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)a * ld * 8 + lane * 128));
                        }
                    }
                }
            }
            cta_sync();
            if (!s_abort) {
                for (int t = helper; t < NTS; t += 3) {
                    const double2 v = *reinterpret_cast<const double2*>(&Ts[t * 64 + tsw(r, 2 * q)]);
                    stage_tile(Ob, c_ts.ij[t][0], c_ts.ij[t][1], lane, v.x, v.y);
                }
            }
        } else {
            // ---- warp 15: the factorisation of the measurement diagonal tiles, one step ahead ----
            for (int I = 0; I < nb; ++I) {
                const double2 v = load_pair(Ob, N, m, s_perm, s_y, 8 * I + r, 8 * I + 2 * q);
                *reinterpret_cast<double2*>(&Dg[I * 64 + tsw(r, 2 * q)]) = v;
            }
            double dmin = 1.79e308, dmax = 0.0;
            FCLK(1);
            if (nb > 0) {
                __syncwarp();
                if (!factor_block(Dg, nullptr, Li, s_R, 0, m, p.illcond, dmin, dmax, lane) && lane == 0) s_abort = 1;
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
                    if (!factor_block(Dg + (j + 1) * 64, Zs + (j + 1) * 64, Li, s_R, j + 1, m, p.illcond, dmin, dmax, lane) && lane == 0) s_abort = 1;
                }
            }
            FCLK(3);
            cta_sync();
            FCLK(4);
        }
        if (s_abort) {
            if (tid == 0) { p.route[f] = -1; p.fb[1 + atomicAdd(p.fb, 1)] = f; }      // ekf_chol_tiled routes it (Joseph form / signed factor)
            continue;
        }
        // staged(pa, pc) = x(max) + y(min) with x(p) = tri(p / 8) 64 + (p % 8) 8 and y(p) = (p / 8) 64 + p % 8; y is monotonic in p, so the
        // comparison of two permuted positions is the comparison of their y: the copy-out needs one look-up per stored index
        if (tid < N) s_hl[s_perm[tid]] = make_int2(tri(tid >> 3, 0) * 64 + (tid & 7) * 8, (tid >> 3) * 64 + (tid & 7));
        __syncthreads();
        FCLK(5);
        // mu += K y (:600): permuted row N of the result is y - dmu over the measured columns and -dmu over the others
        if (tid < N && m > 0) {
            const double v = staged(Ob, N, tid);
            s_x[s_perm[tid]] += ((tid < m) ? s_y[tid] : 0.0) - v;
        }
        // rows of Sigma' in their own order, a warp per row: the lower triangle, rows 0..21 complete and the 3x3 diagonal block
        // of each feature complete (what the readers of a lower-mode Sigma rely on)
        for (int a0 = warp; a0 < N; a0 += 4 * FW) {          // four rows in flight per warp: independent shared-memory look-ups
            int rx[4], ry[4], len[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int a = a0 + FW * u;
                const int2 h = s_hl[a < N ? a : 0];
                rx[u] = h.x; ry[u] = h.y;
                len[u] = a < N ? (a < BASE ? N : BASE + 3 * ((a - BASE) / 3) + 3) : 0;
            }
            const int maxlen = max(max(len[0], len[1]), max(len[2], len[3]));
            double* po = Po + (size_t)a0 * ld;
            // (the bounds test stays around the look-up: with the eight look-ups of two column groups issued ahead of predicated
            // stores the launch was 2 % slower, 1.330 against 1.301 ms — DESIGN.md section 8)
            for (int c = lane; c < maxlen; c += 32) {
                const int2 h = s_hl[c];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c < len[u]) po[(size_t)(FW * u) * ld + c] = prune(Ob[ry[u] >= h.y ? rx[u] + h.y : h.x + ry[u]]);
            }
        }
#ifdef EKFVIO_PROFILE_CLOCKS
        if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_fclk[21], (unsigned long long)(t_ - t_prev)); }
#endif
        __syncthreads();
#ifdef EKFVIO_PROFILE_CLOCKS
        if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_fclk[22], (unsigned long long)(t_ - t_prev)); }
#endif
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
    const size_t sm = (size_t)(OB_DOUBLES + NTS * 64) * sizeof(double);
    static bool configured_on[64] = {false};
    bool& configured = configured_on[current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_update_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int grid = p.F < sms_count ? p.F : sms_count;    // persistent: one CTA per SM
    static int variant = -1;                              // EKFVIO_FUSED_VARIANT: scheduling experiments (bit 1: no start-up skew)
    if (variant < 0) { const char* e = getenv("EKFVIO_FUSED_VARIANT"); variant = e ? atoi(e) : 0; }
    ekf_update_fused<<<grid, FW * 32, sm, st>>>(p, Pin, Pout, z, R, pass, variant);
    return cudaGetLastError();
}

}  // namespace ekfvio

#ifdef EKFVIO_PROFILE_CLOCKS
// slots 0..7: warp 0 (0 map, 1 load, 2 panel, 3 update, 4 barrier wait, 5 store, 6 tail); 8..15: the same marks on warp 15
extern "C" void ekfvio_debug_fused_clocks(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_fclk, sizeof(unsigned long long) * 24);
    if (reset) { unsigned long long z[24] = {0}; cudaMemcpyToSymbol(g_fclk, z, sizeof(z)); }
}
#endif
