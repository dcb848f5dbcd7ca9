// Large feature-augmented states (n > 51: N = 22 + 3n up to ~1000, m = 2n up to ~700): the EKF
// measurement update as blocked dense linear algebra on FP64 tensor-core tiles (DMMA.8x8x4), with
// several CTAs per filter.  Same mathematics as the small-state path (ekf_tiled.cu):
//
//   S = Sigma(idx,idx) + R,   L L' = sym(upper(S))            (TightlyCoupledEKF.cpp:559-561, 577-578)
//   K = Sigma(:,idx) L^-T L^-1, sparseView, mu += K y          (:580, :600)
//   W = Sigma(:,idx) - K S,   Sigma' = Sigma - K Sigma(idx,:) - W K'   (:586-596, 625, Joseph form)
//
// organised LAPACK-style in 64-wide column blocks:
//   potrf_diag + trsm(PANEL) + gemm(SYRK) per block      -> L
//   trsm(FWD) + gemm(FWDUPD) per block, ascending        -> Z = C L^-T
//   trsm(BWD) + gemm(BWDUPD) per block, descending       -> K = Z L^-1
//   finalize (prune, mu += K y), gemm(W), gemm(JOSEPH)
// The 64x64 diagonal blocks are handled exactly like the small path handles a whole S: 8x8 swizzled
// tiles in shared memory, explicit inverses of the 8x8 diagonal tiles only, strips of 16 rows carried
// through the substitution in registers.  Everything else is one 64x64-tile DMMA GEMM kernel
// (D = C - A B') with cp.async-staged operands in three access modes.
#include "ekf_common.cuh"
#include "ekf_kernels.h"
#include "ekf_tiles.cuh"

using namespace ekfvio;

namespace {

constexpr int BLK = 64;             // column block
constexpr int NB8 = 8;              // 8x8 tiles per block side
constexpr int NT8 = NB8 * (NB8 + 1) / 2;

// Symmetric filters take the reduced update Sigma' = Sigma - Z Z', Z = Sigma(:,idx) inv(L)' (see ekf_tiled.cu /
// DESIGN.md section 2): y rides through the forward substitution as panel row N (v = inv(L) y), the backward
// substitution, W and the second phase of the covariance update are skipped.
__device__ __forceinline__ bool reduced_update(const EkfPtrs& p, int f) { return p.asym[f] == 0 && !(p.flags & EKFVIO_FLAG_LITERAL_JOSEPH); }

// ---- measurement map + residual (one warp per filter) ---------------------------------------
__global__ void ekf_large_idx(EkfPtrs p, const double* __restrict__ z, const double* __restrict__ Rin, const uint8_t* __restrict__ pass) {
    const int f = blockIdx.x, lane = threadIdx.x;
    const int n = p.nfeat[f], nmax = p.nmax;
    const double* zf = z + (size_t)f * nmax * 2;
    const double* Rf = Rin + (size_t)f * nmax * 4;
    const uint8_t* pf = pass + (size_t)f * nmax;
    const double* feat_g = p.feat + (size_t)f * nmax * 3;
    int* idx_g = p.idx + (size_t)f * p.mmax;
    double* y_g = p.y + (size_t)f * p.mmax;
    int m = 0;
    for (int base = 0; base < n; base += 32) {   // formFeatureMeasurementMap (:634-661) + bookkeeping of :506-529
        int i = base + lane;
        bool pr = i < n && pf[i] != 0;
        unsigned mask = __ballot_sync(0xffffffffu, pr);
        int pos = m + 2 * __popc(mask & ((1u << lane) - 1u));
        if (pr) {
            idx_g[pos] = BASE + 3 * i; idx_g[pos + 1] = BASE + 3 * i + 1;
            double zx = zf[2 * i], zy = zf[2 * i + 1];
            y_g[pos] = zx - feat_g[3 * i];
            y_g[pos + 1] = zy - feat_g[3 * i + 1];
            p.klt_last[((size_t)f * nmax + i) * 2] = zx;
            p.klt_last[((size_t)f * nmax + i) * 2 + 1] = zy;
            if (Rf[4 * i + 1] != Rf[4 * i + 2]) p.asym[f] = 1;
        } else if (i < n) {
            p.dflags[(size_t)f * nmax + i] = 1;
        }
        m += 2 * __popc(mask);
    }
    if (lane == 0) p.m[f] = m;
    if (m == 0 && lane == 0) {   // K is N x 0: only the quaternion renormalisation of :605-609 acts
        double* mu_g = p.mu + (size_t)f * BASE;
        double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
        mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
    }
}

// ---- S (full), lower(L) <- upper(S), C = Sigma(:,idx) into the K and W panels ----------------------
__global__ void __launch_bounds__(256) ekf_large_gather(EkfPtrs p, LargePtrs lp, const double* __restrict__ Pin, const double* __restrict__ Rin) {
    const int f = blockIdx.x, part = blockIdx.y, nparts = gridDim.y, tid = threadIdx.x;
    const int m = p.m[f];
    if (m == 0) return;
    const int N = BASE + 3 * p.nfeat[f], ld = p.ldP, mp = lp.mp;
    const int me = ((m + BLK - 1) / BLK) * BLK;                 // extent with identity tail
    const double* Pi = Pin + (size_t)f * ld * ld;
    const double* Rf = Rin + (size_t)f * p.nmax * 4;
    const int* idx = p.idx + (size_t)f * p.mmax;
    double* Sg = lp.S + (size_t)f * mp * mp;
    double* Lg = lp.L + (size_t)f * mp * mp;
    double* Kf = p.K + (size_t)f * ld * p.ldK;
    double* Wf = p.W + (size_t)f * ld * p.ldK;
    // S in 16 x 32 tiles, one warp each: rows of S and rows of lower(L) are both written in full
    // 128-byte lines (the transpose goes through shared memory).
    __shared__ double tile[8][16][33];
    {
        const int lane = tid & 31, warp = tid >> 5;
        const int tcols = me / 32, trows = me / 16;
        for (int t = part * 8 + warp; t < trows * tcols; t += nparts * 8) {
            const int a0 = (t / tcols) * 16, b0 = (t % tcols) * 32;
            const int b = b0 + lane;
            const int cb = b < m ? idx[b] : 0;
#pragma unroll 4
            for (int aa = 0; aa < 16; ++aa) {
                const int a = a0 + aa;
                double v;
                if (a < m && b < m) {
                    const int ra = idx[a];
                    v = Pi[(size_t)ra * ld + cb];
                    if ((a >> 1) == (b >> 1)) v += Rf[4 * ((ra - BASE) / 3) + (a & 1) * 2 + (b & 1)];
                } else v = (a == b) ? 1.0 : 0.0;
                Sg[(size_t)a * mp + b] = v;
                tile[warp][aa][lane] = v;
            }
            __syncwarp();
            if (b0 + 31 >= a0) {                                  // lower(L)(b,a) <- upper(S)(a,b)
                const int aa = lane & 15;
#pragma unroll 4
                for (int bb = lane >> 4; bb < 32; bb += 2)
                    if (b0 + bb >= a0 + aa) Lg[(size_t)(b0 + bb) * mp + a0 + aa] = tile[warp][aa][bb];
            }
            __syncwarp();
        }
    }
    const bool red = reduced_update(p, f);
    const double* y = p.y + (size_t)f * p.mmax;
    const int Nr = N + (red ? 1 : 0);                            // reduced update: row N of the panel is y
    const int Ne = ((Nr + BLK - 1) / BLK) * BLK < ld ? ((Nr + BLK - 1) / BLK) * BLK : ld;
    for (int i = part; i < Ne; i += nparts) {                    // C rows (zero beyond N / m)
        for (int b = tid; b < me; b += 256) {
            double v = (i < N && b < m) ? Pi[(size_t)i * ld + idx[b]] : 0.0;
            if (red && i == N && b < m) v = y[b];
            size_t o = kw_at(ld, i, b);
            Kf[o] = v;
            if (!red) Wf[o] = v;
        }
    }
}

// ---- diagonal block: Cholesky of the 64x64 block jb of L, tiles + inverse tiles to scratch --------
__global__ void __launch_bounds__(128) ekf_large_potrf(EkfPtrs p, LargePtrs lp, int jb) {
    __shared__ __align__(16) double Ls[NT8 * 64];
    __shared__ __align__(16) double Li[NB8 * 64];
    __shared__ int s_bad;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int m = p.m[f];
    if (jb * BLK >= m) return;
    const int mp = lp.mp;
    double* Lg = lp.L + (size_t)f * mp * mp + (size_t)jb * BLK * mp + jb * BLK;
    if (tid == 0) s_bad = 0;
    for (int e = tid; e < NT8 * 64; e += 128) {
        int t = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
        int ib = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
        while (ib * (ib + 1) / 2 > t) --ib;
        while ((ib + 1) * (ib + 2) / 2 <= t) ++ib;
        int kb = t - ib * (ib + 1) / 2;
        int a = ib * 8 + rr, b = kb * 8 + cc;
        Ls[t * 64 + tsw(rr, cc)] = (a >= b) ? Lg[(size_t)a * mp + b] : 0.0;
    }
    __syncthreads();
    chol_tiles<4>(Ls, Li, NB8, &s_bad);
    if (tid == 0 && s_bad) atomicOr(&p.status[f], 1);
    double* Tg = lp.T + ((size_t)f * lp.nblk + jb) * (NT8 + NB8) * 64;
    for (int e = tid; e < NT8 * 64; e += 128) Tg[e] = Ls[e];
    for (int e = tid; e < NB8 * 64; e += 128) Tg[NT8 * 64 + e] = Li[e];
    for (int e = tid; e < BLK * BLK; e += 128) {             // factored block back into L (row-major, lower)
        int a = e >> 6, b = e & 63;
        if (a >= b) Lg[(size_t)a * mp + b] = Ls[tile_of(a >> 3, b >> 3) + tsw(a & 7, b & 7)];
    }
}

// ---- triangular solve of 16-row strips against the diagonal block jb -----------------------------
//   MODE 0 (PANEL): rows below the block in L (row-major):   X Ljj' = A          (Cholesky panel)
//   MODE 1 (FWD)  : rows of the K panel (chunk-major):       Z Ljj' = C_jb       (forward substitution)
//   MODE 2 (BWD)  : rows of the K panel:                     K Ljj  = Z_jb       (backward substitution)
template <int MODE>
__global__ void __launch_bounds__(256) ekf_large_trsm(EkfPtrs p, LargePtrs lp, int jb) {
    __shared__ __align__(16) double Ls[NT8 * 64];
    __shared__ __align__(16) double Li[NB8 * 64];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m[f];
    if (jb * BLK >= m) return;
    const int N = BASE + 3 * p.nfeat[f], ld = p.ldP, mp = lp.mp;
    const int me = ((m + BLK - 1) / BLK) * BLK;
    if (MODE == 2 && reduced_update(p, f)) return;
    const int row_begin = MODE == 0 ? (jb + 1) * BLK : 0;
    const int row_end = MODE == 0 ? me : (MODE == 1 && reduced_update(p, f) ? N + 1 : N);
    const int i0 = row_begin + (blockIdx.x * 8 + warp) * 16;
    if (row_begin + blockIdx.x * 128 >= row_end) return;
    const double* Tg = lp.T + ((size_t)f * lp.nblk + jb) * (NT8 + NB8) * 64;
    for (int e = tid * 2; e < NT8 * 64; e += 512) cp_async16(Ls + e, Tg + e);
    for (int e = tid * 2; e < NB8 * 64; e += 512) cp_async16(Li + e, Tg + NT8 * 64 + e);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (i0 >= row_end) return;
    const int r = lane >> 2, q = lane & 3;
    double* Lg = lp.L + (size_t)f * mp * mp;
    double* Kf = p.K + (size_t)f * ld * p.ldK;
    const int c0 = jb * BLK;
    double k0[2][NB8], k1[2][NB8];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int row = i0 + rt * 8 + r;
#pragma unroll
        for (int t = 0; t < NB8; ++t) {
            double2 v = make_double2(0.0, 0.0);
            if (row < row_end) v = MODE == 0 ? *reinterpret_cast<const double2*>(Lg + (size_t)row * mp + c0 + t * 8 + 2 * q)
                                            : *reinterpret_cast<const double2*>(Kf + kw_at(ld, row, c0 + t * 8 + 2 * q));
            k0[rt][t] = v.x; k1[rt][t] = v.y;
        }
    }
    if (MODE != 2) {
#pragma unroll
        for (int t = 0; t < NB8; ++t) {
            const double* I8 = Li + t * 64;
            const double b0 = I8[tsw(r, q)], b1 = I8[tsw(r, 4 + q)];
            double za[2][2];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                double a0, a1;
                cfrag_to_afrag(k0[rt][t], k1[rt][t], lane, a0, a1);
                double t0 = 0.0, t1 = 0.0;
                dmma884(t0, t1, a0, b0);
                dmma884(t0, t1, a1, b1);
                k0[rt][t] = t0; k1[rt][t] = t1;
                cfrag_to_afrag(t0, t1, lane, za[rt][0], za[rt][1]);
            }
#pragma unroll
            for (int t2 = 0; t2 < NB8; ++t2) {
                if (t2 > t) {
                    const double* T = Ls + tile_of(t2, t);
                    const double l0 = T[tsw(r, q)], l1 = T[tsw(r, 4 + q)];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        dmma884(k0[rt][t2], k1[rt][t2], -za[rt][0], l0);
                        dmma884(k0[rt][t2], k1[rt][t2], -za[rt][1], l1);
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int tr = 0; tr < NB8; ++tr) {
            const int t = NB8 - 1 - tr;
            const double* I8 = Li + t * 64;
            const double b0 = I8[tsw(q, r)], b1 = I8[tsw(4 + q, r)];
            double ka[2][2];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                double a0, a1;
                cfrag_to_afrag(k0[rt][t], k1[rt][t], lane, a0, a1);
                double t0 = 0.0, t1 = 0.0;
                dmma884(t0, t1, a0, b0);
                dmma884(t0, t1, a1, b1);
                k0[rt][t] = t0; k1[rt][t] = t1;
                cfrag_to_afrag(t0, t1, lane, ka[rt][0], ka[rt][1]);
            }
#pragma unroll
            for (int t2 = 0; t2 < NB8; ++t2) {
                if (t2 < t) {
                    const double* T = Ls + tile_of(t, t2);
                    const double l0 = T[tsw(q, r)], l1 = T[tsw(4 + q, r)];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        dmma884(k0[rt][t2], k1[rt][t2], -ka[rt][0], l0);
                        dmma884(k0[rt][t2], k1[rt][t2], -ka[rt][1], l1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int row = i0 + rt * 8 + r;
        if (row >= row_end) continue;
#pragma unroll
        for (int t = 0; t < NB8; ++t) {
            double2 v = make_double2(k0[rt][t], k1[rt][t]);
            if (MODE == 0) *reinterpret_cast<double2*>(Lg + (size_t)row * mp + c0 + t * 8 + 2 * q) = v;
            else *reinterpret_cast<double2*>(Kf + kw_at(ld, row, c0 + t * 8 + 2 * q)) = v;
        }
    }
}

// ---- sparseView on K (:580), mu += K y (:600), quaternion renormalisation (:605-609) ---------------
__global__ void __launch_bounds__(256) ekf_large_finalize(EkfPtrs p) {
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m[f];
    if (m == 0) return;
    const int N = BASE + 3 * p.nfeat[f], ld = p.ldP;
    double* Kf = p.K + (size_t)f * ld * p.ldK;
    const double* y = p.y + (size_t)f * p.mmax;
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * p.nmax * 3;
    // base rows (and the normalisation below) belong to part 0; feature rows are dealt over all parts
    const int part = blockIdx.y, nparts = gridDim.y;
    const int nbase_rounds = part == 0 ? (BASE + 7) / 8 : 0;
    const int nfrows = N - BASE;
    const int nfeat_rounds = (nfrows + 8 * nparts - 1) / (8 * nparts);
    const bool red = reduced_update(p, f);                       // then the panel holds Z and its row N is v = inv(L) y
    for (int r = 0; r < nbase_rounds + nfeat_rounds; ++r) {
        const int i = r < nbase_rounds ? r * 8 + warp : BASE + ((r - nbase_rounds) * nparts + part) * 8 + warp;
        if (r < nbase_rounds ? i >= BASE : i >= N) continue;
        double dot = 0.0;
        for (int k = lane; k < m; k += 32) {
            size_t o = kw_at(ld, i, k);
            if (red) dot += Kf[o] * Kf[kw_at(ld, N, k)];
            else {
                double v = prune(Kf[o]);
                Kf[o] = v;
                dot += v * y[k];
            }
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (lane == 0) { if (i < BASE) mu_g[i] += dot; else feat_g[i - BASE] += dot; }
    }
    __syncthreads();
    if (tid == 0 && part == 0) {
        double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
        mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
        bool fin = true;
        for (int i = 0; i < BASE; ++i) fin = fin && isfinite(mu_g[i]);
        if (!fin) atomicOr(&p.status[f], 2);
    }
}

// ---- the 64x64-tile DMMA GEMM:  D = C - A B'  (one or two A/B phases) ---------------------------
enum { OP_SYRK = 0, OP_FWDUPD = 1, OP_BWDUPD = 2, OP_W = 3, OP_JOSEPH = 4 };
constexpr int GKC = 16, GLDA = GKC + 4, GLDB2 = BLK + 4, GNST = 3;
constexpr int G_A = BLK * GLDA;                                   // A stage [64][20]
constexpr int G_B = (BLK * GLDA > GKC * GLDB2) ? BLK * GLDA : GKC * GLDB2;
constexpr int G_STAGE = G_A + G_B;

// Operand access: mode 0 row-major k-contiguous (base + row*ld + k), mode 1 chunk-major panel
// (kw_at(ldP, row, k)), mode 2 row(k)-major j-contiguous (base + krow*ld + r0 + j, krow optionally gathered;
// r0 is then the column offset of the tile).
struct Opnd { const double* base; int mode; int ld; int r0; int k0; const int* gather; };

__global__ void __launch_bounds__(256) ekf_large_gemm(EkfPtrs p, LargePtrs lp, const double* __restrict__ Pin, double* __restrict__ Pout, int op, int jb,
                                                      int sym) {
    extern __shared__ __align__(16) double smg[];
    const int f = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m[f];
    const int N = BASE + 3 * p.nfeat[f], ld = p.ldP, mp = lp.mp;
    const int nblk = (m + BLK - 1) / BLK, nrt = (N + BLK - 1) / BLK;
    if (op != OP_JOSEPH && (m == 0 || jb >= nblk)) return;
    const bool red = reduced_update(p, f);
    if (red && (op == OP_BWDUPD || op == OP_W)) return;
    if (op == OP_JOSEPH && sym >= 0 && (p.asym[f] != 0) != (sym == 0)) return;   // sym=1: symmetric filters, sym=0: the others, sym=-1: all
    const double* Lg = lp.L + (size_t)f * mp * mp;
    const double* Sg = lp.S + (size_t)f * mp * mp;
    double* Kf = p.K + (size_t)f * ld * p.ldK;
    double* Wf = p.W + (size_t)f * ld * p.ldK;
    const double* Pi = Pin ? Pin + (size_t)f * ld * ld : nullptr;
    double* Po = Pout ? Pout + (size_t)f * ld * ld : nullptr;
    const int* idx = p.idx + (size_t)f * p.mmax;

    // tile coordinates and operand descriptors per operation
    int ti, tj, nphase = 1, Kd[2] = {BLK, 0};
    Opnd A[2], B[2];
    const double* C0; double* D; int cmode, cld, ci0, cj0, row_lim, col_lim;
    bool tri = false, mirror = false, do_prune = false;
    if (op == OP_SYRK) {                 // L(ti,tk) -= L(ti,jb) L(tk,jb)',  ti >= tk > jb
        const int t = nblk - 1 - jb, e = blockIdx.x;
        if (e >= t * (t + 1) / 2) return;
        int ii = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
        while (ii * (ii + 1) / 2 > e) --ii;
        while ((ii + 1) * (ii + 2) / 2 <= e) ++ii;
        ti = jb + 1 + ii; tj = jb + 1 + (e - ii * (ii + 1) / 2);
        A[0] = {Lg, 0, mp, ti * BLK, jb * BLK, nullptr}; B[0] = {Lg, 0, mp, tj * BLK, jb * BLK, nullptr};
        C0 = Lg; D = const_cast<double*>(Lg); cmode = 0; cld = mp; ci0 = ti * BLK; cj0 = tj * BLK; row_lim = nblk * BLK; col_lim = nblk * BLK;
    } else if (op == OP_FWDUPD || op == OP_BWDUPD) {   // K(:,j') -= Z(:,jb) L(j',jb)'  |  Z(:,j') -= K(:,jb) L(jb,j')
        ti = blockIdx.x; tj = op == OP_FWDUPD ? jb + 1 + blockIdx.y : blockIdx.y;
        if (ti >= (N + (red ? 1 : 0) + BLK - 1) / BLK || (op == OP_FWDUPD ? tj >= nblk : tj >= jb)) return;   // (reduced update: row N carries y)
        A[0] = {Kf, 1, ld, ti * BLK, jb * BLK, nullptr};
        if (op == OP_FWDUPD) B[0] = {Lg, 0, mp, tj * BLK, jb * BLK, nullptr};
        else B[0] = {Lg, 2, mp, tj * BLK, jb * BLK, nullptr};
        C0 = Kf; D = Kf; cmode = 1; cld = ld; ci0 = ti * BLK; cj0 = tj * BLK; row_lim = ld; col_lim = nblk * BLK;
    } else if (op == OP_W) {              // W(:,j') -= K S(:,j')
        ti = blockIdx.x; tj = blockIdx.y;
        if (ti >= nrt || tj >= nblk) return;
        A[0] = {Kf, 1, ld, ti * BLK, 0, nullptr}; B[0] = {Sg, 2, mp, tj * BLK, 0, nullptr};
        Kd[0] = nblk * BLK;
        C0 = Wf; D = Wf; cmode = 1; cld = ld; ci0 = ti * BLK; cj0 = tj * BLK; row_lim = ld; col_lim = nblk * BLK;
    } else {                              // Sigma'(ti,tj) = Sigma - K Sigma(idx,:) - W K'
        tri = sym == 1;
        if (tri) {
            const int e = blockIdx.x;
            if (e >= nrt * (nrt + 1) / 2) return;
            int ii = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
            while (ii * (ii + 1) / 2 > e) --ii;
            while ((ii + 1) * (ii + 2) / 2 <= e) ++ii;
            ti = ii; tj = e - ii * (ii + 1) / 2;
        } else {
            ti = blockIdx.x / lp.nrt_max; tj = blockIdx.x % lp.nrt_max;
            if (ti >= nrt || tj >= nrt) return;
        }
        mirror = tri; do_prune = true;
        nphase = 2; Kd[0] = Kd[1] = ((m + GKC - 1) / GKC) * GKC;
        A[0] = {Kf, 1, ld, ti * BLK, 0, nullptr}; B[0] = {Pi, 2, ld, tj * BLK, 0, idx};
        A[1] = {Wf, 1, ld, ti * BLK, 0, nullptr}; B[1] = {Kf, 1, ld, tj * BLK, 0, nullptr};
        if (red) { nphase = 1; B[0] = {Kf, 1, ld, tj * BLK, 0, nullptr}; }          // Sigma - Z Z'
        C0 = Pi; D = Po; cmode = 0; cld = ld; ci0 = ti * BLK; cj0 = tj * BLK; row_lim = N; col_lim = N;
    }

    const int r = lane >> 2, q = lane & 3;
    const int wr = (warp & 3) * 16, wc = (warp >> 2) * 32;           // warp sub-tile: 16 rows x 32 cols
    double c0[2][4], c1[2][4];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int row = ci0 + wr + rt * 8 + r, col = cj0 + wc + ct * 8 + 2 * q;
            double2 v = make_double2(0.0, 0.0);
            if (row < (cmode == 0 ? cld : ld) && col + 1 < (cmode == 0 ? cld : p.ldK))
                v = cmode == 0 ? *reinterpret_cast<const double2*>(C0 + (size_t)row * cld + col) : *reinterpret_cast<const double2*>(C0 + kw_at(ld, row, col));
            c0[rt][ct] = v.x; c1[rt][ct] = v.y;
        }

    for (int ph = 0; ph < nphase; ++ph) {
        const Opnd a = A[ph], b = B[ph];
        const int nch = Kd[ph] / GKC;
        auto stage = [&](int c) {
            if (c < nch) {
                double* As = smg + (c % GNST) * G_STAGE;
                double* Bs = As + G_A;
                const int k0 = c * GKC;
                for (int t = tid; t < BLK * (GKC / 2); t += 256) {       // A: 64 rows x 16 k
                    const int row = t >> 3, seg = t & 7;
                    double* dst = &As[row * GLDA + seg * 2];
                    const int gr = a.r0 + row, gk = a.k0 + k0 + seg * 2;
                    const bool ok = gr < (a.mode == 1 ? ld : a.ld) && gk < (a.mode == 1 ? p.ldK : a.ld);
                    if (ok) cp_async16(dst, a.mode == 1 ? a.base + kw_at(ld, gr, gk) : a.base + (size_t)gr * a.ld + gk);
                    else { dst[0] = 0.0; dst[1] = 0.0; }
                }
                if (b.mode != 2) {                                        // B k-contiguous: 64 rows(j) x 16 k
                    for (int t = tid; t < BLK * (GKC / 2); t += 256) {
                        const int row = t >> 3, seg = t & 7;
                        double* dst = &Bs[row * GLDA + seg * 2];
                        const int gr = b.r0 + row, gk = b.k0 + k0 + seg * 2;
                        const bool ok = gr < (b.mode == 1 ? ld : b.ld) && gk < (b.mode == 1 ? p.ldK : b.ld);
                        if (ok) cp_async16(dst, b.mode == 1 ? b.base + kw_at(ld, gr, gk) : b.base + (size_t)gr * b.ld + gk);
                        else { dst[0] = 0.0; dst[1] = 0.0; }
                    }
                } else {                                                  // B j-contiguous: 16 k-rows x 64 j
                    for (int t = tid; t < GKC * (BLK / 2); t += 256) {
                        const int k = t >> 5, seg = t & 31;
                        double* dst = &Bs[k * GLDB2 + seg * 2];
                        const int gk = b.k0 + k0 + k;
                        const bool ok = b.gather ? gk < m : gk < b.ld;
                        const int krow = ok ? (b.gather ? b.gather[gk] : gk) : 0;
                        const bool cok = b.r0 + seg * 2 + 1 < b.ld;
                        if (ok && cok) cp_async16(dst, b.base + (size_t)krow * b.ld + b.r0 + seg * 2);
                        else { dst[0] = 0.0; dst[1] = 0.0; }
                    }
                }
            }
            cp_async_commit();
        };
        stage(0); stage(1);
        for (int c = 0; c < nch; ++c) {
            cp_async_wait<GNST - 2>();
            __syncthreads();
            stage(c + 2);
            const double* As = smg + (c % GNST) * G_STAGE;
            const double* Bs = As + G_A;
#pragma unroll
            for (int kk = 0; kk < GKC / 4; ++kk) {
                double fa[2], fb[4];
                fa[0] = As[(wr + r) * GLDA + kk * 4 + q];
                fa[1] = As[(wr + 8 + r) * GLDA + kk * 4 + q];
#pragma unroll
                for (int ct = 0; ct < 4; ++ct)
                    fb[ct] = b.mode != 2 ? Bs[(wc + ct * 8 + r) * GLDA + kk * 4 + q] : Bs[(kk * 4 + q) * GLDB2 + wc + ct * 8 + r];
#pragma unroll
                for (int ct = 0; ct < 4; ++ct) {
                    dmma884(c0[0][ct], c1[0][ct], -fa[0], fb[ct]);
                    dmma884(c0[1][ct], c1[1][ct], -fa[1], fb[ct]);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

#pragma unroll
    for (int rt = 0; rt < 2; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) {
            const int row = ci0 + wr + rt * 8 + r, col = cj0 + wc + ct * 8 + 2 * q;
            double v0 = c0[rt][ct], v1 = c1[rt][ct];
            if (do_prune) { v0 = prune(v0); v1 = prune(v1); }
            if (row >= row_lim) continue;
            if (cmode == 1) {
                if (col < col_lim) *reinterpret_cast<double2*>(D + kw_at(ld, row, col)) = make_double2(v0, v1);
            } else if (!tri || ti != tj) {
                if (col + 1 < col_lim) *reinterpret_cast<double2*>(D + (size_t)row * cld + col) = make_double2(v0, v1);
                else if (col < col_lim) D[(size_t)row * cld + col] = v0;
                if (mirror) {
                    if (col < col_lim) D[(size_t)col * cld + row] = v0;
                    if (col + 1 < col_lim) D[(size_t)(col + 1) * cld + row] = v1;
                }
            } else {                       // diagonal tile of the symmetric update: lower part + mirror
                if (col <= row && col < col_lim) { D[(size_t)row * cld + col] = v0; if (col != row) D[(size_t)col * cld + row] = v0; }
                if (col + 1 <= row && col + 1 < col_lim) { D[(size_t)row * cld + col + 1] = v1; if (col + 1 != row) D[(size_t)(col + 1) * cld + row] = v1; }
            }
        }
}

}  // namespace

namespace ekfvio {

size_t large_scratch_doubles_S(int mmax) { size_t mp = ((size_t)mmax + BLK - 1) / BLK * BLK; return mp * mp; }
size_t large_scratch_doubles_T(int mmax) { size_t nb = ((size_t)mmax + BLK - 1) / BLK; return nb * (NT8 + NB8) * 64; }

// The whole measurement update for large states.  `timer_mark(slot)` lets the API layer time the
// three stages with the same slots as the small path (1 factorisation, 3 gain, 2 covariance).
cudaError_t launch_update_large(const EkfPtrs& p, const LargePtrs& lp, const double* Pin, double* Pout, const double* z, const double* R,
                                const uint8_t* pass, cudaStream_t st, long long* launches, KernelTimer* timer) {
    auto mark = [&](void*, int slot, cudaStream_t s) { if (timer) { timer->end(s); if (slot >= 0) timer->begin(slot, s); } };
    void* mark_ctx = nullptr;
    static bool configured_on[64] = {false};
    bool& configured = configured_on[current_device_slot()];
    const size_t sm = (size_t)GNST * G_STAGE * sizeof(double);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_large_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int F = p.F, nblk = lp.nblk, nrt = lp.nrt_max;
    long long n = 0;
    mark(mark_ctx, 1, st);
    ekf_large_idx<<<F, 32, 0, st>>>(p, z, R, pass); ++n;
    ekf_large_gather<<<dim3(F, 16), 256, 0, st>>>(p, lp, Pin, R); ++n;
    for (int jb = 0; jb < nblk; ++jb) {
        ekf_large_potrf<<<F, 128, 0, st>>>(p, lp, jb); ++n;
        const int rows_below = (nblk - jb - 1) * BLK;
        if (rows_below > 0) {
            ekf_large_trsm<0><<<dim3((rows_below + 127) / 128, F), 256, 0, st>>>(p, lp, jb); ++n;
            const int t = nblk - 1 - jb;
            ekf_large_gemm<<<dim3(t * (t + 1) / 2, 1, F), 256, sm, st>>>(p, lp, nullptr, nullptr, OP_SYRK, jb, 0); ++n;
        }
    }
    mark(mark_ctx, 3, st);
    const int rowgrp = (p.Nmax + 1 + 127) / 128;                 // + the y row of the reduced update
    for (int jb = 0; jb < nblk; ++jb) {
        ekf_large_trsm<1><<<dim3(rowgrp, F), 256, 0, st>>>(p, lp, jb); ++n;
        if (jb + 1 < nblk) { ekf_large_gemm<<<dim3(nrt, nblk - jb - 1, F), 256, sm, st>>>(p, lp, nullptr, nullptr, OP_FWDUPD, jb, 0); ++n; }
    }
    for (int jb = nblk - 1; jb >= 0; --jb) {
        ekf_large_trsm<2><<<dim3(rowgrp, F), 256, 0, st>>>(p, lp, jb); ++n;
        if (jb > 0) { ekf_large_gemm<<<dim3(nrt, jb, F), 256, sm, st>>>(p, lp, nullptr, nullptr, OP_BWDUPD, jb, 0); ++n; }
    }
    ekf_large_finalize<<<dim3(F, 16), 256, 0, st>>>(p); ++n;
    ekf_large_gemm<<<dim3(nrt, nblk, F), 256, sm, st>>>(p, lp, nullptr, nullptr, OP_W, 0, 0); ++n;
    mark(mark_ctx, 2, st);
    ekf_large_gemm<<<dim3(nrt * (nrt + 1) / 2, 1, F), 256, sm, st>>>(p, lp, Pin, Pout, OP_JOSEPH, 0, 1); ++n;   // symmetric filters
    ekf_large_gemm<<<dim3(nrt * nrt, 1, F), 256, sm, st>>>(p, lp, Pin, Pout, OP_JOSEPH, 0, 0); ++n;            // asymmetric ones (early exit otherwise)
    mark(mark_ctx, -1, st);
    if (launches) *launches += n;
    return cudaGetLastError();
}

}  // namespace ekfvio
