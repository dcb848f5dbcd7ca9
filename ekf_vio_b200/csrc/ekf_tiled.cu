// Tiled FP64 tensor-core (DMMA.8x8x4) kernels for the EKF measurement update on sm_100a — the
// fast path for feature-augmented states up to N = 22 + 3n <= 256.
//
//   ekf_joseph_tiled : Sigma' = Sigma - K Sigma(idx,:) - W K'     (TightlyCoupledEKF.cpp:586-596,
//                      625 in selection form; W = Sigma(:,idx) - K S is the Joseph residual panel)
//
// One CTA per filter, one warp per 16-row strip of Sigma'; the strip's accumulators stay in
// registers (2 x 11 DMMA tiles per pass), the B operands stream through shared memory in 8-deep
// k-chunks with cp.async double buffering, the A operands come straight from global memory one
// chunk ahead.  Shared-memory leading dimensions are == 4 (mod 8) doubles so every fragment load
// is bank-conflict free.
#include <cstdlib>

#include "ekf_common.cuh"
#include "ekf_kernels.h"
#include "ekf_tiles.cuh"

using namespace ekfvio;

namespace {

#ifdef EKFVIO_PROFILE_CLOCKS
__device__ unsigned long long g_clk[8];
#define CLK_MARK(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_clk[i], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define CLK_MARK(i) do {} while (0)
#endif

constexpr int JT = 11;              // column tiles per pass
constexpr int KC = 16;              // k-chunk depth
constexpr int NST = 3;              // cp.async stages
constexpr int LDG = JT * 8 + 4;     // G-phase chunk: [KC][LDG]
constexpr int LDK = KC + 4;         // K-phase chunk: [JT*8][LDK]
constexpr int BS_DOUBLES = (JT * 8 * LDK > KC * LDG) ? JT * 8 * LDK : KC * LDG;

template <int NW>
__device__ __forceinline__ void joseph_tiled_filter(const EkfPtrs& p, const double* __restrict__ Pin, double* __restrict__ Pout, int f) {
    extern __shared__ __align__(16) double Bs_all[];   // NST * BS_DOUBLES
    __shared__ int s_idx[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.nfeat[f], N = BASE + 3 * n, m = p.m[f];
    const int ld = p.ldP, ldK = p.ldK;
    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Po = Pout + (size_t)f * ld * ld;
    const double* Kf = p.K + (size_t)f * ld * ldK;
    const double* Wf = p.W + (size_t)f * ld * ldK;
    const int* idx = p.idx + (size_t)f * p.mmax;
    for (int i = tid; i < m; i += NW * 32) s_idx[i] = idx[i];
    __syncthreads();

    const int r = lane >> 2, q = lane & 3;
    const int i0 = warp * 16;
    const bool active = i0 < N;
    const int nct = (N + 7) >> 3, npass = (nct + JT - 1) / JT, nch = (m + KC - 1) / KC;
    const int row0 = min(i0 + r, ld - 1), row1 = min(i0 + 8 + r, ld - 1);

    for (int pass = 0; pass < npass; ++pass) {
        const int j0 = pass * JT * 8;
        double c0[2][JT], c1[2][JT];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int t = 0; t < JT; ++t) {
                int row = rt ? row1 : row0, col = j0 + t * 8 + 2 * q;
                double2 v = make_double2(0.0, 0.0);
                if (active && col < ld) v = *reinterpret_cast<const double2*>(Pi + (size_t)row * ld + col);
                c0[rt][t] = v.x; c1[rt][t] = v.y;
            }
        for (int phase = 0; phase < 2; ++phase) {
            const double* Am = phase == 0 ? Kf : Wf;
            const double* A0 = Am + (size_t)row0 * 16 + q;   // + chunk * ld*16 (chunk-major panels, KC == 16)
            const double* A1 = Am + (size_t)row1 * 16 + q;
            auto stage = [&](int c) {
                double* B = Bs_all + (c % NST) * BS_DOUBLES;
                const int k0 = c * KC;
                if (c < nch) {
                    if (phase == 0) {   // rows idx[k] of Sigma, columns j0 .. j0 + 87
                        for (int t = tid; t < KC * (JT * 4); t += NW * 32) {
                            int k = t / (JT * 4), seg = t % (JT * 4);
                            double* dst = &B[k * LDG + seg * 2];
                            int col = j0 + seg * 2;
                            if (k0 + k < m && col < ld) cp_async16(dst, Pi + (size_t)s_idx[k0 + k] * ld + col);
                            else { dst[0] = 0.0; dst[1] = 0.0; }
                        }
                    } else {            // rows j0 .. j0 + 87 of K, columns k0 .. k0 + KC-1
                        for (int t = tid; t < JT * 8 * (KC / 2); t += NW * 32) {
                            int col = t / (KC / 2), seg = t % (KC / 2);
                            double* dst = &B[col * LDK + seg * 2];
                            if (j0 + col < ld && k0 + seg * 2 < m) cp_async16(dst, Kf + kw_at(ld, j0 + col, k0 + seg * 2));   // m is even; k >= m must read as zero
                            else { dst[0] = 0.0; dst[1] = 0.0; }
                        }
                    }
                }
                cp_async_commit();   // one group per stage call, possibly empty: keeps the wait count uniform
            };
            auto load_a = [&](int c, double (&a)[2][KC / 4]) {
                const int k0 = c * KC;
#pragma unroll
                for (int kk = 0; kk < KC / 4; ++kk) {
                    bool ok = active && c < nch && k0 + kk * 4 + q < ldK;
                    a[0][kk] = ok ? A0[(size_t)c * ld * 16 + kk * 4] : 0.0;
                    a[1][kk] = ok ? A1[(size_t)c * ld * 16 + kk * 4] : 0.0;
                }
            };
            double a_cur[2][KC / 4], a_nxt[2][KC / 4];
            stage(0); stage(1);
            load_a(0, a_cur);
            for (int c = 0; c < nch; ++c) {
                cp_async_wait<NST - 2>();      // chunk c has landed (one younger group may be in flight)
                __syncthreads();               // ... for everyone, and everyone is done with chunk c-1
                stage(c + 2);                  // refills the buffer chunk c-1 used
                load_a(c + 1, a_nxt);
                if (active) {
                    const double* B = Bs_all + (c % NST) * BS_DOUBLES;
#pragma unroll
                    for (int kk = 0; kk < KC / 4; ++kk) {
#pragma unroll
                        for (int t = 0; t < JT; ++t) {
                            double b = (phase == 0) ? B[(kk * 4 + q) * LDG + t * 8 + r] : B[(t * 8 + r) * LDK + kk * 4 + q];
                            dmma884(c0[0][t], c1[0][t], -a_cur[0][kk], b);
                            dmma884(c0[1][t], c1[1][t], -a_cur[1][kk], b);
                        }
                    }
                }
#pragma unroll
                for (int kk = 0; kk < KC / 4; ++kk) { a_cur[0][kk] = a_nxt[0][kk]; a_cur[1][kk] = a_nxt[1][kk]; }
            }
            cp_async_wait<0>();
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int t = 0; t < JT; ++t) {
                    int row = i0 + rt * 8 + r, col = j0 + t * 8 + 2 * q;
                    if (row < N) {
                        double* o = Po + (size_t)row * ld + col;
                        if (col + 1 < N) *reinterpret_cast<double2*>(o) = make_double2(prune(c0[rt][t]), prune(c1[rt][t]));
                        else if (col < N) o[0] = prune(c0[rt][t]);
                    }
                }
        }
    }
}

// Persistent launch: a CTA walks over the filters and serves those the symmetric kernels leave out (only_asym) —
// when there are none, the whole launch is a few flag reads.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) ekf_joseph_tiled(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout, int only_asym) {
    const int total = p.fb ? p.fb[0] : p.F;                // behind ekf_update_fused: only the filters it left alone
    for (int li = blockIdx.x; li < total; li += gridDim.x) {
        const int f = p.fb ? p.fb[1 + li] : li;
        if (only_asym && p.route[f] != ROUTE_JOSEPH_FULL) continue;   // the symmetric routes went to ekf_joseph_sym
        joseph_tiled_filter<NW>(p, Pin, Pout, f);
        __syncthreads();                                  // shared memory is reused by the next filter
    }
}


// Completes the feature rows of a Sigma in lower form (row r valid up to its diagonal block) from the columns below the diagonal.
// The column reads are strided (one cache line per lane): eight of them are in flight per lane before the first store, because a
// single filter's latency through this kernel is what the launch behind ekf_update_fused costs.
template <int NWC>
__device__ __forceinline__ void mirror_lower_rows(double* Pw, int ld, int N, int warp, int lane) {
    for (int r = BASE + warp; r < N; r += NWC) {
        for (int c0 = r + 1 + lane; c0 < N; c0 += 256) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int c = c0 + 32 * k; v[k] = c < N ? Pw[(size_t)c * ld + r] : 0.0; }
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int c = c0 + 32 * k; if (c < N) Pw[(size_t)r * ld + c] = v[k]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ekf_gain_tiled: measurement map, residual, S, Cholesky, K, W, state update — the DMMA version
// of ekf_gain_general (TightlyCoupledEKF.cpp:475-580, 600-620).
//
// S (m x m, padded to 8*nb with an identity tail) lives in shared memory as 8x8 tiles, twice: Ss,
// the full S kept for W = Sigma(:,idx) - K S (no symmetry assumed: the Joseph form only cancels
// rounding errors of K if W is the residual against the very S that Sigma(idx,:) and Sigma(:,idx)
// define), and Ls, the lower triangle taken from upper(S) as the reference's LDLT does, factored
// in place (L L' = S).
// Tile element (r, c) sits at r*8 + (c ^ 4*((r>>1)&1)) so that both the row-fragment access
// (lane -> [lane/4][lane%4 + 4kk]) and the column-fragment access ([lane%4 + 4kk][lane/4]) hit 16
// distinct banks per half-warp.  Each warp then owns a 16-row strip of Sigma(:,idx) and carries it
// through both triangular solves entirely in registers (right-looking, 8-wide column blocks,
// diagonal blocks applied through their explicit 8x8 inverses).
// Kernel 1 of 2: measurement map, residual vector, lower(S) from upper(S), blocked right-looking
// Cholesky (8x8 tiles) and the explicit inverses of the diagonal tiles.  128 threads per filter
// and ~54 KB of shared memory, so four filters share an SM and hide each other's serial
// diagonal-tile steps.  L tiles and inverse tiles go to global scratch in their swizzled layout.
template <int NB, int NWC>
__device__ __forceinline__ void chol_tiled_filter(const EkfPtrs& p, const double* __restrict__ Pin, const double* __restrict__ z,
                                                  const double* __restrict__ Rin, const uint8_t* __restrict__ pass, const int f) {
    extern __shared__ __align__(16) double smc[];
    constexpr int NT = NB * (NB + 1) / 2;
    constexpr int NTH = NWC * 32;
    double* Ls = smc;                      // NT tiles
    double* Li = Ls + NT * 64;             // NB inverse diagonal tiles
    int* s_idx = reinterpret_cast<int*>(Li + NB * 64);   // NB*8
    __shared__ int s_m, s_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.nfeat[f];
    const int ld = p.ldP, nmax = p.nmax;
    const double* Pi = Pin + (size_t)f * ld * ld;
    const double* zf = z + (size_t)f * nmax * 2;
    const double* Rf = Rin + (size_t)f * nmax * 4;
    const uint8_t* pf = pass + (size_t)f * nmax;
    double* mu_g = p.mu + (size_t)f * BASE;
    const double* feat_g = p.feat + (size_t)f * nmax * 3;
    int* idx_g = p.idx + (size_t)f * p.mmax;
    double* y_g = p.y + (size_t)f * p.mmax;

    __shared__ int s_was_sym;
    if (warp == 0) {  // formFeatureMeasurementMap (:634-661) + bookkeeping of :506-529, ballot-compacted
        if (lane == 0) s_was_sym = p.asym[f] == 0;
        __syncwarp();
        int m = 0;
        for (int base = 0; base < n; base += 32) {
            int i = base + lane;
            bool pr = i < n && pf[i] != 0;
            unsigned mask = __ballot_sync(0xffffffffu, pr);
            int pos = m + 2 * __popc(mask & ((1u << lane) - 1u));
            if (pr) {
                s_idx[pos] = BASE + 3 * i; s_idx[pos + 1] = BASE + 3 * i + 1;
                idx_g[pos] = BASE + 3 * i; idx_g[pos + 1] = BASE + 3 * i + 1;
                double zx = zf[2 * i], zy = zf[2 * i + 1];
                y_g[pos] = zx - feat_g[3 * i];
                y_g[pos + 1] = zy - feat_g[3 * i + 1];
                p.klt_last[((size_t)f * nmax + i) * 2] = zx;
                p.klt_last[((size_t)f * nmax + i) * 2 + 1] = zy;
                if (Rf[4 * i + 1] != Rf[4 * i + 2]) p.asym[f] = 1;   // sticky: Sigma turns asymmetric with this update
            } else if (i < n) {
                p.dflags[(size_t)f * nmax + i] = 1;
            }
            m += 2 * __popc(mask);
        }
        for (int a = m + lane; a < NB * 8; a += 32) s_idx[a] = 0;
        if (lane == 0) { s_m = m; s_bad = 0; p.m[f] = m; }
    }
    __syncthreads();
    const int m = s_m;
    if (m == 0) {
        if (tid == 0) {  // K is N x 0: only the quaternion renormalisation of :605-609 acts
            double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
            mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
            p.route[f] = (p.asym[f] || (p.flags & 0x400u)) ? ROUTE_JOSEPH_FULL : ROUTE_SYM;      // the covariance kernels only copy Sigma (m = 0)
        }
        return;
    }
    const int nb = (m + 7) >> 3;
    const bool sym_now = p.asym[f] == 0;
    if (p.sigma_lower && s_was_sym && !sym_now) {
        // the filter turns asymmetric with this update (asymmetric R): complete its Sigma, of which the last process()
        // wrote the feature rows only up to their diagonal blocks
        double* Pw = const_cast<double*>(Pi);
        const int N = BASE + 3 * n;
        mirror_lower_rows<NWC>(Pw, ld, N, warp, lane);
        __syncthreads();
    }
    const bool low = p.sigma_lower && sym_now;     // then Sigma(idx[b], idx[a]) is read through its mirror image

    // lower(a,b), a >= b  <-  upper(S)(b,a) = Sigma(idx[b], idx[a]) + R(b,a)  (SimplicialLDLT::compute(S')
    // reads upper(S), :578); identity tail.  Four independent gathers in flight per thread.
    {
        // 64 threads per tile, thread (tile of the pass, rr, cc); the tile pair (ib, jb) of the lower-triangular list is
        // advanced incrementally; four tiles (= four independent gathers) in flight per thread
        const int ntiles = nb * (nb + 1) / 2;
        const int rr = (tid >> 3) & 7, cc = tid & 7;
        constexpr int TPP = NTH / 64;               // tiles per pass of the CTA
        int ib = 0, jb = tid >> 6;
        while (jb > ib) { jb -= ib + 1; ++ib; }
        for (int t0 = tid >> 6; t0 < ntiles; t0 += 4 * TPP) {
            double v[4]; int o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + TPP * u;
                v[u] = 0.0; o[u] = -1;
                if (t < ntiles) {
                    const int a = ib * 8 + rr, b = jb * 8 + cc;
                    o[u] = t * 64 + tsw(rr, cc);
                    if (a >= b) {
                        if (a < m) {
                            v[u] = low ? Pi[(size_t)s_idx[a] * ld + s_idx[b]] : Pi[(size_t)s_idx[b] * ld + s_idx[a]];
                            if ((a >> 1) == (b >> 1)) v[u] += Rf[4 * ((s_idx[a] - BASE) / 3) + (b & 1) * 2 + (a & 1)];
                        } else {
                            v[u] = (a == b) ? 1.0 : 0.0;
                        }
                    }
                }
                jb += TPP;
                while (jb > ib) { jb -= ib + 1; ++ib; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (o[u] >= 0) Ls[o[u]] = v[u];
        }
    }
    __syncthreads();

    // S = L J L', J = diag(+-1): the reference's unpivoted LDL^T (SimplicialLDLT, TightlyCoupledEKF.cpp:577-580) carries on with an S
    // that is not positive definite, and so does this factorisation.
    __shared__ double s_sgn[NB * 8];
    __shared__ int s_neg, s_route;
    if (tid == 0) s_neg = 0;
    for (int a = tid; a < NB * 8; a += NTH) s_sgn[a] = 1.0;
    __syncthreads();
    chol_tiles<NWC>(Ls, Li, nb, &s_bad, s_sgn, &s_neg);
    // Routing of this filter's update (ekf_kernels.h ROUTE_*).  Sigma - Z Z' is only taken where it is as accurate as the Joseph
    // form: S positive definite with a pivot ratio below ILLCOND_RATIO.  Beyond that it would cancel the small posterior variances
    // against a huge prior, so such filters get the Joseph form term by term on the tiled kernels (no symmetry assumed), which
    // also handle the signed factor.  status bit3 records "S was not positive definite" (informational: the reference does not
    // notice), bit0 a zero pivot (the reference's Eigen::NumericalIssue, :579).
    if (warp == 0) {
        double dmax = 0.0, dmin = 1.79e308;
        for (int a = lane; a < m; a += 32) {
            const double d = Ls[tile_of(a >> 3, a >> 3) + tsw(a & 7, a & 7)];
            dmax = fmax(dmax, d); dmin = fmin(dmin, d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o)); dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o)); }
        if (lane == 0) {
            int route = ROUTE_SYM;
            if (p.flags & EKFVIO_FLAG_LITERAL_JOSEPH) route = ROUTE_JOSEPH_SYM;
            if (!(dmax * dmax <= p.illcond * dmin * dmin)) route = ROUTE_JOSEPH_SYM;          // (also catches NaN)
            if (s_neg) { route = ROUTE_JOSEPH_SYM; atomicOr(&p.status[f], 8); }
            if (!sym_now || (p.flags & 0x400u)) route = ROUTE_JOSEPH_FULL;                        // (0x400: diagnostic, never the symmetric kernel)
            if (s_bad) atomicOr(&p.status[f], 1);
            s_route = route; p.route[f] = route;
        }
    }
    __syncthreads();
    const int route = s_route;
    if (low && route != ROUTE_SYM) {
        // lower mode, and this filter leaves the symmetric fast path for this update: the kernels it goes to read all of Sigma,
        // of which the last process() wrote the feature rows only up to their diagonal blocks — complete it
        double* Pw = const_cast<double*>(Pi);
        const int N = BASE + 3 * n;
        mirror_lower_rows<NWC>(Pw, ld, N, warp, lane);
    }
    // Symmetric filters (ekf_fwd_tiled): v = inv(L) y replaces y, so that K y = Z v needs no K.  Block forward
    // substitution by one warp: t = y_jb - sum_k L(jb,k) v_k, v_jb = inv(L_jb,jb) t (explicit inverse tiles).
    if (warp == 0 && route == ROUTE_SYM) {
        __shared__ double s_vv[NB * 8], s_t[8];
        const int r = lane >> 2, q = lane & 3;
        for (int jb = 0; jb < nb; ++jb) {
            double acc = 0.0;
            for (int kb = 0; kb < jb; ++kb) {
                const double* T = Ls + tile_of(jb, kb);
                acc += T[tsw(r, 2 * q)] * s_vv[kb * 8 + 2 * q] + T[tsw(r, 2 * q + 1)] * s_vv[kb * 8 + 2 * q + 1];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            const int a = jb * 8 + r;
            if (q == 0) s_t[r] = (a < m ? y_g[a] : 0.0) - acc;
            __syncwarp();
            const double* I8 = Li + jb * 64;
            double vv = I8[tsw(r, 2 * q)] * s_t[2 * q] + I8[tsw(r, 2 * q + 1)] * s_t[2 * q + 1];
            vv += __shfl_xor_sync(0xffffffffu, vv, 1);
            vv += __shfl_xor_sync(0xffffffffu, vv, 2);
            if (q == 0) s_vv[a] = vv;
            __syncwarp();
        }
        for (int a = lane; a < m; a += 32) y_g[a] = s_vv[a];
    }
    // factor and inverse tiles to global scratch (same swizzled layout)
    double* Lg = p.L + (size_t)f * ((NT + NB) * 64 + NB * 8);
    const int used = nb * (nb + 1) / 2 * 64;
    for (int e = tid * 2; e < used; e += 2 * NTH) *reinterpret_cast<double2*>(Lg + e) = *reinterpret_cast<const double2*>(Ls + e);
    for (int e = tid * 2; e < nb * 64; e += 2 * NTH) *reinterpret_cast<double2*>(Lg + NT * 64 + e) = *reinterpret_cast<const double2*>(Li + e);
    for (int a = tid; a < NB * 8; a += NTH) Lg[(NT + NB) * 64 + a] = s_sgn[a];
}

// One CTA per filter — or, behind ekf_update_fused (p.fb != nullptr), a small grid that walks over the list of filters the fused
// kernel left alone (normally none: the launch is one read of the list length per CTA).
template <int NB, int NWC>
__global__ void __launch_bounds__(NWC * 32, 4) ekf_chol_tiled(EkfPtrs p, const double* __restrict__ Pin, const double* __restrict__ z,
                                                         const double* __restrict__ Rin, const uint8_t* __restrict__ pass) {
    if (p.fb == nullptr) { chol_tiled_filter<NB, NWC>(p, Pin, z, Rin, pass, blockIdx.x); return; }
    const int cnt = p.fb[0];
    for (int li = blockIdx.x; li < cnt; li += gridDim.x) {
        chol_tiled_filter<NB, NWC>(p, Pin, z, Rin, pass, p.fb[1 + li]);
        __syncthreads();
    }
}

// Kernel 2 of 2: K = Sigma(:,idx) inv(L)' inv(L), sparseView, mu += K y, W = Sigma(:,idx) - K S,
// quaternion renormalisation.  One warp per 16-row strip, the strip lives in registers through both
// triangular solves (right-looking over 8-wide column blocks).
template <int NW, int NB>
__device__ __forceinline__ void solve_tiled_filter(const EkfPtrs& p, const double* __restrict__ Pin, const double* __restrict__ Rin, int f) {
    extern __shared__ __align__(16) double smg[];
    constexpr int NT = NB * (NB + 1) / 2;
    double* Ls = smg;                      // NT tiles
    double* Li = Ls + NT * 64;             // NB inverse diagonal tiles (contiguous with Ls, as in scratch)
    double* Ss = Li + NB * 64;             // NB*NB tiles: the full (possibly asymmetric) S
    double* s_y = Ss + NB * NB * 64;       // NB*8
    int* s_idx = reinterpret_cast<int*>(s_y + NB * 8);   // NB*8
    __shared__ short s_inv[NW * 16];
    __shared__ double s_sgn[NB * 8];                     // J of S = L J L' (all +1 unless S is not positive definite)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m[f];
    if (m == 0) return;
#ifdef EKFVIO_PROFILE_CLOCKS
    long long t_prev = clock64();
#endif
    const int n = p.nfeat[f], N = BASE + 3 * n;
    const int ld = p.ldP, ldK = p.ldK, nmax = p.nmax;
    const double* Pi = Pin + (size_t)f * ld * ld;
    const double* Rf = Rin + (size_t)f * nmax * 4;
    double* Kf = p.K + (size_t)f * ld * ldK;
    double* Wf = p.W + (size_t)f * ld * ldK;
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * nmax * 3;
    const int nb = (m + 7) >> 3;
    const int r = lane >> 2, q = lane & 3;

    {   // L and inverse tiles: straight copy from scratch
        const double* Lg = p.L + (size_t)f * ((NT + NB) * 64 + NB * 8);
        const int used = nb * (nb + 1) / 2 * 64;
        for (int e = tid * 2; e < used; e += NW * 64) cp_async16(Ls + e, Lg + e);
        for (int e = tid * 2; e < nb * 64; e += NW * 64) cp_async16(Li + e, Lg + NT * 64 + e);
        cp_async_commit();
        const int* idx_g = p.idx + (size_t)f * p.mmax;
        const double* y_g = p.y + (size_t)f * p.mmax;
        for (int a = tid; a < NB * 8; a += NW * 32) { s_idx[a] = a < m ? idx_g[a] : 0; s_y[a] = a < m ? y_g[a] : 0.0; s_sgn[a] = Lg[(NT + NB) * 64 + a]; }
    }
    __syncthreads();
    // this warp's 16-row strip of Sigma(:,idx): the gathers are issued now, so that their latency
    // overlaps with the arrival of the L tiles
    const int i0 = warp * 16;
    double k0[2][NB], k1[2][NB];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int row = i0 + rt * 8 + r;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            int a = jb * 8 + 2 * q;
            bool ok = row < N && a < m && jb < nb;
            k0[rt][jb] = ok ? Pi[(size_t)row * ld + s_idx[a]] : 0.0;
            k1[rt][jb] = ok ? Pi[(size_t)row * ld + s_idx[a + 1]] : 0.0;
        }
    }
    // inverse measurement map: state row -> measurement index (or -1); identity tail rows of Ss
    for (int i = tid; i < ld; i += NW * 32) s_inv[i] = -1;
    __syncthreads();
    for (int a = tid; a < m; a += NW * 32) s_inv[s_idx[a]] = a;
    for (int e = tid; e < (nb * 8 - m) * nb * 8; e += NW * 32) {
        int a = m + e / (nb * 8), b = e % (nb * 8);
        Ss[((a >> 3) * NB + (b >> 3)) * 64 + tsw(a & 7, b & 7)] = (a == b) ? 1.0 : 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();

    CLK_MARK(4);
    if (i0 < N) {
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int row = i0 + rt * 8 + r;
            // a measured state row of Sigma(:,idx) is a row of S = Sigma(idx,idx) + R: the full S (no
            // symmetry assumed) is assembled from the strips instead of being gathered a second time
            const int am = row < N ? s_inv[row] : -1;
            if (am >= 0) {
                const double* rr = Rf + 4 * ((row - BASE) / 3) + (am & 1) * 2;
#pragma unroll
                for (int jb = 0; jb < NB; ++jb) {
                    if (jb < nb) {
                        int b = jb * 8 + 2 * q;
                        double v0 = k0[rt][jb], v1 = k1[rt][jb];
                        if (b == (am & ~1)) { v0 += rr[0]; v1 += rr[1]; }
                        *reinterpret_cast<double2*>(&Ss[((am >> 3) * NB + jb) * 64 + tsw(am & 7, 2 * q)]) = make_double2(v0, v1);
                    }
                }
            }
        }
        // forward: Z L' = C
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            if (jb < nb) {
                const double* I8 = Li + jb * 64;
                const double b0 = I8[tsw(r, q)], b1 = I8[tsw(r, 4 + q)];   // B[k][col] = inv(L)[col][k]
                double za[2][2];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    double a0, a1;
                    cfrag_to_afrag(k0[rt][jb], k1[rt][jb], lane, a0, a1);
                    double t0 = 0.0, t1 = 0.0;
                    dmma884(t0, t1, a0, b0);
                    dmma884(t0, t1, a1, b1);
                    k0[rt][jb] = t0; k1[rt][jb] = t1;
                    cfrag_to_afrag(t0, t1, lane, za[rt][0], za[rt][1]);
                }
#pragma unroll
                for (int j2 = 0; j2 < NB; ++j2) {
                    if (j2 > jb && j2 < nb) {
                        const double* T = Ls + tile_of(j2, jb);          // B[k][col] = L(j2,jb)[col][k]
                        const double l0 = T[tsw(r, q)], l1 = T[tsw(r, 4 + q)];
#pragma unroll
                        for (int rt = 0; rt < 2; ++rt) {
                            dmma884(k0[rt][j2], k1[rt][j2], -za[rt][0], l0);
                            dmma884(k0[rt][j2], k1[rt][j2], -za[rt][1], l1);
                        }
                    }
                }
            }
        }
        CLK_MARK(5);
        // K = Sigma(:,idx) inv(S) = (C inv(L)') J inv(L): the signs between the two substitutions
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            if (jb < nb) {
                const double j0 = s_sgn[jb * 8 + 2 * q], j1 = s_sgn[jb * 8 + 2 * q + 1];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) { k0[rt][jb] *= j0; k1[rt][jb] *= j1; }
            }
        }
        // backward: K L = Z J
#pragma unroll
        for (int jr = 0; jr < NB; ++jr) {
            const int jb = NB - 1 - jr;
            if (jb < nb) {
                const double* I8 = Li + jb * 64;
                const double b0 = I8[tsw(q, r)], b1 = I8[tsw(4 + q, r)];   // B[k][col] = inv(L)[k][col]
                double ka[2][2];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    double a0, a1;
                    cfrag_to_afrag(k0[rt][jb], k1[rt][jb], lane, a0, a1);
                    double t0 = 0.0, t1 = 0.0;
                    dmma884(t0, t1, a0, b0);
                    dmma884(t0, t1, a1, b1);
                    k0[rt][jb] = t0; k1[rt][jb] = t1;
                    cfrag_to_afrag(t0, t1, lane, ka[rt][0], ka[rt][1]);
                }
#pragma unroll
                for (int j2 = 0; j2 < NB; ++j2) {
                    if (j2 < jb) {
                        const double* T = Ls + tile_of(jb, j2);              // B[k][col] = L(jb,j2)[k][col]
                        const double l0 = T[tsw(q, r)], l1 = T[tsw(4 + q, r)];
#pragma unroll
                        for (int rt = 0; rt < 2; ++rt) {
                            dmma884(k0[rt][j2], k1[rt][j2], -ka[rt][0], l0);
                            dmma884(k0[rt][j2], k1[rt][j2], -ka[rt][1], l1);
                        }
                    }
                }
            }
        }
        // sparseView (:580), mu += K y (:600), K to global
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int row = i0 + rt * 8 + r;
            double dot = 0.0;
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                if (jb < nb) {
                    double a = prune(k0[rt][jb]), b = prune(k1[rt][jb]);
                    dot += a * s_y[jb * 8 + 2 * q] + b * s_y[jb * 8 + 2 * q + 1];
                    if (row < ld) *reinterpret_cast<double2*>(Kf + kw_at(ld, row, jb * 8 + 2 * q)) = make_double2(a, b);
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            if (q == 0 && row < N) { if (row < BASE) mu_g[row] += dot; else feat_g[row - BASE] += dot; }
        }
    }
    CLK_MARK(6);
    __syncthreads();   // every strip has contributed its rows of S, and K is in global memory
    if (i0 < N) {
        // W = Sigma(:,idx) - K S with the full S
        double w0[2][NB], w1[2][NB];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int row = i0 + rt * 8 + r;
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                int a = jb * 8 + 2 * q;
                bool ok = row < N && a < m && jb < nb;
                w0[rt][jb] = ok ? Pi[(size_t)row * ld + s_idx[a]] : 0.0;
                w1[rt][jb] = ok ? Pi[(size_t)row * ld + s_idx[a + 1]] : 0.0;
            }
        }
        const int rowa = min(i0 + r, ld - 1), rowb = min(i0 + 8 + r, ld - 1);
        const double* kra = Kf + (size_t)rowa * 16 + q;   // block kb: + (kb/2) * ld*16 + (kb%2) * 8
        const double* krb = Kf + (size_t)rowb * 16 + q;
        const size_t cst = (size_t)ld * 16;
        double ka[2][2] = {{ldg_pinned(kra), ldg_pinned(kra + 4)}, {ldg_pinned(krb), ldg_pinned(krb + 4)}};
        for (int kb = 0; kb < nb; ++kb) {
            double kn[2][2] = {{0, 0}, {0, 0}};
            if (kb + 1 < nb) {   // issued before this block's DMMAs (volatile asm keeps program order)
                const size_t o = (size_t)((kb + 1) >> 1) * cst + ((kb + 1) & 1) * 8;
                kn[0][0] = ldg_pinned(kra + o); kn[0][1] = ldg_pinned(kra + o + 4);
                kn[1][0] = ldg_pinned(krb + o); kn[1][1] = ldg_pinned(krb + o + 4);
            }
#pragma unroll
            for (int j2 = 0; j2 < NB; ++j2) {
                if (j2 < nb) {
                    const double* T = Ss + (kb * NB + j2) * 64;   // B[k][col] = S(kb*8+k, j2*8+col)
                    const double s0 = T[tsw(q, r)], s1 = T[tsw(4 + q, r)];
#pragma unroll
                    for (int rt = 0; rt < 2; ++rt) {
                        dmma884(w0[rt][j2], w1[rt][j2], -ka[rt][0], s0);
                        dmma884(w0[rt][j2], w1[rt][j2], -ka[rt][1], s1);
                    }
                }
            }
            ka[0][0] = kn[0][0]; ka[0][1] = kn[0][1]; ka[1][0] = kn[1][0]; ka[1][1] = kn[1][1];
        }
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int row = i0 + rt * 8 + r;
#pragma unroll
            for (int jb = 0; jb < NB; ++jb)
                if (jb < nb && row < ld) *reinterpret_cast<double2*>(Wf + kw_at(ld, row, jb * 8 + 2 * q)) = make_double2(w0[rt][jb], w1[rt][jb]);
        }
    }
    __syncthreads();
    CLK_MARK(7);
    if (tid == 0) {  // renormalise the quaternion (:605-609) and flag non-finite states
        double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
        mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
        bool fin = true;
        for (int i = 0; i < BASE; ++i) fin = fin && isfinite(mu_g[i]);
        if (!fin) atomicOr(&p.status[f], 2);
    }
}

template <int NW, int NB>
__global__ void __launch_bounds__(NW * 32, 1) ekf_solve_tiled(EkfPtrs p, const double* __restrict__ Pin, const double* __restrict__ Rin) {
    const int total = p.fb ? p.fb[0] : p.F;                // behind ekf_update_fused: only the filters it left alone
    for (int li = blockIdx.x; li < total; li += gridDim.x) {   // persistent: see ekf_joseph_tiled
        const int f = p.fb ? p.fb[1 + li] : li;
        if (p.route[f] == ROUTE_SYM || p.route[f] == ROUTE_DONE) continue;   // ekf_fwd_tiled / ekf_update_fused
        solve_tiled_filter<NW, NB>(p, Pin, Rin, f);
        __syncthreads();
    }
}

// Symmetric Sigma and R (the normal case): the Joseph form (I-KH) Sigma (I-KH)' + K R K' equals
// Sigma - Z Z' with Z = Sigma(:,idx) inv(L)', S = L L', and K y = Z (inv(L) y).  Only the forward
// substitution is needed; neither K nor W = Sigma(:,idx) - K S is formed.  Two CTAs per filter (six
// 16-row strips each, strips in registers as in ekf_solve_tiled) and two CTAs per SM, so that one CTA's
// factor copy and gathers overlap the other's substitution.  v = inv(L) y comes from ekf_chol_tiled (in p.y).
template <int NW, int NB>
__device__ __forceinline__ void fwd_tiled_filter(const EkfPtrs& p, const double* __restrict__ Pin, const int f) {
    extern __shared__ __align__(16) double smf[];
    constexpr int NT = NB * (NB + 1) / 2;
    double* Ls = smf;                      // NT tiles
    double* Li = Ls + NT * 64;             // NB inverse diagonal tiles
    double* s_v = Li + NB * 64;            // NB*8: inv(L) y
    int* s_idx = reinterpret_cast<int*>(s_v + NB * 8);   // NB*8

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m[f];
    if (m == 0 || p.route[f] != ROUTE_SYM) return;  // the others: ekf_solve_tiled
    const int n = p.nfeat[f], N = BASE + 3 * n;
    const int ld = p.ldP, ldK = p.ldK, nmax = p.nmax;
    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Zf = p.K + (size_t)f * ld * ldK;
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * nmax * 3;
    const int nb = (m + 7) >> 3;
    const int r = lane >> 2, q = lane & 3;
    {
        const double* Lg = p.L + (size_t)f * ((NT + NB) * 64 + NB * 8);
        const int used = nb * (nb + 1) / 2 * 64;
        for (int e = tid * 2; e < used; e += NW * 64) cp_async16(Ls + e, Lg + e);
        for (int e = tid * 2; e < nb * 64; e += NW * 64) cp_async16(Li + e, Lg + NT * 64 + e);
        cp_async_commit();
        const int* idx_g = p.idx + (size_t)f * p.mmax;
        const double* v_g = p.y + (size_t)f * p.mmax;
        for (int a = tid; a < NB * 8; a += NW * 32) { s_idx[a] = a < m ? idx_g[a] : 0; s_v[a] = a < m ? v_g[a] : 0.0; }
    }
    __syncthreads();
    const int i0 = (blockIdx.y * NW + warp) * 16;
    double k0[2][NB], k1[2][NB];
#pragma unroll
    for (int rt = 0; rt < 2; ++rt) {
        const int row = i0 + rt * 8 + r;
        // lower mode: a feature row is taken up to its diagonal only (what ekf_mirror_lower_kernel would copy above it is
        // exactly what is read through the mirror image here, so results do not depend on whether a mirror pass ran)
        const int ext = (!p.sigma_lower || row < BASE) ? ld : row + 1;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            int a = jb * 8 + 2 * q;
            bool ok = row < N && a < m && jb < nb;
            // lower mode: element (row, col) beyond the row's diagonal block is read as (col, row)
            const int c0 = s_idx[a], c1 = s_idx[a + 1];
            k0[rt][jb] = ok ? (c0 < ext ? Pi[(size_t)row * ld + c0] : Pi[(size_t)c0 * ld + row]) : 0.0;
            k1[rt][jb] = ok ? (c1 < ext ? Pi[(size_t)row * ld + c1] : Pi[(size_t)c1 * ld + row]) : 0.0;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (i0 < N) {
        // forward: Z L' = C
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            if (jb < nb) {
                const double* I8 = Li + jb * 64;
                const double b0 = I8[tsw(r, q)], b1 = I8[tsw(r, 4 + q)];   // B[k][col] = inv(L)[col][k]
                double za[2][2];
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    double a0, a1;
                    cfrag_to_afrag(k0[rt][jb], k1[rt][jb], lane, a0, a1);
                    double t0 = 0.0, t1 = 0.0;
                    dmma884(t0, t1, a0, b0);
                    dmma884(t0, t1, a1, b1);
                    k0[rt][jb] = t0; k1[rt][jb] = t1;
                    cfrag_to_afrag(t0, t1, lane, za[rt][0], za[rt][1]);
                }
#pragma unroll
                for (int j2 = 0; j2 < NB; ++j2) {
                    if (j2 > jb && j2 < nb) {
                        const double* T = Ls + tile_of(j2, jb);          // B[k][col] = L(j2,jb)[col][k]
                        const double l0 = T[tsw(r, q)], l1 = T[tsw(r, 4 + q)];
#pragma unroll
                        for (int rt = 0; rt < 2; ++rt) {
                            dmma884(k0[rt][j2], k1[rt][j2], -za[rt][0], l0);
                            dmma884(k0[rt][j2], k1[rt][j2], -za[rt][1], l1);
                        }
                    }
                }
            }
        }
        // mu += K y = Z v (:600); Z to the gain panel for the covariance kernel
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int row = i0 + rt * 8 + r;
            double dot = 0.0;
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) {
                if (jb < nb) {
                    const double a = k0[rt][jb], b = k1[rt][jb];
                    dot += a * s_v[jb * 8 + 2 * q] + b * s_v[jb * 8 + 2 * q + 1];
                    if (row < ld) *reinterpret_cast<double2*>(Zf + kw_at(ld, row, jb * 8 + 2 * q)) = make_double2(a, b);
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            if (q == 0 && row < N) { if (row < BASE) mu_g[row] += dot; else feat_g[row - BASE] += dot; }
        }
    }
    if (blockIdx.y == 0) {                 // the base rows (quaternion) all live in the first CTA's strips
        __syncthreads();
        if (tid == 0) {  // renormalise the quaternion (:605-609) and flag non-finite states
            double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
            mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
            bool fin = true;
            for (int i = 0; i < BASE; ++i) fin = fin && isfinite(mu_g[i]);
            if (!fin) atomicOr(&p.status[f], 2);
        }
    }
}

template <int NW, int NB>
__global__ void __launch_bounds__(NW * 32, 2) ekf_fwd_tiled(EkfPtrs p, const double* __restrict__ Pin) {
    if (p.fb == nullptr) { fwd_tiled_filter<NW, NB>(p, Pin, blockIdx.x); return; }
    const int cnt = p.fb[0];
    for (int li = blockIdx.x; li < cnt; li += gridDim.x) {
        fwd_tiled_filter<NW, NB>(p, Pin, p.fb[1 + li]);
        __syncthreads();
    }
}

template <int NW, int NB>
cudaError_t launch_gain_tiled_t(int which, const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, cudaStream_t st) {
    static bool configured_on[64] = {false};
    bool& configured = configured_on[current_device_slot()];
    constexpr int NT = NB * (NB + 1) / 2;
    const size_t sm_c = (size_t)(NT + NB) * 64 * sizeof(double) + NB * 8 * sizeof(int);
    const size_t sm_s = (size_t)((NT + NB) * 64 + NB * NB * 64 + NB * 8) * sizeof(double) + NB * 8 * sizeof(int);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_chol_tiled<NB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ekf_chol_tiled<NB, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ekf_solve_tiled<NW, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (which == 0) {
        static int chol_warps = 0;                 // EKFVIO_CHOL_WARPS=4 / 8 (experiments)
        if (!chol_warps) { const char* e = getenv("EKFVIO_CHOL_WARPS"); chol_warps = (e && atoi(e) == 8) ? 8 : 4; }
        // behind ekf_update_fused: the filters it left alone (0 ... 5 % of the batch per step on the config-3 streams,
        // tools/fallback_stats_probe.py).  The factorisation of one filter is a serial chain, so up to four CTAs share an SM (53 KB of
        // shared memory each) and the list is served in one wave; eight warps per filter
        const int grid = (p.fb && p.F > 4 * 148) ? 4 * 148 : p.F;
        if (p.fb) ekf_chol_tiled<NB, 8><<<grid, 256, sm_c, st>>>(p, Pin, z, R, pass);
        else if (chol_warps == 4) ekf_chol_tiled<NB, 4><<<grid, 128, sm_c, st>>>(p, Pin, z, R, pass);
        else ekf_chol_tiled<NB, 8><<<grid, 256, sm_c, st>>>(p, Pin, z, R, pass);
    }
    else {
        // symmetric filters: forward substitution only, two CTAs per filter; the others (and all of them under
        // EKFVIO_FLAG_LITERAL_JOSEPH): the full solve.  Each kernel skips the filters of the other.
        constexpr int NWF = (NW + 1) / 2;
        const size_t sm_f = (size_t)((NT + NB) * 64 + NB * 8) * sizeof(double) + NB * 8 * sizeof(int);
        static bool configured_f_on[64] = {false};
        bool& configured_f = configured_f_on[current_device_slot()];
        if (!configured_f) {
            cudaError_t e = cudaFuncSetAttribute(ekf_fwd_tiled<NWF, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f);
            if (e != cudaSuccess) return e;
            configured_f = true;
        }
        if (!(p.flags & EKFVIO_FLAG_LITERAL_JOSEPH)) ekf_fwd_tiled<NWF, NB><<<dim3((p.fb && p.F > 148) ? 148 : p.F, 2), NWF * 32, sm_f, st>>>(p, Pin);
        ekf_solve_tiled<NW, NB><<<p.F < 148 ? p.F : 148, NW * 32, sm_s, st>>>(p, Pin, R);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// ekf_joseph_sym: the covariance update for filters whose Sigma and R are symmetric (the normal
// case; p.asym[f] == 0).  Only the 16x16 blocks on or below the diagonal are computed —
//   Sigma'(I,J) = Sigma(I,J) - K(I,:) Sigma(idx,J) - W(I,:) K(J,:)'     for J <= I
// and written together with their mirror image, so Sigma' is exactly symmetric (the reference's
// Sigma' is symmetric up to rounding; its fixSigma() is a no-op, TightlyCoupledEKF.cpp:716-718).
// Both operands of every DMMA come from shared memory: per k-chunk the A panel (K or W, all rows)
// and the B panel (rows idx[k] of Sigma, or K) are staged with cp.async, three stages deep.  The
// lower-triangular block list is dealt round-robin to the warps, IPW blocks of 2x2 tiles each.
constexpr int SKC = 16;             // k-chunk depth
constexpr int SLDA = SKC + 4;       // A panel / K-as-B panel: [rows][SLDA]
constexpr int SNST = 3;

template <int NW, int NBLK, int IPW>
__global__ void __launch_bounds__(NW * 32, 1) ekf_joseph_sym(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout) {
    constexpr int ROWS = NBLK * 16;
    constexpr int SLDB = ROWS + 4;                       // G panel: [SKC][SLDB]
    constexpr int A_DOUBLES = ROWS * SLDA;
    constexpr int B_DOUBLES = (SKC * SLDB > A_DOUBLES) ? SKC * SLDB : A_DOUBLES;
    constexpr int STAGE = A_DOUBLES + B_DOUBLES;
    constexpr int TLD = 10;                              // per-warp 8x8 transpose tile, padded
    extern __shared__ __align__(16) double sms[];        // SNST stages | NW transpose tiles
    __shared__ int s_idx[2][2 * 104];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int ld = p.ldP, ldK = p.ldK;
    double* tw = sms + SNST * STAGE + warp * 8 * TLD;

    // Persistent CTA: filters f = blockIdx.x, blockIdx.x + gridDim.x, ...  The cp.async stage stream
    // runs on across the two phases of a filter and into the next filter, so the pipeline never
    // drains: while a filter's result is being stored, the first panels of the next one are in flight.
    // schur: the gain kernel left Z = Sigma(:,idx) inv(L)' in the K panel and Sigma' = Sigma - Z Z' is one phase;
    // otherwise (EKFVIO_FLAG_LITERAL_JOSEPH) the K and W panels go through the two phases of the Joseph form.
    // (per filter: ROUTE_SYM one phase with Z, ROUTE_JOSEPH_SYM two phases with K and W)
    // (behind ekf_update_fused the CTA walks over the list of filters that kernel left alone instead of the whole batch)
    const int total = p.fb ? p.fb[0] : p.F;
    struct Meta { int li, f, N, m, nch, nv; bool schur; const double* Pi; const double* Kf; const double* Wf; };
    auto meta = [&](int li) {
        const int f = li < total ? (p.fb ? p.fb[1 + li] : li) : p.F;
        Meta M; M.li = li; M.f = f; M.schur = true;
        const int route = f < p.F ? p.route[f] : ROUTE_JOSEPH_FULL;
        if (route != ROUTE_JOSEPH_FULL && route != ROUTE_DONE) {
            M.N = BASE + 3 * p.nfeat[f]; M.m = p.m[f]; M.schur = route == ROUTE_SYM;
        } else { M.N = 0; M.m = 0; }                        // nothing to do (asymmetric filters: ekf_joseph_tiled)
        const bool schur = M.schur;
        M.nch = (M.m + SKC - 1) / SKC; M.nv = schur ? M.nch : 2 * M.nch;
        M.Pi = Pin + (size_t)f * ld * ld; M.Kf = p.K + (size_t)f * ld * ldK; M.Wf = p.W + (size_t)f * ld * ldK;
        return M;
    };
    int gs = 0;                                            // stages issued so far -> buffer gs % SNST
    auto issue = [&](const Meta& M, int v, const int* sidx) {
        if (v < M.nv) {
            const bool schur = M.schur;
            double* A = sms + (gs % SNST) * STAGE;
            double* B = A + A_DOUBLES;
            const int phase = schur ? 1 : (v >= M.nch), k0 = ((phase && !schur) ? v - M.nch : v) * SKC, m = M.m;
            const double* Am = (phase && !schur) ? M.Wf : M.Kf;
            for (int t = tid; t < ROWS * (SKC / 2); t += NW * 32) {   // A panel: rows of K or W
                int row = t / (SKC / 2), seg = t % (SKC / 2);
                double* dst = &A[row * SLDA + seg * 2];
                if (row < ld && k0 + seg * 2 < m) cp_async16(dst, Am + kw_at(ld, row, k0 + seg * 2));
                else { dst[0] = 0.0; dst[1] = 0.0; }
            }
            if (!phase) {                                             // B panel: rows idx[k] of Sigma
                for (int t = tid; t < SKC * (ROWS / 2); t += NW * 32) {
                    int k = t / (ROWS / 2), seg = t % (ROWS / 2);
                    double* dst = &B[k * SLDB + seg * 2];
                    if (k0 + k < m && seg * 2 < ld) cp_async16(dst, M.Pi + (size_t)sidx[k0 + k] * ld + seg * 2);
                    else { dst[0] = 0.0; dst[1] = 0.0; }
                }
            } else if (!schur) {                                      // B panel: rows of K (Z Z': the A panel serves as both operands)
                for (int t = tid; t < ROWS * (SKC / 2); t += NW * 32) {
                    int row = t / (SKC / 2), seg = t % (SKC / 2);
                    double* dst = &B[row * SLDA + seg * 2];
                    if (row < ld && k0 + seg * 2 < m) cp_async16(dst, M.Kf + kw_at(ld, row, k0 + seg * 2));
                    else { dst[0] = 0.0; dst[1] = 0.0; }
                }
            }
        }
        cp_async_commit();   // exactly one group per call (possibly empty): keeps the wait arithmetic uniform
        ++gs;
    };

    Meta cur = meta(blockIdx.x);
    int ib = 0;
    for (int i = tid; i < cur.m; i += NW * 32) s_idx[0][i] = p.idx[(size_t)cur.f * p.mmax + i];
    __syncthreads();
    issue(cur, 0, s_idx[0]); issue(cur, 1, s_idx[0]);
    // Ring position of the current / next filter's stage 0.  Stage v of a filter is group g0 + v: a filter with fewer than two
    // stages (nothing to do here, or m = 0) still advances the ring by the two (empty) groups issued for it, so the consumer must
    // follow the issue counter, not a count of stages consumed.
    int g0_cur = 0, g0_nxt = 0;

    while (cur.li < total) {
        Meta nxt = meta(cur.li + gridDim.x);
        for (int i = tid; i < nxt.m; i += NW * 32) s_idx[ib ^ 1][i] = p.idx[(size_t)nxt.f * p.mmax + i];
        __syncthreads();
        const int N = cur.N;
        const int nblk = (N + 15) >> 4, nitems = nblk * (nblk + 1) / 2;
        int bi[IPW], bj[IPW], aoff[IPW], boffg[IPW], boffk[IPW];
        double c0[IPW][4], c1[IPW][4];                    // tile t = rt*2 + ct of the 16x16 block
#pragma unroll
        for (int it = 0; it < IPW; ++it) {
            int e = warp + it * NW;
            bool ok = e < nitems;
            int i = 0, j = 0;
            if (ok) {
                i = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
                while (i * (i + 1) / 2 > e) --i;
                while ((i + 1) * (i + 2) / 2 <= e) ++i;
                j = e - i * (i + 1) / 2;
            }
            bi[it] = ok ? i : -1; bj[it] = j;
            aoff[it] = (i * 16 + r) * SLDA + q;
            boffg[it] = q * SLDB + j * 16 + r;
            boffk[it] = (j * 16 + r) * SLDA + q;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                double2 v = make_double2(0.0, 0.0);
                int row = i * 16 + (t >> 1) * 8 + r, col = j * 16 + (t & 1) * 8 + 2 * q;
                if (ok && row < ld && col < ld) v = *reinterpret_cast<const double2*>(cur.Pi + (size_t)row * ld + col);
                c0[it][t] = v.x; c1[it][t] = v.y;
            }
        }
        int issued_next = 0;
        for (int v = 0; v < cur.nv; ++v) {
            cp_async_wait<SNST - 2>();
            __syncthreads();
            if (v + 2 < cur.nv) issue(cur, v + 2, s_idx[ib]);
            else { if (issued_next == 0) g0_nxt = gs; issue(nxt, issued_next, s_idx[ib ^ 1]); ++issued_next; }
            const double* A = sms + ((g0_cur + v) % SNST) * STAGE;
            const double* B = cur.schur ? A : A + A_DOUBLES;
            const bool phase1 = cur.schur || v >= cur.nch;
#pragma unroll
            for (int kk = 0; kk < SKC / 4; ++kk) {
                double fa0[IPW], fa1[IPW], fb0[IPW], fb1[IPW];
#pragma unroll
                for (int it = 0; it < IPW; ++it) {   // idle slots (bi < 0) recompute block (0,0) and are never stored
                    const double* Ai = A + aoff[it] + kk * 4;
                    fa0[it] = Ai[0]; fa1[it] = Ai[8 * SLDA];
                    if (!phase1) { const double* Bi = B + boffg[it] + kk * 4 * SLDB; fb0[it] = Bi[0]; fb1[it] = Bi[8]; }
                    else { const double* Bi = B + boffk[it] + kk * 4; fb0[it] = Bi[0]; fb1[it] = Bi[8 * SLDA]; }
                }
#pragma unroll
                for (int it = 0; it < IPW; ++it) {
                    dmma884(c0[it][0], c1[it][0], -fa0[it], fb0[it]);
                    dmma884(c0[it][1], c1[it][1], -fa0[it], fb1[it]);
                    dmma884(c0[it][2], c1[it][2], -fa1[it], fb0[it]);
                    dmma884(c0[it][3], c1[it][3], -fa1[it], fb1[it]);
                }
            }
        }
        if (issued_next < 2 && cur.nv > 0) __syncthreads();   // a one-stage filter: the refill below reuses the buffer just consumed
        while (issued_next < 2) { if (issued_next == 0) g0_nxt = gs; issue(nxt, issued_next, s_idx[ib ^ 1]); ++issued_next; }   // short filters (nv < 2)

        // epilogue: Sigma'(I,J) and its mirror Sigma'(J,I); the mirror goes through an 8x8 transpose in
        // shared memory so that both stores are row-contiguous
        if (N > 0) {
            double* Po = Pout + (size_t)cur.f * ld * ld;
#pragma unroll
            for (int it = 0; it < IPW; ++it) {
                if (bi[it] < 0) continue;
                const bool diag = bi[it] == bj[it];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int rt = t >> 1, ct = t & 1;
                    const int row = bi[it] * 16 + rt * 8 + r, col = bj[it] * 16 + ct * 8 + 2 * q;
                    const double v0 = prune(c0[it][t]), v1 = prune(c1[it][t]);
                    if (diag && ct > rt) continue;                       // upper tile of a diagonal block: mirrored from (ct,rt)
                    const bool dtile = diag && ct == rt;                 // tile on the diagonal: keep col <= row, mirror the rest
                    if (row < N) {
                        if (!dtile) {
                            if (col + 1 < N) *reinterpret_cast<double2*>(Po + (size_t)row * ld + col) = make_double2(v0, v1);
                            else if (col < N) Po[(size_t)row * ld + col] = v0;
                        } else {
                            if (col <= row) Po[(size_t)row * ld + col] = v0;
                            if (col + 1 <= row) Po[(size_t)row * ld + col + 1] = v1;
                        }
                    }
                    __syncwarp();
                    tw[r * TLD + 2 * q] = v0; tw[r * TLD + 2 * q + 1] = v1;
                    __syncwarp();
                    // lane (r, q) now writes row' = tile column r, columns' = tile rows 2q, 2q+1
                    const double m0 = tw[(2 * q) * TLD + r], m1 = tw[(2 * q + 1) * TLD + r];
                    const int mrow = bj[it] * 16 + ct * 8 + r, mcol = bi[it] * 16 + rt * 8 + 2 * q;
                    if (mrow < N) {
                        if (!dtile) {
                            if (mcol + 1 < N) *reinterpret_cast<double2*>(Po + (size_t)mrow * ld + mcol) = make_double2(m0, m1);
                            else if (mcol < N) Po[(size_t)mrow * ld + mcol] = m0;
                        } else {
                            if (mcol > mrow && mcol < N) Po[(size_t)mrow * ld + mcol] = m0;
                            if (mcol + 1 > mrow && mcol + 1 < N) Po[(size_t)mrow * ld + mcol + 1] = m1;
                        }
                    }
                }
            }
        }
        cur = nxt; ib ^= 1; g0_cur = g0_nxt;
    }
    cp_async_wait<0>();
}

template <int NW, int NBLK, int IPW>
cudaError_t launch_joseph_sym_t(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st) {
    constexpr int ROWS = NBLK * 16;
    constexpr int A_DOUBLES = ROWS * SLDA;
    constexpr int B_DOUBLES = (SKC * (ROWS + 4) > A_DOUBLES) ? SKC * (ROWS + 4) : A_DOUBLES;
    const size_t sm = (size_t)(SNST * (A_DOUBLES + B_DOUBLES) + NW * 8 * 10) * sizeof(double);
    static int sms_count_on[64] = {0};
    int& sms_count = sms_count_on[current_device_slot()];
    if (!sms_count) {
        cudaError_t e = cudaFuncSetAttribute(ekf_joseph_sym<NW, NBLK, IPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const int grid = p.F < sms_count ? p.F : sms_count;    // persistent: one CTA per SM
    ekf_joseph_sym<NW, NBLK, IPW><<<grid, NW * 32, sm, st>>>(p, Pin, Pout);
    return cudaGetLastError();
}

}  // namespace

namespace ekfvio {

bool joseph_tiled_supported(const EkfPtrs& p) { return p.Nmax <= 256 && p.mmax <= 256; }
bool joseph_sym_supported(const EkfPtrs& p) { return p.Nmax <= 176 && p.mmax <= 208; }

// Symmetric filters only (p.asym[f] == 0); the others are left to launch_joseph_tiled.
cudaError_t launch_joseph_sym(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st) {
    if (p.Nmax <= 128) return launch_joseph_sym_t<9, 8, 4>(p, Pin, Pout, st);      // 36 blocks
    return launch_joseph_sym_t<11, 11, 6>(p, Pin, Pout, st);                       // 66 blocks, 6 per warp
}

bool gain_tiled_supported(const EkfPtrs& p) { return p.Nmax <= 176 && p.mmax <= 104 && p.L != nullptr; }
size_t gain_tiled_scratch_doubles(int mmax) { int NB = mmax <= 64 ? 8 : 13; return (size_t)(NB * (NB + 1) / 2 + NB) * 64 + NB * 8; }   // L tiles | inverse diagonal tiles | signs

// which = 0: Cholesky kernel, 1: solve kernel (two launches so the API layer can time them apart)
cudaError_t launch_gain_tiled(int which, const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, cudaStream_t st) {
    if (p.Nmax <= 128 && p.mmax <= 64) return launch_gain_tiled_t<8, 8>(which, p, Pin, z, R, pass, st);
    return launch_gain_tiled_t<11, 13>(which, p, Pin, z, R, pass, st);
}

cudaError_t launch_joseph_tiled(const EkfPtrs& p, const double* Pin, double* Pout, int only_asym, cudaStream_t st) {
    const int strips = (p.Nmax + 15) / 16;
    const size_t sm = (size_t)NST * BS_DOUBLES * sizeof(double);
    static bool configured_on[64] = {false};
    bool& configured = configured_on[current_device_slot()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_joseph_tiled<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ekf_joseph_tiled<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ekf_joseph_tiled<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int grid = p.F < 148 ? p.F : 148;                 // persistent CTAs (one per SM)
    if (strips <= 8) ekf_joseph_tiled<8><<<grid, 8 * 32, sm, st>>>(p, Pin, Pout, only_asym);
    else if (strips <= 11) ekf_joseph_tiled<11><<<grid, 11 * 32, sm, st>>>(p, Pin, Pout, only_asym);
    else ekf_joseph_tiled<16><<<grid, 16 * 32, sm, st>>>(p, Pin, Pout, only_asym);
    return cudaGetLastError();
}

}  // namespace ekfvio

#ifdef EKFVIO_PROFILE_CLOCKS
extern "C" void ekfvio_debug_clocks(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_clk, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_clk, z, sizeof(z)); }
}
#endif

