// General-path EKF kernels (any feature count, any — possibly asymmetric — Sigma and R).
// One CTA per filter, FP64 CUDA-core arithmetic, all panels in global/L2.  These kernels define the
// batched semantics and back the tiled fast path (ekf_tiled.cu) whenever its preconditions do
// not hold.  Reference: include/ekf_vio/TightlyCoupledEKF.cpp (line numbers per function below).
#include "ekf_common.cuh"
#include "ekf_kernels.h"

using namespace ekfvio;

namespace {

constexpr int PT = 256;  // threads per CTA for the general kernels

// ---------------------------------------------------------------------------------------------
// Finite-difference Jacobian pieces + state propagation, shared by process and linearize.
// smem layout (doubles): mu[22] | fdb[32][22] | A[22][23] | dq[8][4] | bvec[19][3] (+pad) | feat[3n] | B[3n][9] | D[n][9]
struct ProcSmem {
    double* mu; double* fdb; double* A; double* dq; double* bvec; double* feat; double* B; double* D;
    __device__ ProcSmem(double* base, int n) {
        mu = base; fdb = mu + 22; A = fdb + 32 * 22; dq = A + 22 * 23; bvec = dq + 32; feat = bvec + 64; B = feat + 3 * n; D = B + 27 * n;
    }
    static __host__ __device__ size_t doubles(int n) { return 22 + 32 * 22 + 22 * 23 + 32 + 64 + 3 * (size_t)n + 27 * (size_t)n + 9 * (size_t)n; }
};

// numericallyLinearizeProcess (TightlyCoupledEKF.cpp:176-325) as independent evaluations.
// Fills s.A (22x22, ld 23), s.B (3n x 9: d feature / d base cols 7..15), s.D (n x 3 x 3) and
// leaves the dq_inv the reference's convolveFeature cache ends with in s.dq[1] (fresh, base omega).
// x / z given r = RN(1 / z): one multiply and two fused corrections give the correctly rounded
// quotient (Markstein), i.e. the bits a division would — a third of the instructions.
__device__ __forceinline__ double div_with_rcp(double x, double z, double r) {
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-q0, z, x);
    return __fma_rn(rem, r, q0);
}

__device__ void linearize_block(const ProcSmem& s, int n, double dt, const double* cache7, bool fresh_cache) {
    const int tid = threadIdx.x;
    const double two_d = 2 * DELTA_SHIFT;
    // base-state evaluations: column j = t/2, t%2 == 0 -> x+d, 1 -> (x+d)-2d
    if (tid < 32) {
        double tm[22];
#pragma unroll
        for (int i = 0; i < 22; ++i) tm[i] = s.mu[i];
        int j = tid >> 1;
        double v = tm[j] + DELTA_SHIFT;
        if (tid & 1) v = v - two_d;
        tm[j] = v;
        double o[22];
        convolve_base(tm, dt, o);
#pragma unroll
        for (int i = 0; i < 22; ++i) s.fdb[tid * 22 + i] = o[i];
    }
    // the 8 dq_inv variants convolveFeature would use:
    //   0: columns 7..9  (cached value if the cache key matches the base omega — E2 — else fresh)
    //   1: fresh for the base omega (columns 13..15, D blocks, state propagation)
    //   2+2c+s: omega_c perturbed (+ / -) for columns 10..12
    if (tid >= 32 && tid < 40) {
        int q = tid - 32;
        double ox = s.mu[10], oy = s.mu[11], oz = s.mu[12];
        Q4 r;
        if (q == 0 && !fresh_cache && cache7[0] == ox && cache7[1] == oy && cache7[2] == oz) {
            r = {cache7[3], cache7[4], cache7[5], cache7[6]};
        } else {
            if (q >= 2) {
                int c = (q - 2) >> 1;
                double* o = (c == 0) ? &ox : (c == 1) ? &oy : &oz;
                double v = *o + DELTA_SHIFT;
                if (q & 1) v = v - two_d;
                *o = v;
            }
            r = delta_quat(ox, oy, oz, dt, -1.0);
        }
        s.dq[q * 4 + 0] = r.w; s.dq[q * 4 + 1] = r.x; s.dq[q * 4 + 2] = r.y; s.dq[q * 4 + 3] = r.z;
    }
    __syncthreads();
    // A = d base / d base
    const double r2d = 1.0 / two_d;
    for (int e = tid; e < 22 * 22; e += blockDim.x) {
        int i = e / 22, j = e % 22;
        double v;
        if (j < 16) v = div_with_rcp(s.fdb[(2 * j) * 22 + i] - s.fdb[(2 * j + 1) * 22 + i], two_d, r2d);
        else v = (i == j) ? 1.0 : 0.0;
        s.A[i * 23 + j] = v;
    }
    // convolveFeature evaluates  (R fp - R tr) / z  with fp from the feature and tr, R from the base
    // state.  R tr does not depend on the feature: the 19 variants (columns 7..15 x {+,-}, and the
    // unperturbed one) are computed once per filter.
    const double hdt2 = 0.5 * dt * dt;
    if (tid >= 64 && tid < 64 + 19) {
        const int t = tid - 64;
        V3 vel{s.mu[7], s.mu[8], s.mu[9]}, acc{s.mu[13], s.mu[14], s.mu[15]};
        int qi = 1;
        if (t < 18) {
            const int j = 7 + (t >> 1), sgn = t & 1;
            double val = s.mu[j] + DELTA_SHIFT;
            if (sgn) val = val - two_d;
            if (j <= 9) { (j == 7 ? vel.x : j == 8 ? vel.y : vel.z) = val; qi = 0; }
            else if (j <= 12) { qi = 2 + 2 * (j - 10) + sgn; }
            else { (j == 13 ? acc.x : j == 14 ? acc.y : acc.z) = val; }
        }
        const Q4 dq{s.dq[qi * 4], s.dq[qi * 4 + 1], s.dq[qi * 4 + 2], s.dq[qi * 4 + 3]};
        const V3 tr{dt * vel.x + hdt2 * acc.x, dt * vel.y + hdt2 * acc.y, dt * vel.z + hdt2 * acc.z};
        const V3 bv = qrot(dq, tr);
        s.bvec[t * 3] = bv.x; s.bvec[t * 3 + 1] = bv.y; s.bvec[t * 3 + 2] = bv.z;
    }
    __syncthreads();
    // feature evaluations, item (feature f, group g): g = 0 -> columns 7..9 (cached dq_inv, E2),
    // g = 1 -> columns 13..15, g = 2 -> columns 10..12 (one dq_inv per evaluation), g = 3 -> own u, v, rho
    auto project = [&](const V3& a, const double* bv, double* o) {   // (x/z, y/z, 1/z) of a - b
        const double x = a.x - bv[0], y = a.y - bv[1], z = a.z - bv[2];
        const double iz = 1.0 / z;   // quotients stay correctly rounded: one-ulp differences here grow ~5000x through the gain
        o[0] = div_with_rcp(x, z, iz); o[1] = div_with_rcp(y, z, iz); o[2] = iz;
    };
    for (int e = tid; e < n * 4; e += blockDim.x) {
        const int f = e >> 2, g = e & 3;
        const double u = s.feat[3 * f], v = s.feat[3 * f + 1], rho = s.feat[3 * f + 2];
        double hi[3], lo[3];
        if (g < 3) {
            V3 fp;
            fp.z = 1.0 / rho; fp.x = u * fp.z; fp.y = v * fp.z;
            if (g < 2) {
                const Q4 dq{s.dq[g * 4], s.dq[g * 4 + 1], s.dq[g * 4 + 2], s.dq[g * 4 + 3]};
                const V3 a = qrot(dq, fp);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = (g == 0 ? 0 : 6) + k;
                    project(a, s.bvec + (2 * c) * 3, hi);
                    project(a, s.bvec + (2 * c + 1) * 3, lo);
#pragma unroll
                    for (int r = 0; r < 3; ++r) s.B[(3 * f + r) * 9 + c] = div_with_rcp(hi[r] - lo[r], two_d, r2d);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = 3 + k;
#pragma unroll
                    for (int sgn = 0; sgn < 2; ++sgn) {
                        const int qi = 2 + 2 * k + sgn;
                        const Q4 dq{s.dq[qi * 4], s.dq[qi * 4 + 1], s.dq[qi * 4 + 2], s.dq[qi * 4 + 3]};
                        project(qrot(dq, fp), s.bvec + (2 * c + sgn) * 3, sgn == 0 ? hi : lo);
                    }
#pragma unroll
                    for (int r = 0; r < 3; ++r) s.B[(3 * f + r) * 9 + c] = div_with_rcp(hi[r] - lo[r], two_d, r2d);
                }
            }
        } else {
            const Q4 dq{s.dq[4], s.dq[5], s.dq[6], s.dq[7]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double t[3] = {u, v, rho};
                const double up = t[k] + DELTA_SHIFT;
#pragma unroll
                for (int sgn = 0; sgn < 2; ++sgn) {
                    t[k] = sgn == 0 ? up : up - two_d;
                    V3 fp;
                    fp.z = 1.0 / t[2]; fp.x = t[0] * fp.z; fp.y = t[1] * fp.z;
                    project(qrot(dq, fp), s.bvec + 18 * 3, sgn == 0 ? hi : lo);
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) s.D[f * 9 + r * 3 + k] = div_with_rcp(hi[r] - lo[r], two_d, r2d);
            }
        }
    }
    __syncthreads();
}

// First half of process(dt) for batches that fill the GPU: linearisation (A, B, D into the idle W
// panel), dq cache refresh and state propagation, in small CTAs so that the serial finite-difference
// chains of many filters overlap.  ekf_process_general(pre = 1) then does the covariance pass.
__global__ void __launch_bounds__(128) ekf_linearize_kernel(EkfPtrs p, const double* __restrict__ dts) {
    extern __shared__ double sm[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = p.nfeat[f];
    ProcSmem s(sm, n);
    const double dt = dts[f];
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * p.nmax * 3;
    double* cache_g = p.cache + (size_t)f * 7;
    for (int i = tid; i < BASE; i += blockDim.x) s.mu[i] = mu_g[i];
    for (int i = tid; i < 3 * n; i += blockDim.x) s.feat[i] = feat_g[i];
    __syncthreads();
    linearize_block(s, n, dt, cache_g, (p.flags & EKFVIO_FLAG_FRESH_DQ_CACHE) != 0);
    if (n > 0 && tid == 0) {
        cache_g[0] = s.mu[10]; cache_g[1] = s.mu[11]; cache_g[2] = s.mu[12];
        cache_g[3] = s.dq[4]; cache_g[4] = s.dq[5]; cache_g[5] = s.dq[6]; cache_g[6] = s.dq[7];
    }
    double* lin = p.W + (size_t)f * p.ldP * p.ldK;
    for (int i = tid; i < 22 * 23; i += blockDim.x) lin[i] = s.A[i];
    for (int i = tid; i < 27 * n; i += blockDim.x) lin[22 * 23 + i] = s.B[i];
    for (int i = tid; i < 9 * n; i += blockDim.x) lin[22 * 23 + 27 * p.nmax + i] = s.D[i];
    for (int fi = tid; fi < n; fi += blockDim.x) {
        V3 vel{s.mu[7], s.mu[8], s.mu[9]}, acc{s.mu[13], s.mu[14], s.mu[15]};
        Q4 dq{s.dq[4], s.dq[5], s.dq[6], s.dq[7]};
        double o[3];
        convolve_feature(dq, vel, acc, dt, s.feat[3 * fi], s.feat[3 * fi + 1], s.feat[3 * fi + 2], o);
        feat_g[3 * fi] = o[0]; feat_g[3 * fi + 1] = o[1]; feat_g[3 * fi + 2] = o[2];
    }
    if (tid == 0) {
        double o[22];
        convolve_base(s.mu, dt, o);
        for (int i = 0; i < 22; ++i) mu_g[i] = o[i];
    }
}

// process(dt) — TightlyCoupledEKF.cpp:96-121.  mode 0: full step (state + covariance).
// mode 1: linearize only (dense F written to F_out, state untouched except the dq cache).
__global__ void __launch_bounds__(PT, 2) ekf_process_general(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout,
                                                          const double* __restrict__ dts, int mode, double* __restrict__ F_out, int fused,
                                                          int pre, int lower) {
    extern __shared__ double sm[];
    const int f = blockIdx.x, tid = threadIdx.x;
    if (lower == 2 && p.asym[f] == 0) return;   // symmetric filters were served by ekf_process_cov_tiles (ekf_process_tiles.cu)
    const int n = p.nfeat[f], N = BASE + 3 * n;
    const int ld = p.ldP;
    ProcSmem s(sm, n);
    const double dt = dts[f];
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * p.nmax * 3;
    double* cache_g = p.cache + (size_t)f * 7;
    if (pre) {
        // A, B, D were left in the (idle) W panel by ekf_linearize_kernel, which also moved the state
        const double* lin = p.W + (size_t)f * ld * p.ldK;
        for (int i = tid; i < 22 * 23; i += blockDim.x) s.A[i] = lin[i];
        for (int i = tid; i < 27 * n; i += blockDim.x) s.B[i] = lin[22 * 23 + i];
        for (int i = tid; i < 9 * n; i += blockDim.x) s.D[i] = lin[22 * 23 + 27 * p.nmax + i];
        __syncthreads();
    } else {
    for (int i = tid; i < BASE; i += blockDim.x) s.mu[i] = mu_g[i];
    for (int i = tid; i < 3 * n; i += blockDim.x) s.feat[i] = feat_g[i];
    __syncthreads();
    linearize_block(s, n, dt, cache_g, (p.flags & EKFVIO_FLAG_FRESH_DQ_CACHE) != 0);
    }

    // Row-split launch (gridDim.y > 1, large states): every CTA of a filter reads the old state and
    // dq cache, so the new ones are staged in the (idle) gain panel and committed by
    // ekf_commit_state_kernel afterwards.
    const int ysplit = gridDim.y, yb = blockIdx.y;
    if (ysplit > 1) {
        double* stage = p.K + (size_t)f * ld * p.ldK;
        mu_g = stage; feat_g = stage + BASE; cache_g = stage + BASE + 3 * p.nmax;
    }
    if (n > 0 && tid == 0 && yb == 0 && !pre) {  // cache ends fresh for the base omega (see linearize_block)
        cache_g[0] = s.mu[10]; cache_g[1] = s.mu[11]; cache_g[2] = s.mu[12];
        cache_g[3] = s.dq[4]; cache_g[4] = s.dq[5]; cache_g[5] = s.dq[6]; cache_g[6] = s.dq[7];
    }
    if (mode == 1) {
        double* Fo = F_out + (size_t)f * p.Nmax * p.Nmax;
        const int LN = p.Nmax;
        for (int e = tid; e < N * N; e += blockDim.x) {
            int i = e / N, j = e % N;
            double v = 0.0;
            if (i < BASE) { if (j < BASE) v = s.A[i * 23 + j]; }
            else {
                int fi = (i - BASE) / 3, r = (i - BASE) % 3;
                if (j >= 7 && j <= 15) v = s.B[(3 * fi + r) * 9 + (j - 7)];
                else if (j >= BASE && (j - BASE) / 3 == fi) v = s.D[fi * 9 + r * 3 + (j - BASE) % 3];
            }
            Fo[(size_t)i * LN + j] = v;
        }
        return;
    }

    // state propagation: features with the OLD base state (:102-104), then the base state (:107)
    if (yb == 0 && !pre)
    for (int fi = tid; fi < n; fi += blockDim.x) {
        V3 vel{s.mu[7], s.mu[8], s.mu[9]}, acc{s.mu[13], s.mu[14], s.mu[15]};
        Q4 dq{s.dq[4], s.dq[5], s.dq[6], s.dq[7]};
        double o[3];
        convolve_feature(dq, vel, acc, dt, s.feat[3 * fi], s.feat[3 * fi + 1], s.feat[3 * fi + 2], o);
        feat_g[3 * fi] = o[0]; feat_g[3 * fi + 1] = o[1]; feat_g[3 * fi + 2] = o[2];
    }
    if (tid == 0 && yb == 0 && !pre) {
        double o[22];
        convolve_base(s.mu, dt, o);
        for (int i = 0; i < 22; ++i) mu_g[i] = o[i];
    }

    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Po = Pout + (size_t)f * ld * ld;
    if (p.flags & 0x800u) return;   // debug bit: linearisation + state only (timing experiments)
    if (fused) {
        // Sigma' = F Sigma F' + Q fused by row blocks: a warp takes three rows I of F, forms
        // T = F(I,:) Sigma (3 x N) in its slice of shared memory, then Sigma'(I,:) = T F'.
        // Sigma is read once and written once.  Rows 7..15 of Sigma (needed by every feature block)
        // are staged once per CTA; a warp's own three rows arrive by cp.async one task ahead.
        const int Np = (N + 3) & ~3;
        const int Npm = ((BASE + 3 * p.nmax) + 3) & ~3;
        double* Pb = sm + ((ProcSmem::doubles(p.nmax) + 1) & ~(size_t)1);   // [9][Np]: rows 7..15 (16-byte aligned for cp.async)
        const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
        double* Tw = Pb + 9 * (size_t)Npm + (size_t)warp * 6 * Npm;   // [3][Np]
        double* Own = Tw + 3 * (size_t)Npm;                           // [3][Np]
        for (int e = tid; e < 9 * N; e += blockDim.x) { int k = e / N, c = e - k * N; Pb[k * Np + c] = Pi[(size_t)(7 + k) * ld + c]; }
        const int ntask = 8 + n;
        // Lower mode (symmetric filter whose next consumer is the reduced update): a feature row block is only
        // needed up to its own diagonal block — columns c < BASE + 3 fi + 3 of T and of Sigma' — which removes
        // almost half of both passes and of the traffic; rows 0..21 stay complete.  The upper part of the feature
        // rows of the output is then stale (ekf_api.cu mirrors it on demand).
        const bool low = lower == 1 && p.asym[f] == 0;
        auto prefetch = [&](int t) {
            if (t >= 8 && t < ntask) {
                const double* own = Pi + (size_t)(BASE + 3 * (t - 8)) * ld;
                const int cn = low ? BASE + 3 * (t - 8) + 3 : N, half = ((cn + 3) & ~3) / 2;
                for (int e = lane; e < 3 * half; e += 32) {
                    int r = e / half, c2 = (e - r * half) * 2;
                    unsigned dst = (unsigned)__cvta_generic_to_shared(Own + r * Np + c2);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(own + (size_t)r * ld + c2) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        prefetch(warp + nw);            // tasks 0..7 are the base-row tasks, one per warp in the first round
        __syncthreads();
        for (int t = warp; t < ntask; t += nw) {
            const bool base_task = t < 8;
            const int fi = t - 8;
            const int r0 = base_task ? 3 * t : BASE + 3 * fi;
            const int nr = base_task ? min(3, BASE - r0) : 3;
            // pass 1: T(r, c), lanes over columns
            if (base_task) {
                for (int c = lane; c < N; c += 32) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
                    for (int k = 0; k < 22; ++k) {
                        double pv = Pi[(size_t)k * ld + c];
                        a0 += s.A[r0 * 23 + k] * pv;
                        if (nr > 1) a1 += s.A[(r0 + 1) * 23 + k] * pv;
                        if (nr > 2) a2 += s.A[(r0 + 2) * 23 + k] * pv;
                    }
                    Tw[c] = a0; Tw[Np + c] = a1; Tw[2 * Np + c] = a2;
                }
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                double bb[27], dd[9];
#pragma unroll
                for (int k = 0; k < 27; ++k) bb[k] = s.B[fi * 27 + k];
#pragma unroll
                for (int k = 0; k < 9; ++k) dd[k] = s.D[fi * 9 + k];
                const int cn = low ? r0 + 3 : N;
                for (int c = lane; c < cn; c += 32) {
                    double o0 = Own[c], o1 = Own[Np + c], o2 = Own[2 * Np + c];
                    double a0 = dd[0] * o0 + dd[1] * o1 + dd[2] * o2;
                    double a1 = dd[3] * o0 + dd[4] * o1 + dd[5] * o2;
                    double a2 = dd[6] * o0 + dd[7] * o1 + dd[8] * o2;
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        double pv = Pb[k * Np + c];
                        a0 += bb[k] * pv; a1 += bb[9 + k] * pv; a2 += bb[18 + k] * pv;
                    }
                    Tw[c] = a0; Tw[Np + c] = a1; Tw[2 * Np + c] = a2;
                }
            }
            __syncwarp();
            if (!base_task) prefetch(t + nw);   // Own is free again; overlaps with pass 2
            // pass 2: Sigma'(r, :) = T(r, :) F', lanes over base columns, then over feature blocks
            if (lane < BASE) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
                for (int k = 0; k < 22; ++k) {
                    double av = s.A[lane * 23 + k];
                    a0 += Tw[k] * av; a1 += Tw[Np + k] * av; a2 += Tw[2 * Np + k] * av;
                }
                if (r0 == lane) a0 += process_noise_diag(lane, dt);
                if (r0 + 1 == lane) a1 += process_noise_diag(lane, dt);
                if (r0 + 2 == lane) a2 += process_noise_diag(lane, dt);
                Po[(size_t)r0 * ld + lane] = prune(a0);
                if (nr > 1) Po[(size_t)(r0 + 1) * ld + lane] = prune(a1);
                if (nr > 2) Po[(size_t)(r0 + 2) * ld + lane] = prune(a2);
            }
            {
                double tb[3][9];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int k = 0; k < 9; ++k) tb[r][k] = Tw[r * Np + 7 + k];
                const int nfj = (low && !base_task) ? fi + 1 : n;
                for (int fj = lane; fj < nfj; fj += 32) {
                    double t3[3][3];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int q = 0; q < 3; ++q) t3[r][q] = Tw[r * Np + BASE + 3 * fj + q];
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) {
                        const double* bj = s.B + (3 * fj + c2) * 9;
                        const double* dj = s.D + fj * 9 + c2 * 3;
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
                        for (int k = 0; k < 9; ++k) { double bv = bj[k]; a0 += tb[0][k] * bv; a1 += tb[1][k] * bv; a2 += tb[2][k] * bv; }
#pragma unroll
                        for (int q = 0; q < 3; ++q) { double dv = dj[q]; a0 += t3[0][q] * dv; a1 += t3[1][q] * dv; a2 += t3[2][q] * dv; }
                        const int j = BASE + 3 * fj + c2;
                        if (r0 == j) a0 += process_noise_diag(j, dt);
                        if (r0 + 1 == j) a1 += process_noise_diag(j, dt);
                        if (r0 + 2 == j) a2 += process_noise_diag(j, dt);
                        Po[(size_t)r0 * ld + j] = prune(a0);
                        if (nr > 1) Po[(size_t)(r0 + 1) * ld + j] = prune(a1);
                        if (nr > 2) Po[(size_t)(r0 + 2) * ld + j] = prune(a2);
                    }
                }
            }
            __syncwarp();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    // large states: Sigma' = F Sigma F' + Q in two passes through T = F Sigma held in Pout.
    // pass 1: thread (column c, row group g) — the 22 base entries of column c stay in registers
    {
        const int CW = 64, G = PT / CW;
        const int g = tid / CW;
        for (int c0 = 0; c0 < N; c0 += CW) {
            int c = c0 + tid % CW;
            if (c < N) {
                double pb[22];
#pragma unroll
                for (int k = 0; k < 22; ++k) pb[k] = Pi[(size_t)k * ld + c];
                for (int i = yb + ysplit * g; i < BASE; i += ysplit * G) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 22; ++k) acc += s.A[i * 23 + k] * pb[k];
                    Po[(size_t)i * ld + c] = acc;
                }
                for (int r3 = yb + ysplit * g; r3 < 3 * n; r3 += ysplit * G) {
                    int fi = r3 / 3;
                    const double* b = s.B + r3 * 9;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc += b[k] * pb[7 + k];
                    const double* d = s.D + fi * 9 + (r3 % 3) * 3;
#pragma unroll
                    for (int q = 0; q < 3; ++q) acc += d[q] * Pi[(size_t)(BASE + 3 * fi + q) * ld + c];
                    Po[(size_t)(BASE + r3) * ld + c] = acc;
                }
            }
        }
    }
    __syncthreads();
    // pass 2: one warp per row, in place on the row (a CTA of a row-split launch owns base rows
    // i = yb mod ysplit and feature rows r3 = yb mod ysplit, in both passes)
    {
        const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
        const int nb_own = (BASE - yb + ysplit - 1) / ysplit, nf_own = (3 * n - yb + ysplit - 1) / ysplit;
        for (int t = warp; t < nb_own + nf_own; t += nw) {
            const int i = t < nb_own ? yb + ysplit * t : BASE + yb + ysplit * (t - nb_own);
            double* row = Po + (size_t)i * ld;
            double tb[22];
#pragma unroll
            for (int k = 0; k < 22; ++k) tb[k] = row[k];
            __syncwarp();
            if (lane < BASE) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < 22; ++k) acc += tb[k] * s.A[lane * 23 + k];
                if (i == lane) acc += process_noise_diag(i, dt);
                row[lane] = prune(acc);
            }
            for (int fj = lane; fj < n; fj += 32) {
                double t3[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) t3[q] = row[BASE + 3 * fj + q];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double* b = s.B + (3 * fj + r) * 9;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc += tb[7 + k] * b[k];
                    const double* d = s.D + fj * 9 + r * 3;
#pragma unroll
                    for (int q = 0; q < 3; ++q) acc += t3[q] * d[q];
                    int j = BASE + 3 * fj + r;
                    if (i == j) acc += process_noise_diag(j, dt);
                    row[j] = prune(acc);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// updateWithFeaturePositions, part 1 (TightlyCoupledEKF.cpp:475-580, 600-620): measurement map,
// residual, S = H Sigma H' + R, LDL^T of the upper triangle of S, K, W = Sigma H' - K S, state
// update.  The covariance update itself is ekf_joseph_general.
__device__ void gain_general_filter(const EkfPtrs& p, const double* __restrict__ Pin, const double* __restrict__ z,
                                    const double* __restrict__ Rin, const uint8_t* __restrict__ pass, double* __restrict__ Sg, int f) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int n = p.nfeat[f], N = BASE + 3 * n;
    const int ld = p.ldP, ldK = p.ldK, nmax = p.nmax, mmax = p.mmax;
    const double* Pi = Pin + (size_t)f * ld * ld;
    const double* zf = z + (size_t)f * nmax * 2;
    const double* Rf = Rin + (size_t)f * nmax * 4;
    const uint8_t* pf = pass + (size_t)f * nmax;
    int* idx = p.idx + (size_t)f * mmax;
    double* y = p.y + (size_t)f * mmax;
    double* Kf = p.K + (size_t)f * ld * ldK;
    double* Wf = p.W + (size_t)f * ld * ldK;
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * nmax * 3;
    __shared__ int s_m, s_bad;

    // formFeatureMeasurementMap (:634-661) + bookkeeping of :506-529
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            if (pf[i]) {
                idx[m] = BASE + 3 * i; idx[m + 1] = BASE + 3 * i + 1;
                y[m] = zf[2 * i] - feat_g[3 * i];
                y[m + 1] = zf[2 * i + 1] - feat_g[3 * i + 1];
                p.klt_last[((size_t)f * nmax + i) * 2] = zf[2 * i];
                p.klt_last[((size_t)f * nmax + i) * 2 + 1] = zf[2 * i + 1];
                if (Rf[4 * i + 1] != Rf[4 * i + 2]) p.asym[f] = 1;
                m += 2;
            } else {
                p.dflags[(size_t)f * nmax + i] = 1;
            }
        }
        s_m = m; s_bad = 0;
        p.m[f] = m;
    }
    __syncthreads();
    const int m = s_m;
    if (m == 0) {
        if (tid == 0) {  // K is N x 0: only the quaternion renormalisation of :605-609 acts
            double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
            mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
        }
        return;
    }
    // working matrix for the factorisation: shared memory when it fits, else global scratch
    double* Lw = (p.gain_smem_doubles >= (size_t)m * m + m) ? sm : Sg + (size_t)f * ((size_t)mmax * mmax + mmax);
    double* colv = Lw + (size_t)m * m;
    // lower(Lw)(a,b), a >= b  <-  upper(S)(b,a) = Sigma(idx[b], idx[a]) + R(b,a)   (:559-561, :578)
    for (int e = tid; e < m * m; e += blockDim.x) {
        int a = e / m, b = e % m;
        double v = 0.0;
        if (a >= b) {
            v = Pi[(size_t)idx[b] * ld + idx[a]];
            if ((a >> 1) == (b >> 1)) v += Rf[4 * ((idx[a] - BASE) / 3) + (b & 1) * 2 + (a & 1)];
        }
        Lw[e] = v;
    }
    __syncthreads();
    // right-looking LDL^T, no pivoting: unit L below the diagonal, D on the diagonal
    for (int c = 0; c < m; ++c) {
        double d = Lw[(size_t)c * m + c];
        if (d == 0.0) { if (tid == 0) { s_bad = 1; } }
        for (int r = c + 1 + tid; r < m; r += blockDim.x) {
            double v = Lw[(size_t)r * m + c];
            colv[r] = v;
            Lw[(size_t)r * m + c] = v / d;
        }
        __syncthreads();
        int rem = m - c - 1;
        for (int e = tid; e < rem * rem; e += blockDim.x) {
            int r = c + 1 + e / rem, k = c + 1 + e % rem;
            if (k <= r) Lw[(size_t)r * m + k] -= Lw[(size_t)r * m + c] * colv[k];
        }
        __syncthreads();
    }
    if (tid == 0 && s_bad) atomicOr(&p.status[f], 1);
    // K row by row: K(i,:) = S^-1 Sigma(i, idx)  (:580), entries <= 1e-13 dropped (sparseView)
    for (int i = tid; i < N; i += blockDim.x) {
#define KROW(a) Kf[kw_at(ld, i, (a))]
        const double* prow = Pi + (size_t)i * ld;
        for (int a = 0; a < m; ++a) KROW(a) = prow[idx[a]];
        for (int a = 0; a < m; ++a) {
            double v = KROW(a);
            const double* l = Lw + (size_t)a * m;
            for (int k = 0; k < a; ++k) v -= l[k] * KROW(k);
            KROW(a) = v;
        }
        for (int a = 0; a < m; ++a) KROW(a) /= Lw[(size_t)a * m + a];
        for (int a = m - 1; a >= 0; --a) {
            double v = KROW(a);
            for (int k = a + 1; k < m; ++k) v -= Lw[(size_t)k * m + a] * KROW(k);
            KROW(a) = v;
        }
        double dot = 0.0;
        for (int a = 0; a < m; ++a) { double k = prune(KROW(a)); KROW(a) = k; dot += k * y[a]; }
        for (int a = m; a < ldK; ++a) KROW(a) = 0.0;
#undef KROW
        // mu += K y  (:600)
        if (i < BASE) mu_g[i] += dot; else feat_g[i - BASE] += dot;
    }
    __syncthreads();
    // W(i,b) = Sigma(i,idx[b]) - sum_a K(i,a) S(a,b): one warp per row, lanes over b
    {
        const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
        for (int i = warp; i < N; i += nw) {
            for (int b = lane; b < ldK; b += 32) {
                double acc = 0.0;
                if (b < m) {
                    int jb = idx[b];
                    acc = Pi[(size_t)i * ld + jb];
                    for (int a = 0; a < m; ++a) acc -= Kf[kw_at(ld, i, a)] * Pi[(size_t)idx[a] * ld + jb];
                    const double* r = Rf + 4 * ((jb - BASE) / 3);
                    int b0 = b & ~1;
                    acc -= Kf[kw_at(ld, i, b0)] * r[b & 1] + Kf[kw_at(ld, i, b0 + 1)] * r[2 + (b & 1)];
                }
                Wf[kw_at(ld, i, b)] = acc;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {  // renormalise the quaternion (:605-609) and flag non-finite states
        double qn = sqrt(mu_g[3] * mu_g[3] + mu_g[4] * mu_g[4] + mu_g[5] * mu_g[5] + mu_g[6] * mu_g[6]);
        mu_g[3] /= qn; mu_g[4] /= qn; mu_g[5] /= qn; mu_g[6] /= qn;
        bool fin = true;
        for (int i = 0; i < BASE; ++i) fin = fin && isfinite(mu_g[i]);
        if (!fin) atomicOr(&p.status[f], 2);
    }
}

// One CTA per filter, or — as the fallback behind the tiled kernels (only_route >= 0) — a small persistent grid that walks over
// the batch and serves the filters the Cholesky kernel routed here (normally none: the launch is a few flag reads per CTA).
__global__ void __launch_bounds__(PT) ekf_gain_general(EkfPtrs p, const double* __restrict__ Pin, const double* __restrict__ z,
                                                       const double* __restrict__ Rin, const uint8_t* __restrict__ pass,
                                                       double* __restrict__ Sg, int only_route) {
    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        if (only_route >= 0 && p.route[f] != only_route) continue;
        gain_general_filter(p, Pin, z, Rin, pass, Sg, f);
        __syncthreads();
    }
}

// Joseph update (:586-596, :625) in selection form:
//   Sigma' = (I-KH) Sigma (I-KH)' + K R K' = Sigma - K Sigma(idx,:) - W K',  W = Sigma(:,idx) - K S
// 32x32 output tile per CTA, 256 threads (2x2 each), k-chunks of 16 staged in shared memory.
__device__ void joseph_general_tile(const EkfPtrs& p, const double* __restrict__ Pin, double* __restrict__ Pout, int f) {
    const int n = p.nfeat[f], N = BASE + 3 * n, m = p.m[f];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    if (i0 >= N || j0 >= N) return;
    const int ld = p.ldP, ldK = p.ldK;
    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Po = Pout + (size_t)f * ld * ld;
    const double* Kf = p.K + (size_t)f * ld * ldK;
    const double* Wf = p.W + (size_t)f * ld * ldK;
    const int* idx = p.idx + (size_t)f * p.mmax;
    __shared__ double As[32][17], Bs[16][33];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int phase = 0; phase < 2; ++phase) {
        const double* Am = phase == 0 ? Kf : Wf;
        for (int k0 = 0; k0 < m; k0 += 16) {
            for (int e = tid; e < 32 * 16; e += 256) {
                int r = e / 16, k = e % 16;
                As[r][k] = (i0 + r < N && k0 + k < m) ? Am[kw_at(ld, i0 + r, k0 + k)] : 0.0;
            }
            if (phase == 0) {  // B(k, j) = Sigma(idx[k], j)
                for (int e = tid; e < 16 * 32; e += 256) {
                    int k = e / 32, c = e % 32;
                    Bs[k][c] = (j0 + c < N && k0 + k < m) ? Pi[(size_t)idx[k0 + k] * ld + j0 + c] : 0.0;
                }
            } else {           // B(k, j) = K(j, k)
                for (int e = tid; e < 16 * 32; e += 256) {
                    int c = e / 16, k = e % 16;
                    Bs[k][c] = (j0 + c < N && k0 + k < m) ? Kf[kw_at(ld, j0 + c, k0 + k)] : 0.0;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                double a0 = As[ty * 2][k], a1 = As[ty * 2 + 1][k], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
                acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            int i = i0 + ty * 2 + a, j = j0 + tx * 2 + b;
            if (i < N && j < N) Po[(size_t)i * ld + j] = prune(Pi[(size_t)i * ld + j] - acc[a][b]);
        }
}

__global__ void __launch_bounds__(256) ekf_joseph_general(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout, int only_route) {
    for (int f = blockIdx.z; f < p.F; f += gridDim.z) {     // gridDim.z == F, or a small persistent grid for the fallback launch
        if (only_route >= 0 && p.route[f] != only_route) continue;
        joseph_general_tile(p, Pin, Pout, f);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void ekf_reset_kernel(EkfPtrs p, double* P0) {  // initializeBaseState (:23-56)
    const int f = blockIdx.x, tid = threadIdx.x;
    const int ld = p.ldP;
    double* P = P0 + (size_t)f * ld * ld;
    for (int e = tid; e < BASE * BASE; e += blockDim.x) {
        int i = e / BASE, j = e % BASE;
        double v = 0.0;
        if (i == j) v = (i >= 7 && i <= 15) ? 30.0 : (i >= 16 ? 0.5 : 0.0);
        P[(size_t)i * ld + j] = v;
    }
    if (tid < BASE) p.mu[(size_t)f * BASE + tid] = (tid == 3) ? 1.0 : 0.0;
    if (tid == 0) {
        p.nfeat[f] = 0; p.status[f] = 0; p.m[f] = 0; p.asym[f] = 0; p.route[f] = 0;
        double* c = p.cache + (size_t)f * 7;
        c[0] = c[1] = c[2] = 0.0; c[3] = 1.0; c[4] = c[5] = c[6] = 0.0;
    }
}

__global__ void ekf_add_features_kernel(EkfPtrs p, double* P0, const int* __restrict__ ks, const double* __restrict__ uv, int kmax) {
    // addNewFeatures (:58-94) + Feature::Feature (Feature.cpp:14-20)
    const int f = blockIdx.x, tid = threadIdx.x;
    const int k = ks[f];
    if (k <= 0) return;
    const int n0 = p.nfeat[f];
    if (n0 + k > p.nmax || k > kmax) { if (tid == 0) atomicOr(&p.status[f], 4); return; }
    const int ld = p.ldP, N0 = BASE + 3 * n0, N1 = N0 + 3 * k;
    double* P = P0 + (size_t)f * ld * ld;
    // conservativeResize: new rows and columns are empty (zero); new diagonal set
    for (int e = tid; e < (N1 - N0) * N1; e += blockDim.x) {
        int i = N0 + e / N1, j = e % N1;
        double v = 0.0;
        if (i == j) v = ((i - BASE) % 3 == 2) ? p.depth_var : p.uv_var;
        P[(size_t)i * ld + j] = v;
        if (j < N0) P[(size_t)j * ld + i] = 0.0;
    }
    for (int q = tid; q < k; q += blockDim.x) {
        int fi = n0 + q;
        double u = uv[((size_t)f * kmax + q) * 2], v = uv[((size_t)f * kmax + q) * 2 + 1];
        double* ft = p.feat + ((size_t)f * p.nmax + fi) * 3;
        ft[0] = u; ft[1] = v; ft[2] = 1.0 / p.depth;
        p.klt_last[((size_t)f * p.nmax + fi) * 2] = u;
        p.klt_last[((size_t)f * p.nmax + fi) * 2 + 1] = v;
        p.dflags[(size_t)f * p.nmax + fi] = 0;
    }
    __syncthreads();
    if (tid == 0) p.nfeat[f] = n0 + k;
}

// Feature removal (absent from the reference, which flags lost features — TightlyCoupledEKF.cpp:524-528 —
// but never deletes them): the flagged features are marginalised out, i.e. their mean entries and their rows and columns
// of Sigma are deleted and the survivors close ranks in order.  remove == nullptr: use the delete flags.
__global__ void __launch_bounds__(256) ekf_remove_features_kernel(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout,
                                                                  const uint8_t* __restrict__ remove) {
    extern __shared__ int keep[];              // [nmax] old index of surviving feature k'
    __shared__ int s_n1;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n0 = p.nfeat[f], ld = p.ldP;
    const uint8_t* rm = (remove ? remove : p.dflags) + (size_t)f * p.nmax;
    if (tid == 0) {
        int k = 0;
        for (int i = 0; i < n0; ++i) if (!rm[i]) keep[k++] = i;
        s_n1 = k;
    }
    __syncthreads();
    const int n1 = s_n1, N1 = BASE + 3 * n1;
    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Po = Pout + (size_t)f * ld * ld;
    for (int e = tid; e < N1 * N1; e += blockDim.x) {
        const int i = e / N1, j = e - i * N1;
        const int oi = i < BASE ? i : BASE + 3 * keep[(i - BASE) / 3] + (i - BASE) % 3;
        const int oj = j < BASE ? j : BASE + 3 * keep[(j - BASE) / 3] + (j - BASE) % 3;
        Po[(size_t)i * ld + j] = Pi[(size_t)oi * ld + oj];
    }
    // state vectors: survivors move down (keep[k] >= k, ascending: staged through registers so that no slot is
    // overwritten before it has been read)
    double* feat = p.feat + (size_t)f * p.nmax * 3;
    double* kl = p.klt_last + (size_t)f * p.nmax * 2;
    uint8_t* fl = p.dflags + (size_t)f * p.nmax;
    for (int k0 = 0; k0 < n0; k0 += blockDim.x) {
        const int k = k0 + tid;
        double a = 0, b = 0, c = 0, u = 0, v = 0;
        if (k < n1) { const int o = keep[k]; a = feat[3 * o]; b = feat[3 * o + 1]; c = feat[3 * o + 2]; u = kl[2 * o]; v = kl[2 * o + 1]; }
        __syncthreads();
        if (k < n0) { feat[3 * k] = a; feat[3 * k + 1] = b; feat[3 * k + 2] = c; kl[2 * k] = u; kl[2 * k + 1] = v; fl[k] = 0; }
        __syncthreads();
    }
    if (tid == 0) p.nfeat[f] = n1;
}

__global__ void ekf_check_sigma_kernel(EkfPtrs p, const double* P0, int* neg, double* asym) {  // checkSigma (:699-714)
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = p.nfeat[f], N = BASE + 3 * n, ld = p.ldP;
    const double* P = P0 + (size_t)f * ld * ld;
    int ng = 0; double mx = 0.0;
    for (int e = tid; e < N * N; e += blockDim.x) {
        int i = e / N, j = e % N;
        if (i == j) { if (P[(size_t)i * ld + i] < 0) ++ng; }
        else if (j > i) mx = fmax(mx, fabs(P[(size_t)i * ld + j] - P[(size_t)j * ld + i]));
    }
    __shared__ int s_ng[256]; __shared__ double s_mx[256];
    s_ng[tid] = ng; s_mx[tid] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) { s_ng[tid] += s_ng[tid + s]; s_mx[tid] = fmax(s_mx[tid], s_mx[tid + s]); }
        __syncthreads();
    }
    if (tid == 0) { neg[f] = s_ng[0]; asym[f] = s_mx[0]; }
}

// second half of a row-split process launch: staged state (gain panel) -> mu / feat, dq cache refresh
__global__ void ekf_commit_state_kernel(EkfPtrs p) {
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = p.nfeat[f];
    const double* stage = p.K + (size_t)f * p.ldP * p.ldK;
    double* mu_g = p.mu + (size_t)f * BASE;
    double* feat_g = p.feat + (size_t)f * p.nmax * 3;
    for (int i = tid; i < BASE; i += blockDim.x) mu_g[i] = stage[i];
    for (int i = tid; i < 3 * n; i += blockDim.x) feat_g[i] = stage[BASE + i];
    if (n > 0 && tid < 7) p.cache[(size_t)f * 7 + tid] = stage[BASE + 3 * p.nmax + tid];
}

// Completes Sigma after a lower-mode process(): feature row r was written up to its diagonal block; everything right of
// the diagonal becomes the mirror image of the column below it — the same elements the lower-mode readers take, so a
// mirror pass never changes what the update computes.  (Symmetric filters only; the others were written in full.)
__global__ void ekf_mirror_lower_kernel(EkfPtrs p, double* __restrict__ P0) {
    const int f = blockIdx.y;
    if (p.asym[f] != 0) return;
    const int N = BASE + 3 * p.nfeat[f], ld = p.ldP;
    double* P = P0 + (size_t)f * ld * ld;
    for (int r = BASE + blockIdx.x; r < N; r += gridDim.x) {
        for (int c = r + 1 + threadIdx.x; c < N; c += blockDim.x) P[(size_t)r * ld + c] = P[(size_t)c * ld + r];
    }
}

// One call of TightlyCoupledEKF::convolveBaseState / convolveFeature (:328-395, :397-460) on the device, exactly as a caller of the
// reference class gets it: convolveFeature consults the dq_inv cache keyed on omega ONLY (E2) — a hit reuses the rotation cached
// for whatever dt it was computed with, a miss recomputes it for this dt and stores it.  io[0..21] base state, io[22..24] feature,
// io[25..31] cache (in/out), io[32..56] results (22 base | 3 feature).  which: 0 base, 1 feature.
__global__ void ekf_convolve_single_kernel(double* io, double dt, int which, int fresh) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (which == 0) { convolve_base(io, dt, io + 32); return; }
    double* c = io + 25;
    Q4 dqi;
    if (!fresh && c[0] == io[10] && c[1] == io[11] && c[2] == io[12]) dqi = {c[3], c[4], c[5], c[6]};
    else {
        dqi = delta_quat(io[10], io[11], io[12], dt, -1.0);
        c[0] = io[10]; c[1] = io[11]; c[2] = io[12]; c[3] = dqi.w; c[4] = dqi.x; c[5] = dqi.y; c[6] = dqi.z;
    }
    convolve_feature(dqi, V3{io[7], io[8], io[9]}, V3{io[13], io[14], io[15]}, dt, io[22], io[23], io[24], io + 54);
}

__global__ void ekf_fill_dt_kernel(double* dts, double dt, int F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) dts[i] = dt;
}

// pack / unpack Sigma between the internal [F][ld][ld] layout and the ABI's [F][Nmax][Nmax]
__global__ void ekf_pack_P_kernel(const double* P0, double* dense, int ld, int Nmax, int to_dense) {
    const int f = blockIdx.y;
    const double* src = P0 + (size_t)f * ld * ld;
    double* dst = dense + (size_t)f * Nmax * Nmax;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)Nmax * Nmax; e += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(e / Nmax), j = (int)(e % Nmax);
        if (to_dense) dst[e] = src[(size_t)i * ld + j];
        else const_cast<double*>(src)[(size_t)i * ld + j] = dst[e];
    }
}

__global__ void ekf_accumulate_errors_kernel(EkfPtrs p, const double* __restrict__ truth, double* acc) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double ep = 0, ev = 0, eq = 0, cnt = 0;
    if (f < p.F) {
        const double* a = p.mu + (size_t)f * BASE; const double* t = truth + (size_t)f * BASE;
        for (int i = 0; i < 3; ++i) { double d = a[i] - t[i]; ep += d * d; }
        for (int i = 7; i < 10; ++i) { double d = a[i] - t[i]; ev += d * d; }
        double dp = 0, dm = 0;  // q and -q are the same rotation
        for (int i = 3; i < 7; ++i) { double d1 = a[i] - t[i], d2 = a[i] + t[i]; dp += d1 * d1; dm += d2 * d2; }
        eq = fmin(dp, dm);
        cnt = 1;
        if (!(ep == ep) || !(ev == ev)) { ep = ev = eq = 0; cnt = 0; }
    }
    double mxp = ep;
    for (int o = 16; o > 0; o >>= 1) {
        ep += __shfl_xor_sync(0xffffffff, ep, o); ev += __shfl_xor_sync(0xffffffff, ev, o);
        eq += __shfl_xor_sync(0xffffffff, eq, o); cnt += __shfl_xor_sync(0xffffffff, cnt, o);
        mxp = fmax(mxp, __shfl_xor_sync(0xffffffff, mxp, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc + 0, ep); atomicAdd(acc + 1, ev); atomicAdd(acc + 2, eq); atomicAdd(acc + 3, cnt);
        // max via CAS on the bit pattern (values are non-negative)
        unsigned long long* addr = (unsigned long long*)(acc + 4);
        unsigned long long old = *addr, assumed;
        do { assumed = old; if (__longlong_as_double(assumed) >= mxp) break; old = atomicCAS(addr, assumed, __double_as_longlong(mxp)); } while (assumed != old);
    }
}

}  // namespace

// ---- launchers -------------------------------------------------------------------------------
namespace ekfvio {

static size_t proc_smem_bytes(int nmax) { return ProcSmem::doubles(nmax) * sizeof(double); }
// fused covariance pass: + rows 7..15 of Sigma + a 3-row T slice and a 3-row prefetch slice per warp
static size_t proc_fused_smem_bytes(int nmax) {
    size_t Np = ((size_t)(BASE + 3 * nmax) + 3) & ~(size_t)3;
    return proc_smem_bytes(nmax) + sizeof(double) + (9 + 6 * (PT / 32)) * Np * sizeof(double);
}

// the fused row-block kernel is the one that knows the lower mode
bool process_lower_capable(const EkfPtrs& p) { return proc_fused_smem_bytes(p.nmax) <= 110 * 1024; }

cudaError_t launch_mirror_lower(const EkfPtrs& p, double* P0, cudaStream_t st) {
    ekf_mirror_lower_kernel<<<dim3(16, p.F), 128, 0, st>>>(p, P0);
    return cudaGetLastError();
}

cudaError_t launch_process_general(const EkfPtrs& p, const double* Pin, double* Pout, const double* dts, int mode, double* F_out, cudaStream_t st,
                                   long long* launches, int lower) {
    size_t sm = proc_smem_bytes(p.nmax);
    const int fused = (mode == 0 && proc_fused_smem_bytes(p.nmax) <= 110 * 1024) ? 1 : 0;   // two CTAs per SM
    if (fused) sm = proc_fused_smem_bytes(p.nmax);
    static size_t configured_on[64] = {0};
    size_t& configured = configured_on[current_device_slot()];
    if (sm > 48 * 1024 && sm > configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_process_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = sm;
    }
    // two-pass path: split the rows of a filter over several CTAs when the batch alone cannot fill the GPU
    int ysplit = 1;
    if (!fused && mode == 0 && p.F < 4 * 148 && (size_t)p.ldP * p.ldK >= (size_t)BASE + 3 * p.nmax + 7) {
        ysplit = (4 * 148 + p.F - 1) / p.F;
        if (ysplit > 16) ysplit = 16;
    }
    // fused path on a batch that fills the GPU: linearisation in its own small-CTA launch
    // lower mode: the covariance pass of the symmetric filters runs on DMMA tiles (ekf_process_tiles.cu) behind the same
    // linearisation launch; this kernel then only serves the filters whose Sigma is not symmetric
    const bool tiles = fused && lower && process_tiles_capable(p);
    int pre = 0;
    if (fused && (p.F >= 2 * 148 || tiles) && (size_t)p.ldP * p.ldK >= (size_t)22 * 23 + 36 * (size_t)p.nmax && proc_smem_bytes(p.nmax) <= 48 * 1024) {
        ekf_linearize_kernel<<<p.F, 128, proc_smem_bytes(p.nmax), st>>>(p, dts);
        pre = 1;
        if (launches) *launches += 1;
    }
    int low = (fused && lower) ? 1 : 0;
    if (tiles && pre) {
        cudaError_t e = launch_process_cov_tiles(p, Pin, Pout, dts, st);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
        low = 2;
    }
    ekf_process_general<<<dim3(p.F, ysplit), PT, sm, st>>>(p, Pin, Pout, dts, mode, F_out, fused, pre, low);
    if (ysplit > 1) ekf_commit_state_kernel<<<p.F, 128, 0, st>>>(p);
    if (launches) *launches += ysplit > 1 ? 2 : 1;
    return cudaGetLastError();
}

size_t gain_general_smem_doubles(int mmax) {
    size_t need = (size_t)mmax * mmax + mmax;
    return (need * sizeof(double) <= 200 * 1024) ? need : 0;
}

cudaError_t launch_gain_general(const EkfPtrs& p, const double* Pin, const double* z, const double* R, const uint8_t* pass, double* Sg,
                                cudaStream_t st, int only_route) {
    size_t sm = p.gain_smem_doubles * sizeof(double);
    static size_t configured_on[64] = {0};
    size_t& configured = configured_on[current_device_slot()];
    if (sm > 48 * 1024 && sm > configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_gain_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = sm;
    }
    const int grid = only_route >= 0 ? (p.F < 148 ? p.F : 148) : p.F;
    ekf_gain_general<<<grid, PT, sm, st>>>(p, Pin, z, R, pass, Sg, only_route);
    return cudaGetLastError();
}

cudaError_t launch_joseph_general(const EkfPtrs& p, const double* Pin, double* Pout, cudaStream_t st, int only_route) {
    int tiles = (p.Nmax + 31) / 32;
    dim3 grid(tiles, tiles, only_route >= 0 ? (p.F < 16 ? p.F : 16) : p.F);
    ekf_joseph_general<<<grid, 256, 0, st>>>(p, Pin, Pout, only_route);
    return cudaGetLastError();
}

cudaError_t launch_convolve_single(double* d_io, double dt, int which, int fresh, cudaStream_t st) {
    ekf_convolve_single_kernel<<<1, 32, 0, st>>>(d_io, dt, which, fresh);
    return cudaGetLastError();
}

cudaError_t launch_reset(const EkfPtrs& p, double* P0, cudaStream_t st) {
    ekf_reset_kernel<<<p.F, 128, 0, st>>>(p, P0);
    return cudaGetLastError();
}
cudaError_t launch_remove_features(const EkfPtrs& p, const double* Pin, double* Pout, const uint8_t* remove, cudaStream_t st) {
    ekf_remove_features_kernel<<<p.F, 256, (size_t)(p.nmax > 0 ? p.nmax : 1) * sizeof(int), st>>>(p, Pin, Pout, remove);
    return cudaGetLastError();
}

cudaError_t launch_add_features(const EkfPtrs& p, double* P0, const int* ks, const double* uv, int kmax, cudaStream_t st) {
    ekf_add_features_kernel<<<p.F, 256, 0, st>>>(p, P0, ks, uv, kmax);
    return cudaGetLastError();
}
cudaError_t launch_check_sigma(const EkfPtrs& p, const double* P0, int* neg, double* asym, cudaStream_t st) {
    ekf_check_sigma_kernel<<<p.F, 256, 0, st>>>(p, P0, neg, asym);
    return cudaGetLastError();
}
cudaError_t launch_fill_dt(double* dts, double dt, int F, cudaStream_t st) {
    ekf_fill_dt_kernel<<<(F + 255) / 256, 256, 0, st>>>(dts, dt, F);
    return cudaGetLastError();
}
cudaError_t launch_pack_P(const double* P0, double* dense, int ld, int Nmax, int F, int to_dense, cudaStream_t st) {
    dim3 grid(32, F);
    ekf_pack_P_kernel<<<grid, 256, 0, st>>>(P0, dense, ld, Nmax, to_dense);
    return cudaGetLastError();
}
cudaError_t launch_accumulate_errors(const EkfPtrs& p, const double* truth, double* acc, cudaStream_t st) {
    ekf_accumulate_errors_kernel<<<(p.F + 255) / 256, 256, 0, st>>>(p, truth, acc);
    return cudaGetLastError();
}

}  // namespace ekfvio
