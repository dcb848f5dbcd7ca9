// Shared device helpers for the DMMA tile kernels (ekf_tiled.cu, ekf_large.cu).
#pragma once
#include "ekf_common.cuh"

namespace ekfvio {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// A global load the compiler may not sink below later volatile asm (the DMMAs): keeps software
// prefetches one block ahead of their use.
__device__ __forceinline__ double ldg_pinned(const double* p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


__device__ __forceinline__ int tsw(int r, int c) { return r * 8 + (c ^ (((r >> 1) & 1) << 2)); }
__device__ __forceinline__ int tile_of(int ib, int jb) { return (ib * (ib + 1) / 2 + jb) * 64; }

// C-fragment (row = lane/4, cols 2q,2q+1) -> the two A-fragments (row = lane/4, k = q + 4kk)
__device__ __forceinline__ void cfrag_to_afrag(double c0, double c1, int lane, double& a0, double& a1) {
    const int q = lane & 3, base = lane & ~3;
    double v0 = __shfl_sync(0xffffffffu, c0, base | (q >> 1));
    double v1 = __shfl_sync(0xffffffffu, c1, base | (q >> 1));
    a0 = (q & 1) ? v1 : v0;
    v0 = __shfl_sync(0xffffffffu, c0, base | 2 | (q >> 1));
    v1 = __shfl_sync(0xffffffffu, c1, base | 2 | (q >> 1));
    a1 = (q & 1) ? v1 : v0;
}


// Blocked right-looking factorisation of an (8*nb) x (8*nb) symmetric matrix held as swizzled lower-triangular 8x8 tiles in
// shared memory (Ls), in place; Li receives the explicit inverses of the diagonal tiles.  Called by all NWC warps of the CTA.
//
// sgn == nullptr: Cholesky, A = L L'; *bad_flag is set if a pivot is not positive.
// sgn != nullptr (shared memory, 8*nb floats-as-doubles): the SIGNED factorisation A = L J L', J = diag(sgn), sgn = +-1 — the
// unpivoted LDL^T of the reference (SimplicialLDLT, TightlyCoupledEKF.cpp:577-580: L_ldlt = L |D|^-1/2 ... i.e. L = L_ldlt |D|^1/2,
// J = sign(D)) in Cholesky clothing, so that an S that is not positive definite is carried on with exactly as the reference
// does.  *bad_flag is then set only for a zero (or NaN) pivot, *neg_flag if any pivot was negative.
template <int NWC>
__device__ void chol_tiles(double* Ls, double* Li, int nb, int* bad_flag, double* sgn = nullptr, int* neg_flag = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, q = lane & 3;
    for (int jb = 0; jb < nb; ++jb) {
        if (warp == 0) {   // diagonal tile: lane rr (< 8) owns row rr (the other lanes replicate)
            // Elimination in LDL^T form: one reciprocal per column on the dependency chain, the square roots (one per row, after the
            // loop) only scale the finished columns; L = L_ldl |D|^1/2, J = sign D.
            double* T = Ls + tile_of(jb, jb);
            const int rr = lane & 7;
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[tsw(rr, c)]); a[c] = v.x; a[c + 1] = v.y; }
            double own = 1.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double d = __shfl_sync(0xffffffffu, a[c], c);            // pivot (signed)
                double l[8];
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2) l[c2] = __shfl_sync(0xffffffffu, a[c], c2);
                if (c == rr) own = d;
                const double t = a[c] * __drcp_rn(d);
#pragma unroll
                for (int c2 = c + 1; c2 < 8; ++c2)
                    if (rr >= c2) a[c2] -= t * l[c2];
            }
            double sc = 1.0, ad = own;
            bool bad = false, neg = false;
            if (sgn) {
                if (own < 0.0) { sc = -1.0; ad = -own; neg = true; }
                if (!(ad > 0.0)) bad = true;               // zero or NaN pivot
            } else if (!(own > 0.0)) bad = true;
            const double piv = sqrt(ad), myrs = 1.0 / piv;
            bad = __any_sync(0xffffffffu, bad); neg = __any_sync(0xffffffffu, neg);
            double rs[8], sg[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { rs[c] = __shfl_sync(0xffffffffu, myrs, c); sg[c] = __shfl_sync(0xffffffffu, sc, c); }
            // scaled factor: L(rr, c) = a(rr, c) sign_c / sqrt|d_c|
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = (c < rr) ? a[c] * (rs[c] * sg[c]) : ((c == rr) ? piv : 0.0);
            // column j = rr of inv(L): forward substitution with rows fetched by shuffle
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
                for (int c = 0; c < 8; ++c) T[tsw(rr, c)] = (c <= rr) ? a[c] : 0.0;
                double* I8 = Li + jb * 64;
#pragma unroll
                for (int i = 0; i < 8; ++i) I8[tsw(i, rr)] = (i >= rr) ? x[i] : 0.0;
                if (bad && lane == 0) *bad_flag = 1;
                if (sgn) {
                    if (neg && lane == 0) *neg_flag = 1;
                    sgn[jb * 8 + rr] = sc;
                }
            }
        }
        __syncthreads();
        // signs of this block column: J scales the columns of the panel and the k index of the trailing update
        double s2q0 = 1.0, s2q1 = 1.0, sk0 = 1.0, sk1 = 1.0;
        if (sgn) { s2q0 = sgn[jb * 8 + 2 * q]; s2q1 = sgn[jb * 8 + 2 * q + 1]; sk0 = sgn[jb * 8 + q]; sk1 = sgn[jb * 8 + 4 + q]; }
        {   // panel: L(ib,jb) = A(ib,jb) * inv(Ljj)' * J_jb
            const double* I8 = Li + jb * 64;
            double b0 = I8[tsw(r, q)], b1 = I8[tsw(r, 4 + q)];
            for (int ib = jb + 1 + warp; ib < nb; ib += NWC) {
                double* T = Ls + tile_of(ib, jb);
                double a0 = T[tsw(r, q)], a1 = T[tsw(r, 4 + q)];
                double c0 = 0.0, c1 = 0.0;
                dmma884(c0, c1, a0, b0);
                dmma884(c0, c1, a1, b1);
                __syncwarp();
                *reinterpret_cast<double2*>(&T[tsw(r, 2 * q)]) = make_double2(c0 * s2q0, c1 * s2q1);
            }
        }
        __syncthreads();
        {   // trailing update: A(ib,kb) -= L(ib,jb) J_jb L(kb,jb)'  for ib >= kb > jb
            // lower-triangular tile list (ii, kk2), kk2 <= ii < t, dealt round-robin to the warps; the pair is advanced
            // incrementally (a square root per tile used to cost as much as the two DMMAs it addressed)
            const int t = nb - 1 - jb;
            int ii = 0, kk2 = warp;
            while (kk2 > ii) { kk2 -= ii + 1; ++ii; }
            for (; ii < t; ) {
                int ib = jb + 1 + ii, kb = jb + 1 + kk2;
                const double* TA = Ls + tile_of(ib, jb);
                const double* TB = Ls + tile_of(kb, jb);
                double* TC = Ls + tile_of(ib, kb);
                double2 c = *reinterpret_cast<double2*>(&TC[tsw(r, 2 * q)]);
                dmma884(c.x, c.y, -TA[tsw(r, q)] * sk0, TB[tsw(r, q)]);
                dmma884(c.x, c.y, -TA[tsw(r, 4 + q)] * sk1, TB[tsw(r, 4 + q)]);
                *reinterpret_cast<double2*>(&TC[tsw(r, 2 * q)]) = c;
                kk2 += NWC;
                while (kk2 > ii) { kk2 -= ii + 1; ++ii; }
            }
        }
        __syncthreads();
    }
}

}  // namespace ekfvio
