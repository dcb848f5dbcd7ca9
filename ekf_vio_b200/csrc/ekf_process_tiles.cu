// Covariance pass of process(dt) on FP64 tensor-core tiles (DMMA.8x8x4) for symmetric filters in lower mode.
// Reference: TightlyCoupledEKF::process, include/ekf_vio/TightlyCoupledEKF.cpp:96-121 — Sigma' = F Sigma F' + Q dt, then prune.
//
// F = [[A, 0], [B, D]] with A 22x22, B (3n x 9, columns 7..15 of the base state) and D block diagonal (3x3 per feature), all
// left in the idle W panel by ekf_linearize_kernel.  For feature row blocks I, J (and b = base rows, b9 = base rows 7..15):
//     T_I(:, b)    = B_I Sigma(b9, b) + D_I Sigma(I, b)                                  (3 x 22)        U_I = T_I(:, b9)
//     Sigma'(I, b) = T_I(:, b) A'                                                         Sigma'(b, I) = its transpose
//     Sigma'(I, J) = U_I B_J' + B_I C_J' + D_I Sigma(I, J) D_J'        with  C_J = D_J Sigma(b9, J)'     (3 x 9)
//     Sigma'(b, b) = A Sigma(b, b) A'
// i.e. per 24-row group (eight features) three small GEMMs on 8x8 tiles — [U_I | B_I] (24 x 18) times [B_J | C_J]' for every
// column tile up to the group's own diagonal — on top of the block-scaled term, which is the only part that stays on DFMA
// (a thread per 3x3 block, in place in shared memory, then taken as the accumulator the DMMAs start from).  Per filter:
// ~1 800 DMMAs + 0.13 M DFMAs, against 0.5 M DFMAs with one shared-memory operand per three of them in the row-block kernel
// (ekf_process_general, fused branch), which stays the path of asymmetric filters, full-matrix mode and larger states.
//
// Lower form, as ekf_process_general's lower mode writes it and ekf_update_fused reads it: rows 0..21 complete, a feature row
// up to the end of its own 3x3 diagonal block.  Everything this kernel reads lies inside that form.
#include "ekf_common.cuh"
#include "ekf_kernels.h"

#include <cstdlib>

using namespace ekfvio;

namespace {

constexpr int CT = 256;     // threads per CTA (8 warps), two CTAs per SM
constexpr int LDT = 28;     // row stride of the 24 x 24 operand tiles (== 12 mod 16: conflict-free A- and B-fragment loads)
constexpr int LDR = 20;     // row stride of Rm = [B | C | 0 0]  (== 4 mod 16)

__host__ __device__ inline int own_stride(int ldP) { return (ldP % 16 == 8) ? ldP : ldP + 8; }   // == 8 mod 16: conflict-free double2 rows
__host__ __device__ inline int cov_groups(int nmax) { return (nmax + 7) / 8; }
__host__ __device__ inline size_t cov_smem_doubles(int nmax, int ldP) {
    return 3 * 24 * LDT + (((size_t)9 * nmax + 1) & ~(size_t)1) + (size_t)24 * cov_groups(nmax) * LDR + (size_t)24 * own_stride(ldP) + (size_t)9 * ldP;
}

#ifdef EKFVIO_PROFILE_CLOCKS
__device__ unsigned long long g_pclk[16];
#define PCLK(i) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_pclk[(i)], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define PCLK(i) do {} while (0)
#endif

// 24 x 24 product on 3 x 3 tiles, unit u = rt * 3 + ct: C(rt, ct) += sum_k a(row, k) b(k, col); the callers pass fragment loaders
template <class FA, class FB>
__device__ __forceinline__ void tile_mma(double& c0, double& c1, int ksteps, FA fa, FB fb) {
    for (int s = 0; s < ksteps; ++s) dmma884(c0, c1, fa(s), fb(s));
}

__global__ void __launch_bounds__(CT, 2) ekf_process_cov_tiles(EkfPtrs p, const double* __restrict__ Pin, double* __restrict__ Pout,
                                                               const double* __restrict__ dts) {
    extern __shared__ double sm[];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (p.asym[f] != 0) return;                      // full-matrix filters: ekf_process_general behind this launch
    const int r = lane >> 2, q = lane & 3;
    const int n = p.nfeat[f], N = BASE + 3 * n, ld = p.ldP, nmax = p.nmax;
    const int ldo = own_stride(ld);
    const double dt = dts[f];
    double* As = sm;                                  // A, zero padded to 24 x 24
    double* Sbb = As + 24 * LDT;                      // Sigma(b, b), zero padded
    double* Tb = Sbb + 24 * LDT;                      // T(:, b) of the current row group
    double* Ds = Tb + 24 * LDT;                       // D blocks
    double* Rm = Ds + ((9 * nmax + 1) & ~1);          // [B | C | 0 0] per feature row
    double* Own = Rm + 24 * cov_groups(nmax) * LDR;   // the current group's 24 rows of Sigma
    double* Pb9 = Own + 24 * ldo;                     // rows 7..15 of Sigma (only while the C blocks are formed)
    const double* Pi = Pin + (size_t)f * ld * ld;
    double* Po = Pout + (size_t)f * ld * ld;
    const double* lin = p.W + (size_t)f * ld * p.ldK;
    const double* linB = lin + 22 * 23;
    const double* linD = linB + 27 * nmax;
    const int ngroups = (n + 7) >> 3;
    const int ncta = (3 * n + 7) >> 3;                // feature column tiles that hold anything

    // rows of group g by cp.async: row rho (feature-relative) up to the end of its own diagonal block
    auto load_group = [&](int g) {
        const int rows = min(24, 3 * n - 24 * g);
        for (int rr = warp; rr < rows; rr += CT / 32) {
            const int rho = 24 * g + rr;
            const int ext = (BASE + 3 * (rho / 3) + 3 + 1) >> 1;     // 16-byte chunks
            const double* src = Pi + (size_t)(BASE + rho) * ld;
            double* dst = Own + rr * ldo;
            for (int c2 = lane; c2 < ext; c2 += 32) {
                unsigned d = (unsigned)__cvta_generic_to_shared(dst + 2 * c2);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 2 * c2) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#ifdef EKFVIO_PROFILE_CLOCKS
    long long t_prev = clock64();
#endif
    if (ngroups > 0) load_group(0);

    // ---- per-filter operands: every global read of the prologue is a cp.async issued up front (no dependent load rounds) ----
    auto cpa8 = [](double* dst, const double* src) {
        unsigned d = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    };
    for (int e = tid; e < 24 * LDT; e += CT) {
        const int i = e / LDT, j = e - i * LDT;
        if (i < 22 && j < 22) { cpa8(As + e, lin + i * 23 + j); cpa8(Sbb + e, Pi + (size_t)i * ld + j); }
        else { As[e] = 0.0; Sbb[e] = 0.0; }
    }
    for (int e = tid; e < 9 * n; e += CT) cpa8(Ds + e, linD + e);
    {   // Rm: B part by cp.async, zero columns 18, 19, rows beyond 3n zero (the C part is formed below)
        const int rows = 24 * cov_groups(nmax);
        for (int e = tid; e < rows * 11; e += CT) {
            const int row = e / 11, c = e - row * 11;
            const int col = c < 9 ? c : c + 9;       // 0..8, 18, 19
            if (c < 9 && row < 3 * n) cpa8(Rm + row * LDR + col, linB + row * 9 + c);
            else Rm[row * LDR + col] = 0.0;
        }
        const int half = (N + 1) >> 1;
        for (int e = tid; e < 9 * half; e += CT) {
            const int k = e / half, c2 = e - k * half;
            unsigned d = (unsigned)__cvta_generic_to_shared(Pb9 + k * ld + 2 * c2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Pi + (size_t)(7 + k) * ld + 2 * c2) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    PCLK(0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    PCLK(1);
    // base rows: T = A Sigma(b, b) on 3 x 3 tiles (unit u = rt * 3 + ct), then Sigma'(b, b) = T A' + Q
    for (int u = warp; u < 9; u += CT / 32) {
        const int rt = u / 3, ct = u - 3 * rt;
        double c0 = 0.0, c1 = 0.0;
        tile_mma(c0, c1, 6, [&](int s) { return As[(8 * rt + r) * LDT + 4 * s + q]; }, [&](int s) { return Sbb[(4 * s + q) * LDT + 8 * ct + r]; });
        *reinterpret_cast<double2*>(&Tb[(8 * rt + r) * LDT + 8 * ct + 2 * q]) = make_double2(c0, c1);
    }
    {   // C_j(q, k) = sum_q' D_j(q, q') Sigma(7 + k, 22 + 3 j + q')
        const int rows = 24 * cov_groups(nmax);
        for (int row = tid; row < rows; row += CT) {
            const int j = row / 3, qq = row - 3 * j;
            const bool live = row < 3 * n;
            double d0 = 0.0, d1 = 0.0, d2 = 0.0;
            if (live) { const double* d = Ds + j * 9 + qq * 3; d0 = d[0]; d1 = d[1]; d2 = d[2]; }
            const double* sg = Pb9 + BASE + 3 * j;
#pragma unroll
            for (int k = 0; k < 9; ++k) Rm[row * LDR + 9 + k] = live ? d0 * sg[k * ld] + d1 * sg[k * ld + 1] + d2 * sg[k * ld + 2] : 0.0;
        }
    }
    __syncthreads();
    for (int u = warp; u < 9; u += CT / 32) {
        const int rt = u / 3, ct = u - 3 * rt;
        double c0 = 0.0, c1 = 0.0;
        tile_mma(c0, c1, 6, [&](int s) { return Tb[(8 * rt + r) * LDT + 4 * s + q]; }, [&](int s) { return As[(8 * ct + r) * LDT + 4 * s + q]; });
        const int gr = 8 * rt + r, gc = 8 * ct + 2 * q;
        if (gr == gc) c0 += process_noise_diag(gr, dt);
        if (gr == gc + 1) c1 += process_noise_diag(gr, dt);
        if (gr < BASE) {
            if (gc < BASE) Po[(size_t)gr * ld + gc] = prune(c0);
            if (gc + 1 < BASE) Po[(size_t)gr * ld + gc + 1] = prune(c1);
        }
    }
    PCLK(2);

    // ---- feature row groups ----
    for (int g = 0; g < ngroups; ++g) {
        const int nfe = min(8, n - 8 * g);            // features of this group
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                              // Own(g) landed; the base tiles / previous group are done with Tb
        PCLK(3);
        // T(:, b) = D_I Sigma(I, b) + B_I Sigma(b9, b) on 3 x 3 tiles: the block-scaled part is formed in the accumulator layout, then
        // K = 9 (lanes whose k runs past it read a zero of Rm)
        for (int k = 0; k < 2; ++k) {                        // (the ninth tile goes to the last warp: its share of the block scaling below is the smallest)
            if (k == 1 && warp != CT / 32 - 1) break;
            const int u = k == 0 ? warp : 8;
            const int rt = u / 3, ct = u - 3 * rt, row = 8 * rt + r, il = row / 3, rp = row - 3 * il, col = 8 * ct + 2 * q;
            double c0 = 0.0, c1 = 0.0;
            if (il < nfe && col < BASE) {
                const double* d = Ds + (8 * g + il) * 9 + rp * 3;
                const double* o = Own + (3 * il) * ldo + col;
                c0 = d[0] * o[0] + d[1] * o[ldo] + d[2] * o[2 * ldo];
                c1 = d[0] * o[1] + d[1] * o[ldo + 1] + d[2] * o[2 * ldo + 1];
            }
            const double* arow = Rm + (24 * g + row) * LDR;
            tile_mma(c0, c1, 3, [&](int s) { const int k = 4 * s + q; return arow[k < 9 ? k : 18]; },
                     [&](int s) { return Sbb[(7 + 4 * s + q) * LDT + 8 * ct + r]; });
            *reinterpret_cast<double2*>(&Tb[row * LDT + col]) = make_double2(c0, c1);
        }
        // (b) X_ij = D_i Sigma_ij D_j' in place, blocks on and below the diagonal
        {
            const int J = 8 * g + 8;
            for (int e = tid; e < 8 * J; e += CT) {
                const int i = e / J, j = e - i * J;
                if (i >= nfe || j > 8 * g + i) continue;
                const double* di = Ds + (8 * g + i) * 9;
                const double* dj = Ds + j * 9;
                double* o = Own + (3 * i) * ldo + BASE + 3 * j;
                double sg[3][3], m[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) sg[a][b] = o[a * ldo + b];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) m[a][b] = di[a * 3] * sg[0][b] + di[a * 3 + 1] * sg[1][b] + di[a * 3 + 2] * sg[2][b];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) o[a * ldo + b] = m[a][0] * dj[b * 3] + m[a][1] * dj[b * 3 + 1] + m[a][2] * dj[b * 3 + 2];
            }
        }
        __syncthreads();
        PCLK(4);
        PCLK(5);
        // Sigma'(I, b) = T(:, b) A' and its transpose Sigma'(b, I)
        for (int u = warp; u < 9; u += CT / 32) {
            const int rt = u / 3, ct = u - 3 * rt;
            double c0 = 0.0, c1 = 0.0;
            tile_mma(c0, c1, 6, [&](int s) { return Tb[(8 * rt + r) * LDT + 4 * s + q]; }, [&](int s) { return As[(8 * ct + r) * LDT + 4 * s + q]; });
            const int gr = BASE + 24 * g + 8 * rt + r, gc = 8 * ct + 2 * q;
            if (gr < N) {
                c0 = prune(c0); c1 = prune(c1);
                if (gc < BASE) { Po[(size_t)gr * ld + gc] = c0; Po[(size_t)gc * ld + gr] = c0; }
                if (gc + 1 < BASE) { Po[(size_t)gr * ld + gc + 1] = c1; Po[(size_t)(gc + 1) * ld + gr] = c1; }
            }
        }
        PCLK(6);
        // Sigma'(I, J) = X + [U_I | B_I] [B_J | C_J]', U_I = T(:, b9): a warp takes up to three column tiles, one after the other, and all
        // three row tiles of each.  The accumulators and the A fragments are taken out of Own / Tb first, so that the next
        // group's rows travel while the DMMAs run.
        {
            double a[3][5];
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int k = 4 * s + q;
#pragma unroll
                for (int rt = 0; rt < 3; ++rt)
                    a[rt][s] = (k < 9) ? Tb[(8 * rt + r) * LDT + 7 + k] : Rm[(24 * g + 8 * rt + r) * LDR + (k < 18 ? k - 9 : k)];
            }
            const int nct = min(3 * (g + 1), ncta);
            // column tile of slot un: the first round goes up the warps, the second and third come back down, so that the warps with
            // two T / base-column tiles are not the ones with the most column tiles
            auto ctof = [&](int un) { return un == 0 ? warp : (un + 1) * (CT / 32) - 1 - warp; };
            double2 c[3][3];
#pragma unroll
            for (int un = 0; un < 3; ++un) {
                const int ct = ctof(un);
#pragma unroll
                for (int rt = 0; rt < 3; ++rt)
                    c[un][rt] = (ct < nct) ? *reinterpret_cast<const double2*>(&Own[(8 * rt + r) * ldo + BASE + 8 * ct + 2 * q]) : make_double2(0.0, 0.0);
            }
            __syncthreads();                          // everybody is done with Own and Tb
            PCLK(7);
            if (g + 1 < ngroups) load_group(g + 1);
            const double qf = 0.0001 * dt;            // process_noise_diag of a feature row
#pragma unroll
            for (int un = 0; un < 3; ++un) {
                const int ct = ctof(un);
                if (ct >= nct) continue;
                const double* b0 = Rm + (8 * ct + r) * LDR + q;
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const double bv = b0[4 * s];
#pragma unroll
                    for (int rt = 0; rt < 3; ++rt) dmma884(c[un][rt].x, c[un][rt].y, a[rt][s], bv);
                }
                const int gamma = 8 * ct + 2 * q;                                 // feature-relative column
                double* dst = Po + (size_t)(BASE + 24 * g + r) * ld + BASE + gamma;
                if (ct < 3 * g && 24 * g + 24 <= 3 * n) {                          // left of the group's diagonal square, all rows live
#pragma unroll
                    for (int rt = 0; rt < 3; ++rt)
                        *reinterpret_cast<double2*>(dst + (size_t)8 * rt * ld) = make_double2(prune(c[un][rt].x), prune(c[un][rt].y));
                } else {
#pragma unroll
                    for (int rt = 0; rt < 3; ++rt) {
                        const int rho = 24 * g + 8 * rt + r;                      // feature-relative row
                        if (rho >= 3 * n) continue;
                        double v0 = c[un][rt].x, v1 = c[un][rt].y;
                        if (rho == gamma) v0 += qf;
                        if (rho == gamma + 1) v1 += qf;
                        const int lim = 3 * (rho / 3) + 3;                        // end of the row's own diagonal block
                        if (gamma + 1 < lim) *reinterpret_cast<double2*>(dst + (size_t)8 * rt * ld) = make_double2(prune(v0), prune(v1));
                        else if (gamma < lim) dst[(size_t)8 * rt * ld] = prune(v0);
                    }
                }
            }
        }
        PCLK(8);
    }
}

}  // namespace

namespace ekfvio {

bool process_tiles_capable(const EkfPtrs& p) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("EKFVIO_NO_TILED_PROCESS"); off = (e && e[0] == '1') ? 1 : 0; }
    return !off && !(p.flags & 0x1000u) && p.nmax >= 1 && (3 * p.nmax + 7) / 8 <= 3 * (CT / 32)   /* 0x1000: debug bit, row-block kernel for everything */ && cov_smem_doubles(p.nmax, p.ldP) * sizeof(double) <= 112 * 1024 &&
           (size_t)p.ldP * p.ldK >= (size_t)22 * 23 + 36 * (size_t)p.nmax;
}

cudaError_t launch_process_cov_tiles(const EkfPtrs& p, const double* Pin, double* Pout, const double* dts, cudaStream_t st) {
    const size_t smem = cov_smem_doubles(p.nmax, p.ldP) * sizeof(double);
    static size_t configured_on[64] = {0};
    size_t& configured = configured_on[current_device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ekf_process_cov_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    ekf_process_cov_tiles<<<p.F, CT, smem, st>>>(p, Pin, Pout, dts);
    return cudaGetLastError();
}

}  // namespace ekfvio

#ifdef EKFVIO_PROFILE_CLOCKS
// thread 0 of every CTA: 0 prologue issue, 1 prologue wait, 2 C blocks + base tiles, per group 3 wait for the rows, 4 block scaling,
// 5 T GEMM, 6 base-column GEMM, 7 fragment loads + barrier, 8 DMMAs + stores
extern "C" void ekfvio_debug_cov_clocks(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_pclk, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_pclk, z, sizeof(z)); }
}
#endif
