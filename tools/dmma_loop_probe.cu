// Probe: what limits a shared-memory-fed DMMA.8x8x4 inner loop at one CTA per SM?
// Variants: warps per CTA, blocks (2x2 tiles) per warp, barrier per chunk, fragment sharing.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int SLDA = 20, ROWS = 176;

// MODE 0: per item 2 A + 2 B loads (1:1).  MODE 1: A shared by item pairs (3 LDS per 4 DMMA -> 0.75).
// MODE 2: 4x2-tile items (32x16): 4 A + 2 B per 8 DMMA (0.75).   BAR: __syncthreads every chunk.
template <int IPW, int MODE, int BAR>
__global__ void k_loop(double* out, int chunks, int nwarps_active) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 2 * ROWS * SLDA; i += blockDim.x) sm[i] = 1e-9 * (i & 15);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const double* A = sm; const double* B = sm + ROWS * SLDA;
    int aoff[IPW], boff[IPW];
#pragma unroll
    for (int it = 0; it < IPW; ++it) { int e = (warp + it * 11) % 66; int i = e % 11, j = (e * 7) % 11; aoff[it] = (i * 16 + r) * SLDA + q; boff[it] = (j * 16 + r) * SLDA + q; }
    double c0[IPW][4], c1[IPW][4];
#pragma unroll
    for (int it = 0; it < IPW; ++it)
#pragma unroll
        for (int t = 0; t < 4; ++t) { c0[it][t] = 0; c1[it][t] = 0; }
    for (int c = 0; c < chunks; ++c) {
        if (BAR) __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int it = 0; it < IPW; ++it) {
                double a0, a1, b0, b1;
                if (MODE == 1 && (it & 1)) { a0 = A[aoff[it - 1] + kk * 4]; a1 = A[aoff[it - 1] + 8 * SLDA + kk * 4]; }
                else { a0 = A[aoff[it] + kk * 4]; a1 = A[aoff[it] + 8 * SLDA + kk * 4]; }
                b0 = B[boff[it] + kk * 4]; b1 = B[boff[it] + 8 * SLDA + kk * 4];
                dmma(c0[it][0], c1[it][0], a0, b0);
                dmma(c0[it][1], c1[it][1], a0, b1);
                dmma(c0[it][2], c1[it][2], a1, b0);
                dmma(c0[it][3], c1[it][3], a1, b1);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int it = 0; it < IPW; ++it)
#pragma unroll
        for (int t = 0; t < 4; ++t) s += c0[it][t] + c1[it][t];
    if (s == 123.456) out[0] = s;
}

template <int IPW, int MODE, int BAR> void run(const char* name, int warps, int ctas_per_sm, int sms, double* out) {
    size_t smem = 2 * ROWS * SLDA * sizeof(double);
    CK(cudaFuncSetAttribute(k_loop<IPW, MODE, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = 2000;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k_loop<IPW, MODE, BAR><<<sms * ctas_per_sm, warps * 32, smem>>>(out, chunks, warps);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    double fl = 512.0 * (double)sms * ctas_per_sm * warps * IPW * 16.0 * chunks;
    printf("%-34s warps=%2d ctas/sm=%d ipw=%d  %.2f TFLOP/s\n", name, warps, ctas_per_sm, IPW, fl / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    run<6, 0, 1>("1:1 loads, barrier", 11, 1, sms, out);
    run<6, 0, 0>("1:1 loads, no barrier", 11, 1, sms, out);
    run<6, 0, 1>("1:1 loads, barrier", 12, 1, sms, out);
    run<6, 0, 0>("1:1 loads, no barrier", 12, 1, sms, out);
    run<6, 0, 0>("1:1 loads, no barrier", 16, 1, sms, out);
    run<3, 0, 1>("1:1 loads, barrier", 22, 1, sms, out);
    run<3, 0, 0>("1:1 loads, no barrier", 22, 1, sms, out);
    run<3, 0, 0>("1:1 loads, no barrier", 24, 1, sms, out);
    run<6, 1, 1>("A shared by pairs, barrier", 11, 1, sms, out);
    run<6, 1, 0>("A shared by pairs, no barrier", 12, 1, sms, out);
    run<6, 0, 0>("1:1, 2 CTAs/SM", 8, 2, sms, out);
    run<4, 0, 0>("1:1 ipw=4", 16, 1, sms, out);
    run<2, 0, 0>("1:1 ipw=2", 32, 1, sms, out);
    return 0;
}
