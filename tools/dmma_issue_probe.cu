// Probe: DMMA.8x8x4 issue interval per warp on sm_100a as a function of resident warps and independent accumulator chains,
// next to DFMA.  Register operands only.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_issue_probe tools/dmma_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NCH, int DF>
__global__ void k(double* out, long long* clk, int iters) {
    double c0[NCH], c1[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (DF) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
            else dmma(c0[i], c1[i], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c0[i] + c1[i];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int NCH, int DF> void run(int warps) {
    double* out; long long* clk; cudaMalloc(&out, 8); cudaMalloc(&clk, 8);
    const int iters = 2000;
    k<NCH, DF><<<148, 32 * warps>>>(out, clk, iters); cudaDeviceSynchronize();
    k<NCH, DF><<<148, 32 * warps>>>(out, clk, iters); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * NCH * (DF ? 2 : 1));
    printf("%s warps/SM %2d chains %d: %.1f clk per %s per warp -> %.1f clk per op per SMSP\n", DF ? "DFMA" : "DMMA", warps, NCH, per, DF ? "DFMA" : "DMMA", per / ((warps + 3) / 4));
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<1, 0>(w); run<4, 0>(w); run<8, 0>(w); }
    for (int w : {4, 8, 16}) { run<1, 1>(w); run<8, 1>(w); }
    return 0;
}
