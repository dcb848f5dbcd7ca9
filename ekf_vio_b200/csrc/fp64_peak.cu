// FP64 peak probe: register-resident DMMA.8x8x4 and DFMA loops, timed with CUDA events.
// This is the EKF roofline denominator (MEASURED_PEAKS.json carries no FP64 figure).
#include "ekf_common.cuh"
#include "ekf_kernels.h"

namespace {
__global__ void k_peak_dmma(double* out, int iters, double s) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
    double a = 1.0 + s * threadIdx.x, b = s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ekfvio::dmma884_volatile(c0[i], c1[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += c0[i] + c1[i];
    if (r == 123.456) out[0] = r;
}
__global__ void k_peak_dfma(double* out, int iters, double s) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + s, b = s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += acc[i];
    if (r == 123.456) out[0] = r;
}
}  // namespace

namespace ekfvio {
cudaError_t measure_fp64_peak(double* dmma_tflops, double* dfma_tflops) {
    cudaDeviceProp prop; int dev = 0;
    cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    e = cudaGetDeviceProperties(&prop, dev); if (e != cudaSuccess) return e;
    double* out = nullptr;
    e = cudaMalloc((void**)&out, 64); if (e != cudaSuccess) return e;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = prop.multiProcessorCount * 4, tpb = 512, iters = 20000;
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            if (which == 0) k_peak_dmma<<<blocks, tpb>>>(out, iters, 1e-9); else k_peak_dfma<<<blocks, tpb>>>(out, iters, 1e-9);
            cudaEventRecord(b);
            e = cudaEventSynchronize(b); if (e != cudaSuccess) break;
            float ms = 0; cudaEventElapsedTime(&ms, a, b);
            double fl = which == 0 ? 512.0 * blocks * (tpb / 32) * 8.0 * iters : 2.0 * blocks * (double)tpb * 8.0 * iters;
            double tf = fl / ms * 1e-9;
            if (rep > 0 && tf > best[which]) best[which] = tf;
        }
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    if (dmma_tflops) *dmma_tflops = best[0];
    if (dfma_tflops) *dfma_tflops = best[1];
    return e;
}
}  // namespace ekfvio
