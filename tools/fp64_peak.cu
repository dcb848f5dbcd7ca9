// FP64 peak probe for B200 (sm_100a): DFMA issue peak, DMMA.8x8x4 peak (register
// resident and shared-memory fed), and a device copy for HBM GB/s.  Prints one JSON line.
// Used to fix the FP64 roofline denominator that MEASURED_PEAKS.json does not carry.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double s) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + s, b = s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += acc[i];
    if (r == 123.456) out[0] = r;
}

template <int NT>
__global__ void k_dmma(double *out, int iters, double s) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = i; c1[i] = -i; }
    double a = 1.0 + s * threadIdx.x, b = s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) r += c0[i] + c1[i];
    if (r == 123.456) out[0] = r;
}

// smem-fed: warp tile (8*TA) x (8*TB), operands re-read from shared memory every k-step.
template <int TA, int TB>
__global__ void k_dmma_smem(double *out, int iters, double s) {
    extern __shared__ double sm[];
    const int LD = 108;  // 108 mod 16 == 12 -> conflict-free fragment loads
    for (int i = threadIdx.x; i < 64 * LD * 2; i += blockDim.x) sm[i] = s * (i & 7);
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *A = sm + (warp % 2) * 32 * LD + (lane >> 2) * LD + (lane & 3);
    const double *B = sm + 64 * LD + (warp % 2) * 32 * LD + (lane >> 2) * LD + (lane & 3);
    double c0[TA][TB], c1[TA][TB];
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 5
        for (int k = 0; k < 100; k += 4) {
            double a[TA], b[TB];
#pragma unroll
            for (int i = 0; i < TA; ++i) a[i] = A[i * 8 * LD + k];
#pragma unroll
            for (int j = 0; j < TB; ++j) b[j] = B[j * 8 * LD + k];
#pragma unroll
            for (int i = 0; i < TA; ++i)
#pragma unroll
                for (int j = 0; j < TB; ++j) dmma884(c0[i][j], c1[i][j], a[i], b[j]);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) r += c0[i][j] + c1[i][j];
    if (r == 123.456) out[0] = r;
}

__global__ void k_copy(const double4 *__restrict__ a, double4 *__restrict__ b, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) b[i] = a[i];
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, 1024));
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d", p.name, sms, p.clockRate);
    const int iters = 20000;
    // DFMA: threads x ILP x iters x 2 flop
    for (int tpb : {128, 256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        float ms = time_ms([&] { k_dfma<8><<<blocks, tpb>>>(out, iters, 1e-9); }, 5);
        double fl = 2.0 * blocks * tpb * 8.0 * iters;
        printf(",\"dfma_tflops_tpb%d\":%.3f", tpb, fl / ms * 1e-9);
    }
    // DMMA: warps x NT x iters x 512 flop
    for (int tpb : {128, 256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        float ms = time_ms([&] { k_dmma<8><<<blocks, tpb>>>(out, iters, 1e-9); }, 5);
        double fl = 512.0 * blocks * (tpb / 32) * 8.0 * iters;
        printf(",\"dmma_tflops_tpb%d\":%.3f", tpb, fl / ms * 1e-9);
    }
    {   // low occupancy DMMA: 4 / 8 warps per SM, with 16 independent accumulators
        for (int tpb : {128, 256}) {
            float ms = time_ms([&] { k_dmma<16><<<sms, tpb>>>(out, iters, 1e-9); }, 5);
            double fl = 512.0 * sms * (tpb / 32) * 16.0 * iters;
            printf(",\"dmma_tflops_1cta_tpb%d_nt16\":%.3f", tpb, fl / ms * 1e-9);
        }
    }
    {
        size_t smem = 64 * 108 * 2 * sizeof(double);
        CK(cudaFuncSetAttribute(k_dmma_smem<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_dmma_smem<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_dmma_smem<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_dmma_smem<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int it2 = 400;
        for (int tpb : {256, 512}) {
            int blocks = sms * 2;
            float ms;
            double base = 512.0 * blocks * (tpb / 32) * 25.0 * it2;
            ms = time_ms([&] { k_dmma_smem<2, 2><<<blocks, tpb, smem>>>(out, it2, 1e-9); }, 5);
            printf(",\"dmma_smem_2x2_tpb%d\":%.3f", tpb, base * 4 / ms * 1e-9);
            ms = time_ms([&] { k_dmma_smem<4, 2><<<blocks, tpb, smem>>>(out, it2, 1e-9); }, 5);
            printf(",\"dmma_smem_4x2_tpb%d\":%.3f", tpb, base * 8 / ms * 1e-9);
            ms = time_ms([&] { k_dmma_smem<4, 4><<<blocks, tpb, smem>>>(out, it2, 1e-9); }, 5);
            printf(",\"dmma_smem_4x4_tpb%d\":%.3f", tpb, base * 16 / ms * 1e-9);
            ms = time_ms([&] { k_dmma_smem<2, 8><<<blocks, tpb, smem>>>(out, it2, 1e-9); }, 5);
            printf(",\"dmma_smem_2x8_tpb%d\":%.3f", tpb, base * 16 / ms * 1e-9);
        }
    }
    {
        size_t n = (size_t)1 << 30;  // 1 GiB each way
        double4 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
        CK(cudaMemset(a, 1, n));
        float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n / sizeof(double4)); }, 5);
        printf(",\"copy_gbs\":%.1f", 2.0 * n / ms * 1e-6);
    }
    printf("}\n");
    return 0;
}
