// HBM streaming probe: bandwidth of pure-read, pure-write, copy and a 1:5 read:write mix (the
// traffic shape of the KLT pyramid level-0 launch).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(uint4* d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = make_uint4(i, 1, 2, 3);
}
__global__ void k_read(const uint4* s, size_t n, uint4* sink) {
    uint4 a = make_uint4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = s[i]; a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w; }
    if (a.x == 0x12345678 && a.y == 77) sink[0] = a;
}
__global__ void k_copy(const uint4* s, uint4* d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// read n/5 vectors, write n vectors
__global__ void k_mix(const uint4* s, uint4* d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 5; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = s[i];
#pragma unroll
        for (int k = 0; k < 5; ++k) { d[i + k * (n / 5)] = v; v.x += 1; }
    }
}
int main() {
    const size_t bytes = 2ull << 30, n = bytes / 16;
    uint4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 16, blk = 256;
    auto time = [&](const char* name, auto launch, double moved) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        const int reps = 10;
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-10s %8.1f GB/s\n", name, moved * reps / (ms * 1e-3) / 1e9);
    };
    time("write", [&] { k_write<<<grid, blk>>>(a, n); }, (double)bytes);
    time("read", [&] { k_read<<<grid, blk>>>(a, n, b); }, (double)bytes);
    time("copy", [&] { k_copy<<<grid, blk>>>(a, b, n); }, 2.0 * bytes);
    time("mix 1r:5w", [&] { k_mix<<<grid, blk>>>(a, b, n); }, 1.2 * bytes);
    time("memset", [&] { cudaMemsetAsync(a, 0, bytes); }, (double)bytes);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
