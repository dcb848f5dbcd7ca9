// Minimal TMA probe, one variant per process (CUDA errors are sticky):  tma_min <variant>
//  0: mbarrier only (init, expect_tx 0, arrive, wait)      1: 3-D u32 box via __grid_constant__ descriptor
//  2: 2-D u8 box via __grid_constant__ descriptor           3: 3-D u32, descriptor read from global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(unsigned long long* bar) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__global__ void k_mbar(unsigned* out) {
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    wait0(&bar);
    out[threadIdx.x] = threadIdx.x;
}
template <int RANK>
__global__ void k_tma(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, unsigned* out, int words, int cx, int cy, int cz) {
    __shared__ __align__(128) unsigned tile[4096];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* m = gmap ? gmap : &map;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(words * 4) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)),
                         "l"(m), "r"(smem_u32(&bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)),
                         "l"(m), "r"(smem_u32(&bar)), "r"(cx), "r"(cy) : "memory");
    }
    wait0(&bar);
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int pitch = 640, h = 480, n = 3;
    std::vector<unsigned char> host((size_t)pitch * h * n);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (unsigned char)(i * 7 + (i >> 9));
    unsigned char* d; cudaMalloc(&d, host.size()); cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 4096 * 4);
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    printf("variant %d driver %d runtime %d sizeof(CUtensorMap) %zu alignof %zu\n", variant, drv, rt, sizeof(CUtensorMap), alignof(CUtensorMap));
    if (variant == 0) { k_mbar<<<1, 128>>>(out); printf("mbarrier only -> %s\n", cudaGetErrorString(cudaDeviceSynchronize())); return 0; }
    CUtensorMap m; CUresult r;
    cuuint32_t es[3] = {1, 1, 1};
    if (variant == 2) {
        cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)h * n}, gstr[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {64, 8};
        r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2d u8 -> %d\n", (int)r);
        k_tma<2><<<1, 128>>>(m, nullptr, out, 64 * 8 / 4, 16, 4, 0);
    } else {
        cuuint64_t gdim[3] = {(cuuint64_t)pitch / 4, (cuuint64_t)h, (cuuint64_t)n}, gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
        cuuint32_t box[3] = {16, 8, 1};
        r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 3d u32 -> %d\n", (int)r);
        CUtensorMap* gm = nullptr;
        if (variant == 3) { cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice); }
        if (variant == 5) {
            r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("encode with L2 128B promotion -> %d\n", (int)r);
        }
        int cx = 4, cy = 4;
        if (variant == 4) { cx = -2; cy = -2; }
        if (variant == 6) { cx = -2; cy = 4; }
        if (variant == 7) { cx = 4; cy = -2; }
        if (variant == 8) { cx = 156; cy = 476; }
        k_tma<3><<<1, 128>>>(m, gm, out, 16 * 8, cx, cy, 1);
    }
    printf("kernel -> %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<unsigned> res(128); cudaMemcpy(res.data(), out, 512, cudaMemcpyDeviceToHost);
    printf("first words %08x %08x %08x\n", res[0], res[1], res[2]);
    return 0;
}
