// C-ABI layer for the batched pyramidal KLT tracker (include/ekfvio_c.h, klt section).
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "klt_common.cuh"
#include "klt_kernels.h"

namespace ekfvio {
extern thread_local std::string g_last_error;
int fail(const char* what, cudaError_t e);
int fail_msg(const std::string& msg);
}  // namespace ekfvio

using namespace kltdev;
using ekfvio::fail_msg;

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ekfvio::fail(#x, e_); } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- TMA descriptors -------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is looked up at run time so that the library links against the
// runtime only.  An image batch is a 3-D tensor of 32-bit words: (pitch / 4, height, images), strides (pitch, image stride).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}
static bool tma_eligible(const void* base, int pitch, size_t img_stride) {
    return encode_tiled_fn() && (((size_t)base) & 15) == 0 && (pitch & 15) == 0 && (img_stride & 15) == 0 && pitch >= 16;
}
static bool make_tensor_map(CUtensorMap* out, const void* base, int pitch, int h, int nimg, size_t img_stride, int box_words, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || box_words > 256 || box_rows > 256 || box_words < 1 || box_rows < 1) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)(pitch / 4), (cuuint64_t)h, (cuuint64_t)(nimg > 0 ? nimg : 1)};
    const cuuint64_t gstride[2] = {(cuuint64_t)pitch, (cuuint64_t)img_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_words, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int staged_rows(int h) { return (h + 7) / 8 * 8 + 4; }      // klt_levels_fused_kernel: a thread's 8-row block may overhang the image
constexpr int kHaloX = 16;                                          // klt_kernels.cu HX: boxes start on 16-byte boundaries
constexpr int kTileWords = (256 + 2 * kHaloX) / 4, kTileRows = 64 + 4;   // klt_level0_tma_kernel's staged tile (klt_kernels.cu ATW / ATH)

extern "C" {

void ekfvio_klt_default_params(ekfvio_klt_params* p) {
    p->window_size = 21;        // Params.h:104
    p->max_pyramid_level = 3;   // Params.h:103
    p->max_iterations = 30;     // KLTTracker.cpp:63
    p->epsilon = 0.01;          // KLTTracker.cpp:63
    p->min_eigen = 1e-4;        // Params.h:36
    p->kill_pad = 11;           // Params.h:33
    p->use_initial_flow = 1;    // KLTTracker.cpp:64
}

int ekfvio_klt_destroy(ekfvio_klt* k) {
    if (!k) return 0;
    cudaSetDevice(k->device);
    cudaFree(k->d_slots); cudaFree(k->d_prev_pts); cudaFree(k->d_next_pts); cudaFree(k->d_status); cudaFree(k->d_err); cudaFree(k->d_npts);
    cudaFreeHost(k->h_img); cudaFreeHost(k->h_pts);
    if (k->copy_st) cudaStreamDestroy(k->copy_st);
    for (int i = 0; i < 4; ++i) if (k->ev_chunk[i]) cudaEventDestroy(k->ev_chunk[i]);
    delete[] k->slot_has_derivs; delete[] k->slot_batch; delete[] k->slot_ext;
    delete[] k->tmap_l0; delete[] k->tmap_l1;
    delete k;
    return 0;
}

int ekfvio_klt_create(ekfvio_klt** out, int device, int width, int height, int max_batch, int max_points, int num_slots,
                      const ekfvio_klt_params* params) {
    if (!out || width <= 0 || height <= 0 || max_batch <= 0 || max_points <= 0 || num_slots < 2) return fail_msg("ekfvio_klt_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("ekfvio_klt_create: no CUDA device (this library has no CPU path)");
    CU(cudaSetDevice(device));
    ekfvio_klt* k = new (std::nothrow) ekfvio_klt();
    if (!k) return fail_msg("out of host memory");
    k->device = device; k->width = width; k->height = height; k->max_batch = max_batch; k->max_points = max_points; k->num_slots = num_slots;
    if (params) k->prm = *params; else ekfvio_klt_default_params(&k->prm);
    const int win = k->prm.window_size;
    if (win < 3 || win > 31 || (win & 1) == 0) { delete k; return fail_msg("ekfvio_klt_create: window_size must be odd and within 3..31"); }
    if (k->prm.max_pyramid_level < 0 || k->prm.max_pyramid_level >= KLT_MAX_LEVELS) { delete k; return fail_msg("ekfvio_klt_create: max_pyramid_level out of range"); }
    // level sizes as cv::buildOpticalFlowPyramid: stop when the next level would be <= window
    Pyr& P = k->pyr;
    int w = width, h = height, lv = 0;
    size_t off = 0;
    for (;;) {
        Level& L = P.lv[lv];
        L.w = w; L.h = h;
        L.pitch = (int)align_up((size_t)w, 16);
        L.dpitch = (int)align_up((size_t)w, 4);
        L.img_stride = align_up((size_t)L.pitch * h, 256);
        L.der_stride = align_up((size_t)L.dpitch * h * sizeof(short2), 256);
        L.img_off = off; off += L.img_stride * max_batch;
        L.der_off = off; off += L.der_stride * max_batch;
        ++lv;
        if (lv > k->prm.max_pyramid_level) break;
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win || nh <= win) break;
        w = nw; h = nh;
    }
    P.levels = lv;
    k->slot_bytes = align_up(off, 256);
    k->slot_has_derivs = new bool[num_slots]();
    k->slot_batch = new int[num_slots]();
    k->slot_ext = new ExtLevel0[num_slots]();
    size_t npt = (size_t)max_batch * max_points;
    cudaError_t e = cudaMalloc((void**)&k->d_slots, k->slot_bytes * num_slots);
    if (e == cudaSuccess) e = cudaMalloc((void**)&k->d_prev_pts, npt * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&k->d_next_pts, npt * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&k->d_status, npt);
    if (e == cudaSuccess) e = cudaMalloc((void**)&k->d_err, npt * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&k->d_npts, max_batch * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&k->h_img, 2 * P.lv[0].img_stride * max_batch);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&k->h_pts, npt * 6 * sizeof(float) + max_batch * sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&k->copy_st, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&k->ev_chunk[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { ekfvio_klt_destroy(k); return ekfvio::fail("ekfvio_klt_create alloc", e); }
    size_t smem = track_smem_bytes(win);
    (void)smem;
    // TMA descriptors of the slots' own level-0 / level-1 images, and the plan for the fused levels 1..3
    k->tmap_l0 = new unsigned long long[num_slots][16]();
    k->tmap_l1 = new unsigned long long[num_slots][16]();
    static_assert(sizeof(CUtensorMap) == 16 * sizeof(unsigned long long), "CUtensorMap is 128 bytes");
    k->tma_ok = encode_tiled_fn() != nullptr;
    if (getenv("EKFVIO_KLT_NO_TMA") == nullptr) {
        for (int s = 0; s < num_slots && k->tma_ok; ++s) {
            uint8_t* base = k->d_slots + (size_t)s * k->slot_bytes;
            k->tma_ok = make_tensor_map(reinterpret_cast<CUtensorMap*>(k->tmap_l0[s]), base + P.lv[0].img_off, P.lv[0].pitch, P.lv[0].h, max_batch,
                                        P.lv[0].img_stride, kTileWords, kTileRows);
        }
        if (k->tma_ok && P.levels >= 2 && P.levels <= 4) {
            const Level& L1 = P.lv[1];
            // staged levels in shared memory: level 1 | level 2, level 3 over the (then dead) level 1
            size_t off = 0;
            for (int l = 1; l < P.levels && l <= 2; ++l) off += align_up((size_t)(P.lv[l].pitch + 2 * kHaloX) * staged_rows(P.lv[l].h), 128);
            bool ok = (L1.pitch + 2 * kHaloX) / 4 <= 256 && staged_rows(L1.h) <= 256 && off <= 200 * 1024;
            if (P.levels == 4 && (size_t)(P.lv[3].pitch + 2 * kHaloX) * staged_rows(P.lv[3].h) > (size_t)(L1.pitch + 2 * kHaloX) * staged_rows(L1.h)) ok = false;
            for (int s = 0; s < num_slots && ok; ++s) {
                uint8_t* base = k->d_slots + (size_t)s * k->slot_bytes;
                ok = make_tensor_map(reinterpret_cast<CUtensorMap*>(k->tmap_l1[s]), base + L1.img_off, L1.pitch, L1.h, max_batch, L1.img_stride,
                                     (L1.pitch + 2 * kHaloX) / 4, staged_rows(L1.h));
            }
            if (ok) { k->fused_levels = P.levels - 1; k->fused_smem = off + 128; }      // (+128: the kernel aligns its base)
        }
    } else k->tma_ok = false;
    *out = k;
    return 0;
}

int ekfvio_klt_enable_timing(ekfvio_klt* k, int on) {
    CU(cudaSetDevice(k->device));
    k->timer.reset();
    k->timer.on = on != 0;
    return 0;
}

int ekfvio_klt_get_timing(ekfvio_klt* k, double* ms8, long long* count8) {
    CU(cudaSetDevice(k->device));
    k->timer.resolve();
    for (int i = 0; i < KernelTimer::SLOTS; ++i) { if (ms8) ms8[i] = k->timer.ms[i]; if (count8) count8[i] = k->timer.cnt[i]; }
    return 0;
}

int ekfvio_klt_num_levels(const ekfvio_klt* k) { return k ? k->pyr.levels : 0; }
long long ekfvio_klt_launch_count(const ekfvio_klt* k) { return k ? k->launches : 0; }

// Builds one or two slots level by level; both slots share each level's launch.
static int build_slots(ekfvio_klt* k, int nslots, const int* slots, const uint8_t* const* imgs, int pitch, int batch, const int* with_derivs,
                       cudaStream_t st, int first = 0, int total = -1, bool by_ref = false) {   // images [first, first + batch) of a batch of `total`
    const Pyr& P = k->pyr;
    for (int l = 0; l < P.levels; ++l) {
        const Level& L = P.lv[l];
        if (l == 1 && k->fused_levels > 0) {
            // levels 1 .. fused_levels of every image in one launch: level 1 arrives by TMA and stays in shared memory
            FusedJob fj[2] = {FusedJob{nullptr, 0, 0, 0}, FusedJob{nullptr, 0, 0, 0}};
            CUtensorMap maps[2];
            for (int s = 0; s < 2; ++s) {
                const int sl = s < nslots ? slots[s] : slots[0];
                memcpy(&maps[s], k->tmap_l1[sl], sizeof(CUtensorMap));
                if (s < nslots) fj[s] = FusedJob{k->d_slots + (size_t)slots[s] * k->slot_bytes, with_derivs[s] ? 1 : 0, batch, first};
            }
            FusedLevel fl[3];
            size_t off = 0;
            for (int i = 0; i < k->fused_levels; ++i) {
                const Level& Li = P.lv[1 + i];
                fl[i] = FusedLevel{Li.w, Li.h, Li.pitch, Li.dpitch, Li.pitch + 2 * kHaloX, staged_rows(Li.h), i == 2 ? 0 : (int)off, Li.img_off, Li.der_off, Li.img_stride, Li.der_stride};
                off += align_up((size_t)(Li.pitch + 2 * kHaloX) * staged_rows(Li.h), 128);
            }
            k->timer.begin(1, st);
            CU(launch_levels_fused(maps[0], maps[1], fj[0], fj[1], fl, k->fused_levels, k->fused_smem, st));
            k->timer.end(st);
            k->launches += 1;
            break;
        }
        LevelJob jobs[2];
        bool any = false;
        for (int s = 0; s < 2; ++s) {
            LevelJob& J = jobs[s];
            J = LevelJob{nullptr, 0, 0, nullptr, nullptr, nullptr, 0};
            if (s >= nslots) continue;
            uint8_t* base = k->d_slots + (size_t)slots[s] * k->slot_bytes;
            J.batch = batch;
            if (l == 0 && imgs[s]) {
                J.src = imgs[s] + (size_t)first * pitch * k->height; J.spitch = pitch; J.sstride = (size_t)pitch * k->height;
                if (!by_ref) J.copy_dst = base + L.img_off + (size_t)first * L.img_stride;     // (by reference: level 0 stays where the caller has it)
            } else { J.src = base + L.img_off + (size_t)first * L.img_stride; J.spitch = L.pitch; J.sstride = L.img_stride; }
            if (with_derivs[s]) J.deriv = reinterpret_cast<short2*>(base + L.der_off + (size_t)first * L.der_stride);
            if (l + 1 < P.levels) J.down = base + P.lv[l + 1].img_off + (size_t)first * P.lv[l + 1].img_stride;
            any = any || J.copy_dst || J.deriv || J.down;
        }
        if (!any) continue;
        int npitch = 0; size_t nstride = 0;
        if (l + 1 < P.levels) { npitch = P.lv[l + 1].pitch; nstride = P.lv[l + 1].img_stride; }
        k->timer.begin(l < 3 ? l : 3, st);
        bool done = false;
        if (l == 0 && k->tma_ok) {
            // level 0 staged by TMA: the source is either the caller's image batch (a descriptor is encoded for this call) or the
            // slot's own level 0 (descriptors made at create time).  Sources TMA cannot address (unaligned base / pitch) take the
            // cp.async kernel below.
            CUtensorMap maps[2];
            TmaJob tj[2] = {TmaJob{nullptr, nullptr, nullptr, 0, 0}, TmaJob{nullptr, nullptr, nullptr, 0, 0}};
            bool ok = true;
            for (int s = 0; s < 2 && ok; ++s) {
                const int sl = s < nslots ? slots[s] : slots[0];
                if (s < nslots && imgs[s]) {
                    const uint8_t* src = imgs[s] + (size_t)first * pitch * k->height;
                    ok = tma_eligible(src, pitch, (size_t)pitch * k->height) && make_tensor_map(&maps[s], src, pitch, k->height, batch, (size_t)pitch * k->height, kTileWords, kTileRows);
                    if (ok) tj[s] = TmaJob{jobs[s].copy_dst, jobs[s].deriv, jobs[s].down, batch, 0};
                } else {
                    memcpy(&maps[s], k->tmap_l0[sl], sizeof(CUtensorMap));
                    if (s < nslots) tj[s] = TmaJob{nullptr, jobs[s].deriv, jobs[s].down, batch, first};
                }
            }
            if (ok) {
                CU(launch_level0_tma(maps[0], maps[1], tj[0], tj[1], L.w, L.h, L.pitch, L.img_stride, L.dpitch, L.der_stride / sizeof(short2), npitch, nstride, st));
                done = true;
            }
        }
        if (!done) CU(launch_level(jobs[0], jobs[1], L.w, L.h, L.pitch, L.img_stride, L.dpitch, L.der_stride / sizeof(short2), npitch, nstride, st));
        k->timer.end(st);
        k->launches += 1;
    }
    for (int s = 0; s < nslots; ++s) {
        k->slot_has_derivs[slots[s]] = with_derivs[s] != 0; k->slot_batch[slots[s]] = total < 0 ? batch : total;
        k->slot_ext[slots[s]] = (by_ref && imgs[s]) ? ExtLevel0{imgs[s], pitch, (size_t)pitch * k->height} : ExtLevel0{nullptr, 0, 0};
    }
    return 0;
}

int ekfvio_klt_build_pyramid(ekfvio_klt* k, int slot, const uint8_t* d_imgs, int pitch, int batch, int with_derivs, void* stream) {
    if (slot < 0 || slot >= k->num_slots || batch <= 0 || batch > k->max_batch) return fail_msg("ekfvio_klt_build_pyramid: bad slot or batch");
    CU(cudaSetDevice(k->device));
    return build_slots(k, 1, &slot, &d_imgs, pitch, batch, &with_derivs, (cudaStream_t)stream);
}

int ekfvio_klt_build_pyramid_pair(ekfvio_klt* k, int prev_slot, const uint8_t* d_prev, int next_slot, const uint8_t* d_next, int pitch, int batch,
                                  int next_with_derivs, void* stream) {
    if (prev_slot < 0 || prev_slot >= k->num_slots || next_slot < 0 || next_slot >= k->num_slots || prev_slot == next_slot || batch <= 0 ||
        batch > k->max_batch)
        return fail_msg("ekfvio_klt_build_pyramid_pair: bad slots or batch");
    CU(cudaSetDevice(k->device));
    const int slots[2] = {prev_slot, next_slot};
    const uint8_t* imgs[2] = {d_prev, d_next};
    const int wd[2] = {1, next_with_derivs};
    return build_slots(k, 2, slots, imgs, pitch, batch, wd, (cudaStream_t)stream);
}

int ekfvio_klt_build_pyramid_pair_ref(ekfvio_klt* k, int prev_slot, const uint8_t* d_prev, int next_slot, const uint8_t* d_next, int pitch, int batch,
                                      int next_with_derivs, void* stream) {
    if (prev_slot < 0 || prev_slot >= k->num_slots || next_slot < 0 || next_slot >= k->num_slots || prev_slot == next_slot || batch <= 0 ||
        batch > k->max_batch || !d_prev || !d_next || pitch < k->width)
        return fail_msg("ekfvio_klt_build_pyramid_pair_ref: bad slots, batch or images");
    CU(cudaSetDevice(k->device));
    const int slots[2] = {prev_slot, next_slot};
    const uint8_t* imgs[2] = {d_prev, d_next};
    const int wd[2] = {1, next_with_derivs};
    return build_slots(k, 2, slots, imgs, pitch, batch, wd, (cudaStream_t)stream, 0, -1, true);
}

int ekfvio_klt_track(ekfvio_klt* k, int prev_slot, int next_slot, const float* d_prev_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                     const int* d_npts, int batch, void* stream) {
    if (prev_slot < 0 || prev_slot >= k->num_slots || next_slot < 0 || next_slot >= k->num_slots) return fail_msg("ekfvio_klt_track: bad slot");
    if (!k->slot_has_derivs[prev_slot]) return fail_msg("ekfvio_klt_track: prev slot was built without derivatives");
    if (batch <= 0 || batch > k->slot_batch[prev_slot] || batch > k->slot_batch[next_slot]) return fail_msg("ekfvio_klt_track: batch exceeds the built pyramids");
    CU(cudaSetDevice(k->device));
    k->timer.begin(4, (cudaStream_t)stream);
    CU(launch_track(k->pyr, k->d_slots + (size_t)prev_slot * k->slot_bytes, k->d_slots + (size_t)next_slot * k->slot_bytes, d_prev_pts, d_next_pts,
                    d_status, d_err, d_npts, k->max_points, 0, batch, k->prm, (cudaStream_t)stream, k->slot_ext[prev_slot], k->slot_ext[next_slot]));
    k->timer.end((cudaStream_t)stream);
    k->launches += 1;
    return 0;
}

int ekfvio_klt_postprocess(ekfvio_klt* k, const float* d_next_pts, const uint8_t* d_status, const int* d_npts, const float* d_K9, int batch,
                           float* d_measured, float* d_cov, uint8_t* d_passed, void* stream) {
    CU(cudaSetDevice(k->device));
    CU(launch_postprocess(d_next_pts, d_status, d_npts, d_K9, k->max_points, batch, k->width, k->height, k->prm.kill_pad, d_measured, d_cov,
                          d_passed, (cudaStream_t)stream));
    k->launches += 1;
    return 0;
}

int ekfvio_klt_track_pair_h(ekfvio_klt* k, const uint8_t* h_prev, const uint8_t* h_next, int pitch, int batch, const float* h_prev_pts,
                            float* h_next_pts, uint8_t* h_status, float* h_err, const int* h_npts, void* stream) {
    if (batch <= 0 || batch > k->max_batch) return fail_msg("ekfvio_klt_track_pair_h: bad batch");
    CU(cudaSetDevice(k->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Level& L0 = k->pyr.lv[0];
    uint8_t* s0 = k->d_slots;                   // slot 0 <- prev
    uint8_t* s1 = k->d_slots + k->slot_bytes;   // slot 1 <- next
    size_t npt = (size_t)batch * k->max_points;
    float* hp = k->h_pts;
    memcpy(hp, h_prev_pts, npt * 2 * sizeof(float));
    memcpy(hp + npt * 2, h_next_pts, npt * 2 * sizeof(float));
    int* hn = reinterpret_cast<int*>(hp + (size_t)k->max_batch * k->max_points * 6);
    memcpy(hn, h_npts, batch * sizeof(int));
    CU(cudaMemcpyAsync(k->d_prev_pts, hp, npt * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(k->d_next_pts, hp + npt * 2, npt * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(k->d_npts, hn, batch * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(k->d_status, 0, npt, st));
    CU(cudaMemsetAsync(k->d_err, 0, npt * sizeof(float), st));
    // Frames go straight into level 0 of the slots (no staging copy on the device).  Page-locked caller
    // buffers are DMA'd from where they are; pageable ones go through the tracker's pinned staging area.
    // The batch is cut into chunks: a copy stream uploads chunk c+1 while the caller's stream builds the
    // pyramids of chunk c and tracks its points, so the PCIe transfer hides the kernels.
    cudaPointerAttributes attr;
    bool pinned[2];
    for (int which = 0; which < 2; ++which) {
        pinned[which] = cudaPointerGetAttributes(&attr, which ? h_next : h_prev) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned[which]) cudaGetLastError();
    }
    const int nchunks = batch >= 32 ? 4 : 1, per = (batch + nchunks - 1) / nchunks;
    CU(cudaEventRecord(k->ev_chunk[0], st));                  // the slots may still be read by earlier work on the caller's stream
    CU(cudaStreamWaitEvent(k->copy_st, k->ev_chunk[0], 0));
    const int slots[2] = {0, 1};
    const uint8_t* none[2] = {nullptr, nullptr};
    const int wd[2] = {1, 0};
    for (int c = 0, first = 0; first < batch; ++c, first += per) {
        const int nb = batch - first < per ? batch - first : per;
        for (int which = 0; which < 2; ++which) {
            const uint8_t* h = (which ? h_next : h_prev) + (size_t)first * k->height * pitch;
            uint8_t* dst = (which ? s1 : s0) + L0.img_off + (size_t)first * L0.img_stride;
            if (pinned[which] && L0.img_stride == (size_t)L0.pitch * k->height) {
                CU(cudaMemcpy2DAsync(dst, L0.pitch, h, pitch, k->width, (size_t)k->height * nb, cudaMemcpyHostToDevice, k->copy_st));
            } else if (pinned[which]) {
                for (int b = 0; b < nb; ++b)
                    CU(cudaMemcpy2DAsync(dst + (size_t)b * L0.img_stride, L0.pitch, h + (size_t)b * k->height * pitch, pitch, k->width, k->height,
                                         cudaMemcpyHostToDevice, k->copy_st));
            } else {
                uint8_t* stage = k->h_img + (size_t)which * L0.img_stride * k->max_batch + (size_t)first * L0.img_stride;
                for (int b = 0; b < nb; ++b)
                    for (int y = 0; y < k->height; ++y)
                        memcpy(stage + (size_t)b * L0.img_stride + (size_t)y * L0.pitch, h + ((size_t)b * k->height + y) * pitch, (size_t)k->width);
                CU(cudaMemcpyAsync(dst, stage, L0.img_stride * nb, cudaMemcpyHostToDevice, k->copy_st));
            }
        }
        CU(cudaEventRecord(k->ev_chunk[c & 3], k->copy_st));
        CU(cudaStreamWaitEvent(st, k->ev_chunk[c & 3], 0));
        int rc = build_slots(k, 2, slots, none, 0, nb, wd, st, first, batch);
        if (rc) return rc;
        k->timer.begin(4, st);
        CU(launch_track(k->pyr, s0, s1, k->d_prev_pts, k->d_next_pts, k->d_status, k->d_err, k->d_npts, k->max_points, first, nb, k->prm, st));
        k->timer.end(st);
        k->launches += 1;
    }
    float* ho = hp + npt * 2;  // reuse: next pts | err | status
    CU(cudaMemcpyAsync(ho, k->d_next_pts, npt * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho + npt * 2, k->d_err, npt * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho + npt * 3, k->d_status, npt, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(h_next_pts, ho, npt * 2 * sizeof(float));
    if (h_err) memcpy(h_err, ho + npt * 2, npt * sizeof(float));
    memcpy(h_status, ho + npt * 3, npt);
    k->seq_prev = 1;                            // slot 1 now holds the latest frame (its derivatives are built on demand)
    return 0;
}

// The tracker inside a sequence (KLTTracker as EKFVIO::addFrame drives it, EKFVIO.cpp:201-217): the previous frame's pyramid is
// already on the device from the last call, so only the new frame is uploaded; it is built with derivatives because it is the
// "previous" frame of the next call.  Same chunked upload / build / track overlap as ekfvio_klt_track_pair_h.
int ekfvio_klt_track_next_h(ekfvio_klt* k, const uint8_t* h_next, int pitch, int batch, const float* h_prev_pts, float* h_next_pts,
                            uint8_t* h_status, float* h_err, const int* h_npts, void* stream) {
    if (k->seq_prev < 0) return fail_msg("ekfvio_klt_track_next_h: no previous frame on the device (call ekfvio_klt_track_pair_h first)");
    const int prev_slot = k->seq_prev, next_slot = prev_slot == 0 ? 1 : 0;
    if (batch <= 0 || batch > k->max_batch || batch > k->slot_batch[prev_slot]) return fail_msg("ekfvio_klt_track_next_h: batch does not match the previous frame's");
    CU(cudaSetDevice(k->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Level& L0 = k->pyr.lv[0];
    uint8_t* sp = k->d_slots + (size_t)prev_slot * k->slot_bytes;
    uint8_t* sn = k->d_slots + (size_t)next_slot * k->slot_bytes;
    size_t npt = (size_t)batch * k->max_points;
    float* hp = k->h_pts;
    memcpy(hp, h_prev_pts, npt * 2 * sizeof(float));
    memcpy(hp + npt * 2, h_next_pts, npt * 2 * sizeof(float));
    int* hn = reinterpret_cast<int*>(hp + (size_t)k->max_batch * k->max_points * 6);
    memcpy(hn, h_npts, batch * sizeof(int));
    CU(cudaMemcpyAsync(k->d_prev_pts, hp, npt * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(k->d_next_pts, hp + npt * 2, npt * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(k->d_npts, hn, batch * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(k->d_status, 0, npt, st));
    CU(cudaMemsetAsync(k->d_err, 0, npt * sizeof(float), st));
    const uint8_t* none[1] = {nullptr};
    const int wd[1] = {1};
    if (!k->slot_has_derivs[prev_slot]) {       // the previous frame came in as the "next" of a pair call: add its derivatives
        int rc = build_slots(k, 1, &prev_slot, none, 0, batch, wd, st);
        if (rc) return rc;
    }
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, h_next) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    const int nchunks = batch >= 32 ? 4 : 1, per = (batch + nchunks - 1) / nchunks;
    CU(cudaEventRecord(k->ev_chunk[0], st));
    CU(cudaStreamWaitEvent(k->copy_st, k->ev_chunk[0], 0));
    for (int c = 0, first = 0; first < batch; ++c, first += per) {
        const int nb = batch - first < per ? batch - first : per;
        const uint8_t* h = h_next + (size_t)first * k->height * pitch;
        uint8_t* dst = sn + L0.img_off + (size_t)first * L0.img_stride;
        if (pinned && L0.img_stride == (size_t)L0.pitch * k->height) {
            CU(cudaMemcpy2DAsync(dst, L0.pitch, h, pitch, k->width, (size_t)k->height * nb, cudaMemcpyHostToDevice, k->copy_st));
        } else if (pinned) {
            for (int b = 0; b < nb; ++b)
                CU(cudaMemcpy2DAsync(dst + (size_t)b * L0.img_stride, L0.pitch, h + (size_t)b * k->height * pitch, pitch, k->width, k->height,
                                     cudaMemcpyHostToDevice, k->copy_st));
        } else {
            uint8_t* stage = k->h_img + (size_t)first * L0.img_stride;
            for (int b = 0; b < nb; ++b)
                for (int y = 0; y < k->height; ++y)
                    memcpy(stage + (size_t)b * L0.img_stride + (size_t)y * L0.pitch, h + ((size_t)b * k->height + y) * pitch, (size_t)k->width);
            CU(cudaMemcpyAsync(dst, stage, L0.img_stride * nb, cudaMemcpyHostToDevice, k->copy_st));
        }
        CU(cudaEventRecord(k->ev_chunk[c & 3], k->copy_st));
        CU(cudaStreamWaitEvent(st, k->ev_chunk[c & 3], 0));
        int rc = build_slots(k, 1, &next_slot, none, 0, nb, wd, st, first, batch);
        if (rc) return rc;
        k->timer.begin(4, st);
        CU(launch_track(k->pyr, sp, sn, k->d_prev_pts, k->d_next_pts, k->d_status, k->d_err, k->d_npts, k->max_points, first, nb, k->prm, st));
        k->timer.end(st);
        k->launches += 1;
    }
    float* ho = hp + npt * 2;
    CU(cudaMemcpyAsync(ho, k->d_next_pts, npt * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho + npt * 2, k->d_err, npt * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho + npt * 3, k->d_status, npt, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(h_next_pts, ho, npt * 2 * sizeof(float));
    if (h_err) memcpy(h_err, ho + npt * 2, npt * sizeof(float));
    memcpy(h_status, ho + npt * 3, npt);
    k->seq_prev = next_slot;
    return 0;
}

// KLTTracker::estimateUncertaintySampleBased (KLTTracker.cpp:111-175) for a batch of frame pairs; device pointers.
int ekfvio_klt_sample_uncertainty(int device, const uint8_t* d_ref_imgs, const uint8_t* d_cur_imgs, int width, int height, int pitch, int batch,
                                  const float* d_ref_pts, const float* d_pts, const int* d_npts, int max_points, float* d_cov, void* stream) {
    if (!d_ref_imgs || !d_cur_imgs || !d_ref_pts || !d_pts || !d_npts || !d_cov || width <= 0 || height <= 0 || pitch < width || batch <= 0 || max_points <= 0)
        return fail_msg("ekfvio_klt_sample_uncertainty: bad arguments");
    CU(cudaSetDevice(device));
    CU(launch_sample_uncertainty(d_ref_imgs, d_cur_imgs, width, height, pitch, (size_t)pitch * height, batch, d_ref_pts, d_pts, d_npts, max_points, d_cov,
                                 (cudaStream_t)stream));
    return 0;
}

// The same for one frame pair in host memory (what the facade's single-feature call uses).
int ekfvio_klt_sample_uncertainty_h(int device, const uint8_t* h_ref_img, const uint8_t* h_cur_img, int width, int height, int pitch, const float* h_ref_pts,
                                    const float* h_pts, int n, float* h_cov) {
    if (!h_ref_img || !h_cur_img || !h_ref_pts || !h_pts || !h_cov || n <= 0) return fail_msg("ekfvio_klt_sample_uncertainty_h: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("ekfvio_klt_sample_uncertainty_h: no CUDA device (this library has no CPU path)");
    CU(cudaSetDevice(device));
    const size_t ib = (size_t)pitch * height;
    uint8_t* d_img = nullptr; float* d_f = nullptr; int* d_n = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_img, 2 * ib);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_f, (size_t)n * 8 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_n, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(d_img, h_ref_img, ib, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_img + ib, h_cur_img, ib, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f, h_ref_pts, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f + 2 * n, h_pts, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_n, &n, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_sample_uncertainty(d_img, d_img + ib, width, height, pitch, ib, 1, d_f, d_f + 2 * n, d_n, n, d_f + 4 * n, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(h_cov, d_f + 4 * n, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_img); cudaFree(d_f); cudaFree(d_n);
    if (e != cudaSuccess) return ekfvio::fail("ekfvio_klt_sample_uncertainty_h", e);
    return 0;
}

int ekfvio_klt_read_level(ekfvio_klt* k, int slot, int img, int level, uint8_t* h_img, int16_t* h_deriv, int* w_out, int* h_out) {
    if (slot < 0 || slot >= k->num_slots || level < 0 || level >= k->pyr.levels || img < 0 || img >= k->max_batch) return fail_msg("ekfvio_klt_read_level: bad index");
    CU(cudaSetDevice(k->device));
    const Level& L = k->pyr.lv[level];
    if (w_out) *w_out = L.w;
    if (h_out) *h_out = L.h;
    CU(cudaDeviceSynchronize());
    uint8_t* base = k->d_slots + (size_t)slot * k->slot_bytes;
    if (h_img) {
        const ExtLevel0& E = k->slot_ext[slot];
        if (level == 0 && E.img) CU(cudaMemcpy2D(h_img, L.w, E.img + (size_t)img * E.stride, E.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
        else CU(cudaMemcpy2D(h_img, L.w, base + L.img_off + (size_t)img * L.img_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    }
    if (h_deriv) {
        if (!k->slot_has_derivs[slot]) return fail_msg("ekfvio_klt_read_level: slot has no derivatives");
        CU(cudaMemcpy2D(h_deriv, (size_t)L.w * 4, base + L.der_off + (size_t)img * L.der_stride, (size_t)L.dpitch * 4, (size_t)L.w * 4, L.h, cudaMemcpyDeviceToHost));
    }
    return 0;
}

}  // extern "C"
