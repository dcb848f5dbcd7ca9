// C-ABI layer for the batched EKF (include/ekfvio_c.h).  Owns device memory, picks the kernel
// path, counts launches.  No CPU fallback: every entry point fails if CUDA does.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>

#include "ekf_common.cuh"
#include "ekf_kernels.h"

using namespace ekfvio;

namespace ekfvio {
thread_local std::string g_last_error;
int fail(const char* what, cudaError_t e) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}
int fail_msg(const std::string& msg) { g_last_error = msg; return 1; }
}  // namespace ekfvio

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ekfvio::fail(#x, e_); } while (0)

static EkfPtrs ptrs(const ekfvio_batch* b) {
    EkfPtrs p;
    p.mu = b->d_mu; p.feat = b->d_feat; p.nfeat = b->d_nfeat; p.cache = b->d_cache; p.dflags = b->d_flags;
    p.klt_last = b->d_klt_last; p.status = b->d_status; p.idx = b->d_idx; p.y = b->d_y; p.m = b->d_m; p.K = b->d_K; p.W = b->d_W; p.L = b->d_L; p.asym = b->d_asym; p.route = b->d_route;
    p.F = b->F; p.nmax = b->nmax; p.Nmax = b->Nmax; p.ldP = b->ldP; p.ldK = b->ldK; p.mmax = b->mmax;
    p.flags = b->prm.flags;
    p.fused = 0; p.fb = nullptr;
    p.sigma_lower = b->upper_stale ? 1 : 0;
    p.depth = b->prm.default_point_depth; p.depth_var = b->prm.default_point_depth_variance;
    p.uv_var = b->prm.default_point_homogenous_variance;
    p.gain_smem_doubles = gain_general_smem_doubles(b->mmax);
    p.illcond = b->illcond;
    return p;
}

// Event records that order the private copy stream must not become part of a caller's stream capture
// (the frame loop replays its enqueue sequence as a CUDA graph): while capturing, fall back to plain stream order.
static bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus s = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &s) == cudaSuccess && s != cudaStreamCaptureStatusNone;
}

// Sigma of symmetric filters after a lower-mode process(): mirror the part below the diagonal blocks into the stale part
// above them before anyone but the reduced tiled update reads the matrix.
static int ensure_full_sigma(ekfvio_batch* b, cudaStream_t st) {
    if (!b->upper_stale) return 0;
    cudaError_t e = launch_mirror_lower(ptrs(b), b->d_P[b->cur], st);
    if (e != cudaSuccess) return ekfvio::fail("launch_mirror_lower", e);
    b->upper_stale = false;
    b->launches += 1;
    return 0;
}

extern "C" {

const char* ekfvio_last_error(void) { return ekfvio::g_last_error.c_str(); }

void ekfvio_default_params(ekfvio_params* p) {
    p->default_point_depth = 0.5;
    p->default_point_depth_variance = 100;
    p->default_point_homogenous_variance = 0.00001;
    p->flags = 0;
}

int ekfvio_batch_destroy(ekfvio_batch* b) {
    if (!b) return 0;
    cudaSetDevice(b->device);
    cudaFree(b->d_mu); cudaFree(b->d_feat); cudaFree(b->d_P[0]); cudaFree(b->d_P[1]); cudaFree(b->d_nfeat); cudaFree(b->d_cache);
    cudaFree(b->d_flags); cudaFree(b->d_klt_last); cudaFree(b->d_status); cudaFree(b->d_dt); cudaFree(b->d_K); cudaFree(b->d_W);
    cudaFree(b->d_S); cudaFree(b->d_L); cudaFree(b->d_asym); cudaFree(b->d_route); cudaFree(b->d_fb); cudaFree(b->d_LS); cudaFree(b->d_LL); cudaFree(b->d_LT); cudaFree(b->d_y); cudaFree(b->d_idx); cudaFree(b->d_m); cudaFree(b->d_fjac);
    cudaFree(b->dd_z); cudaFree(b->dd_R); cudaFree(b->dd_pass);
    cudaFreeHost(b->h_z); cudaFreeHost(b->h_R); cudaFreeHost(b->h_pass); cudaFreeHost(b->h_out);
    if (b->copy_st) cudaStreamDestroy(b->copy_st);
    if (b->ev_h2d) cudaEventDestroy(b->ev_h2d);
    if (b->ev_inputs_free) cudaEventDestroy(b->ev_inputs_free);
    if (b->ev_state) cudaEventDestroy(b->ev_state);
    if (b->ev_entry) cudaEventDestroy(b->ev_entry);
    delete b;
    return 0;
}

int ekfvio_batch_create(ekfvio_batch** out, int device, int num_filters, int max_features, const ekfvio_params* params) {
    if (!out || num_filters <= 0 || max_features < 0) return fail_msg("ekfvio_batch_create: bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail_msg("ekfvio_batch_create: no CUDA device (this library has no CPU path)");
    CU(cudaSetDevice(device));
    ekfvio_batch* b = new (std::nothrow) ekfvio_batch();
    if (!b) return fail_msg("out of host memory");
    b->device = device;
    b->F = num_filters; b->nmax = max_features;
    b->Nmax = BASE + 3 * max_features;
    b->mmax = 2 * max_features > 0 ? 2 * max_features : 2;
    b->large = b->Nmax > 176 || b->mmax > 104;                    // beyond the register-resident tiled path
    b->ldP = (b->Nmax + (b->large ? 1 : 0) + 7) / 8 * 8;          // (large path: one spare panel row carries y through the forward substitution)
    b->ldK = b->large ? (b->mmax + 63) / 64 * 64 : (b->mmax + 15) / 16 * 16;    // K / W panels are chunk-major in 16-column chunks (kw_at)
    if (params) b->prm = *params; else ekfvio_default_params(&b->prm);
    b->illcond = ILLCOND_RATIO;
    if (const char* e = getenv("EKFVIO_ILLCOND")) { const double v = atof(e); if (v > 1.0) b->illcond = v; }
    size_t F = b->F, nm = b->nmax > 0 ? b->nmax : 1;
    size_t Pbytes = F * b->ldP * b->ldP * sizeof(double), Kbytes = F * b->ldP * b->ldK * sizeof(double);
#define ALLOC(ptr, bytes) do { cudaError_t e2 = cudaMalloc((void**)&(ptr), (bytes)); if (e2 != cudaSuccess) { ekfvio_batch_destroy(b); return ekfvio::fail("cudaMalloc " #ptr, e2); } cudaMemset((ptr), 0, (bytes)); } while (0)
    ALLOC(b->d_mu, F * BASE * sizeof(double));
    ALLOC(b->d_feat, F * nm * 3 * sizeof(double));
    ALLOC(b->d_P[0], Pbytes);
    ALLOC(b->d_P[1], Pbytes);
    ALLOC(b->d_nfeat, F * sizeof(int));
    ALLOC(b->d_cache, F * 7 * sizeof(double));
    ALLOC(b->d_flags, F * nm);
    ALLOC(b->d_klt_last, F * nm * 2 * sizeof(double));
    ALLOC(b->d_status, F * sizeof(int));
    ALLOC(b->d_dt, F * sizeof(double));
    ALLOC(b->d_K, Kbytes);
    ALLOC(b->d_W, Kbytes);
    ALLOC(b->d_y, F * b->mmax * sizeof(double));
    ALLOC(b->d_idx, F * b->mmax * sizeof(int));
    ALLOC(b->d_m, F * sizeof(int));
    ALLOC(b->d_asym, F * sizeof(int));
    ALLOC(b->d_route, F * sizeof(int));
    ALLOC(b->d_fb, (F + 1) * sizeof(int));
    if (b->Nmax <= 176 && b->mmax <= 104) ALLOC(b->d_L, F * gain_tiled_scratch_doubles(b->mmax) * sizeof(double));
    if (b->large) {
        ALLOC(b->d_LS, F * large_scratch_doubles_S(b->mmax) * sizeof(double));
        ALLOC(b->d_LL, F * large_scratch_doubles_S(b->mmax) * sizeof(double));
        ALLOC(b->d_LT, F * large_scratch_doubles_T(b->mmax) * sizeof(double));
    }
    ALLOC(b->dd_z, F * nm * 2 * sizeof(double));
    ALLOC(b->dd_R, F * nm * 4 * sizeof(double));
    ALLOC(b->dd_pass, F * nm);
#undef ALLOC
    if (cudaMallocHost((void**)&b->h_z, F * nm * 2 * sizeof(double)) != cudaSuccess || cudaMallocHost((void**)&b->h_R, F * nm * 4 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&b->h_pass, F * nm) != cudaSuccess || cudaMallocHost((void**)&b->h_out, F * (BASE + nm * 3) * sizeof(double)) != cudaSuccess) {
        ekfvio_batch_destroy(b);
        return fail_msg("cudaMallocHost failed");
    }
    {
        const EkfPtrs pp = ptrs(b);
        b->lower_ok = !(b->prm.flags & (EKFVIO_FLAG_FORCE_GENERAL_PATH | EKFVIO_FLAG_LITERAL_JOSEPH | 0x100u | 0x200u | 0x400u)) && !b->large &&
                      gain_tiled_supported(pp) && joseph_sym_supported(pp) && process_lower_capable(pp);
        const char* nf = getenv("EKFVIO_NO_FUSED_UPDATE");       // (experiments: the round-1/2 three-kernel update)
        b->fused_ok = b->lower_ok && update_fused_supported(pp) && !(nf && atoi(nf) != 0);
    }
    if (cudaStreamCreateWithFlags(&b->copy_st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_inputs_free, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_state, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->ev_entry, cudaEventDisableTiming) != cudaSuccess) {
        ekfvio_batch_destroy(b);
        return fail_msg("stream / event creation failed");
    }
    *out = b;
    int rc = ekfvio_batch_reset(b, nullptr);
    if (rc) { ekfvio_batch_destroy(b); *out = nullptr; return rc; }
    CU(cudaDeviceSynchronize());
    return 0;
}

int ekfvio_batch_num_filters(const ekfvio_batch* b) { return b ? b->F : 0; }
int ekfvio_batch_max_features(const ekfvio_batch* b) { return b ? b->nmax : 0; }
long long ekfvio_batch_launch_count(const ekfvio_batch* b) { return b ? b->launches : 0; }

int ekfvio_batch_reset(ekfvio_batch* b, void* stream) {
    CU(cudaSetDevice(b->device));
    b->state_ev_valid = false;   // state written on the caller's stream: downloads order behind that stream
    b->upper_stale = false;
    CU(launch_reset(ptrs(b), b->d_P[b->cur], (cudaStream_t)stream));
    b->launches += 1;
    return 0;
}

int ekfvio_batch_add_features(ekfvio_batch* b, const int* d_k, const double* d_uv, int kmax, void* stream) {
    CU(cudaSetDevice(b->device));
    b->state_ev_valid = false;   // state written on the caller's stream: downloads order behind that stream
    CU(launch_add_features(ptrs(b), b->d_P[b->cur], d_k, d_uv, kmax, (cudaStream_t)stream));
    b->launches += 1;
    return 0;
}

int ekfvio_batch_graph_replayed(ekfvio_batch* b, int sigma_buffer_flips) {
    if (!b || sigma_buffer_flips < 0) return fail_msg("ekfvio_batch_graph_replayed: bad arguments");
    b->cur ^= sigma_buffer_flips & 1;
    b->state_ev_valid = b->inputs_ev_valid = false;
    return 0;
}

int ekfvio_batch_graph_state(const ekfvio_batch* b) { return b ? ((b->cur & 1) | (b->upper_stale ? 2 : 0)) : -1; }
int ekfvio_batch_graph_state_restore(ekfvio_batch* b, int token) {
    if (!b || token < 0 || token > 3) return fail_msg("ekfvio_batch_graph_state_restore: bad token");
    b->cur = token & 1; b->upper_stale = (token & 2) != 0;
    b->state_ev_valid = b->inputs_ev_valid = false;
    return 0;
}

int ekfvio_batch_remove_features(ekfvio_batch* b, const uint8_t* d_remove, void* stream) {
    CU(cudaSetDevice(b->device));
    if (b->nmax == 0) return 0;
    if (ensure_full_sigma(b, (cudaStream_t)stream)) return 1;
    b->state_ev_valid = false;
    CU(launch_remove_features(ptrs(b), b->d_P[b->cur], b->d_P[b->cur ^ 1], d_remove, (cudaStream_t)stream));
    b->cur ^= 1;
    b->launches += 1;
    return 0;
}

int ekfvio_batch_process(ekfvio_batch* b, const double* d_dt, void* stream) {
    CU(cudaSetDevice(b->device));
    // (a lower-mode process reads rows 0..21 and each feature row up to its diagonal block only: a stale upper part does no harm)
    if (!b->lower_ok && ensure_full_sigma(b, (cudaStream_t)stream)) return 1;
    b->timer.begin(0, (cudaStream_t)stream);
    CU(launch_process_general(ptrs(b), b->d_P[b->cur], b->d_P[b->cur ^ 1], d_dt, 0, nullptr, (cudaStream_t)stream, &b->launches, b->lower_ok ? 1 : 0));
    b->timer.end((cudaStream_t)stream);
    b->upper_stale = b->lower_ok;
    b->last_stream = (cudaStream_t)stream;
    if (stream_is_capturing((cudaStream_t)stream)) b->state_ev_valid = false;
    else { CU(cudaEventRecord(b->ev_state, (cudaStream_t)stream)); b->state_ev_valid = true; }
    b->cur ^= 1;
    return 0;
}

int ekfvio_batch_process_dt(ekfvio_batch* b, double dt, void* stream) {
    CU(cudaSetDevice(b->device));
    CU(launch_fill_dt(b->d_dt, dt, b->F, (cudaStream_t)stream));
    b->launches += 1;
    return ekfvio_batch_process(b, b->d_dt, stream);
}

int ekfvio_batch_linearize(ekfvio_batch* b, const double* d_dt, double* d_F, void* stream) {
    CU(cudaSetDevice(b->device));
    CU(launch_process_general(ptrs(b), b->d_P[b->cur], nullptr, d_dt, 1, d_F, (cudaStream_t)stream, &b->launches, 0));
    return 0;
}

int ekfvio_batch_update(ekfvio_batch* b, const double* d_z, const double* d_R, const uint8_t* d_pass, void* stream) {
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    EkfPtrs pp = ptrs(b);
    const bool general = (b->prm.flags & EKFVIO_FLAG_FORCE_GENERAL_PATH) != 0;
    if (b->large && !general) {   // blocked multi-CTA-per-filter path for large states
        LargePtrs lp;
        lp.S = b->d_LS; lp.L = b->d_LL; lp.T = b->d_LT;
        lp.mp = (b->mmax + 63) / 64 * 64; lp.nblk = lp.mp / 64; lp.nrt_max = (b->Nmax + 1 + 63) / 64;
        CU(launch_update_large(pp, lp, b->d_P[b->cur], b->d_P[b->cur ^ 1], d_z, d_R, d_pass, st, &b->launches, &b->timer));
        if (stream_is_capturing(st)) b->state_ev_valid = b->inputs_ev_valid = false;
        else { CU(cudaEventRecord(b->ev_state, st)); CU(cudaEventRecord(b->ev_inputs_free, st)); b->state_ev_valid = b->inputs_ev_valid = true; }
        b->cur ^= 1;
        return 0;
    }
    const bool gain_tiled = !(b->prm.flags & (EKFVIO_FLAG_FORCE_GENERAL_PATH | 0x100u)) && gain_tiled_supported(pp);
    const bool need_S_scratch = gain_general_smem_doubles(b->mmax) == 0;
    if (!b->d_S && need_S_scratch) {
        size_t bytes = (size_t)b->F * ((size_t)b->mmax * b->mmax + b->mmax) * sizeof(double);
        CU(cudaMalloc((void**)&b->d_S, bytes));
    }
    const bool fused = b->fused_ok && gain_tiled;
    if (fused) {
        // the whole update of every symmetric, well-conditioned filter in one kernel; it marks those filters ROUTE_DONE and the
        // launches below only serve the rest (normally none: a few flag reads per CTA)
        b->timer.begin(4, st);
        pp.fb = b->d_fb;
        CU(cudaMemsetAsync(b->d_fb, 0, sizeof(int), st));
        CU(launch_update_fused(pp, b->d_P[b->cur], b->d_P[b->cur ^ 1], d_z, d_R, d_pass, st));
        b->timer.end(st);
        b->launches += 1;
        pp.fused = 1;
    }
    if (gain_tiled) {
        b->timer.begin(1, st);
        CU(launch_gain_tiled(0, pp, b->d_P[b->cur], d_z, d_R, d_pass, st));
        b->timer.end(st);
        b->timer.begin(3, st);
        CU(launch_gain_tiled(1, pp, b->d_P[b->cur], d_z, d_R, d_pass, st));
        b->timer.end(st);
        b->launches += (b->prm.flags & EKFVIO_FLAG_LITERAL_JOSEPH) ? 1 : 2;   // forward-only kernel + full solve (each skips the other's filters)
    } else {
        b->timer.begin(1, st);
        CU(launch_gain_general(pp, b->d_P[b->cur], d_z, d_R, d_pass, b->d_S, st));
        b->timer.end(st);
    }
    // the gain kernels have consumed z / R / pass and written the new state: the next upload and the
    // state download may proceed while the covariance update runs
    if (stream_is_capturing(st)) b->state_ev_valid = b->inputs_ev_valid = false;
    else { CU(cudaEventRecord(b->ev_state, st)); CU(cudaEventRecord(b->ev_inputs_free, st)); b->state_ev_valid = b->inputs_ev_valid = true; }
    b->timer.begin(2, st);
    {
        const bool tiled = gain_tiled && !(b->prm.flags & 0x200u) && joseph_tiled_supported(pp);
        const bool sym = tiled && !(b->prm.flags & 0x400u) && joseph_sym_supported(pp);
        if (sym) {   // ROUTE_SYM: lower-triangle kernel; ROUTE_TILED (few): full kernel
            CU(launch_joseph_sym(pp, b->d_P[b->cur], b->d_P[b->cur ^ 1], st));
            CU(launch_joseph_tiled(pp, b->d_P[b->cur], b->d_P[b->cur ^ 1], 1, st));
            b->launches += 1;
        } else if (tiled) CU(launch_joseph_tiled(pp, b->d_P[b->cur], b->d_P[b->cur ^ 1], 0, st));
        else CU(launch_joseph_general(pp, b->d_P[b->cur], b->d_P[b->cur ^ 1], st));
    }
    b->timer.end(st);
    b->cur ^= 1;
    b->upper_stale = fused;     // the covariance kernels write the whole matrix, ekf_update_fused its lower form
    b->launches += 2;
    return 0;
}

int ekfvio_batch_check_sigma(ekfvio_batch* b, int* d_neg_diag, double* d_max_asym, void* stream) {
    CU(cudaSetDevice(b->device));
    if (ensure_full_sigma(b, (cudaStream_t)stream)) return 1;
    CU(launch_check_sigma(ptrs(b), b->d_P[b->cur], d_neg_diag, d_max_asym, (cudaStream_t)stream));
    b->launches += 1;
    return 0;
}

int ekfvio_batch_accumulate_errors(ekfvio_batch* b, const double* d_truth_mu, double* d_acc, void* stream) {
    CU(cudaSetDevice(b->device));
    CU(launch_accumulate_errors(ptrs(b), d_truth_mu, d_acc, (cudaStream_t)stream));
    b->launches += 1;
    return 0;
}

int ekfvio_batch_enable_timing(ekfvio_batch* b, int on) {
    CU(cudaSetDevice(b->device));
    b->timer.reset();
    b->timer.on = on != 0;
    return 0;
}

int ekfvio_batch_get_timing(ekfvio_batch* b, double* ms8, long long* count8) {
    CU(cudaSetDevice(b->device));
    b->timer.resolve();
    for (int i = 0; i < KernelTimer::SLOTS; ++i) { if (ms8) ms8[i] = b->timer.ms[i]; if (count8) count8[i] = b->timer.cnt[i]; }
    return 0;
}

int ekfvio_measure_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops) {
    CU(cudaSetDevice(device));
    CU(measure_fp64_peak(dmma_tflops, dfma_tflops));
    return 0;
}

int ekfvio_batch_get_view(ekfvio_batch* b, ekfvio_batch_view* v) {
    if (b->upper_stale) { CU(cudaSetDevice(b->device)); if (ensure_full_sigma(b, b->last_stream)) return 1; }   // zero-copy readers see the whole matrix (ordered behind the last process())
    v->d_mu = b->d_mu; v->d_feat = b->d_feat; v->d_P = b->d_P[b->cur]; v->d_nfeat = b->d_nfeat; v->d_status = b->d_status;
    v->ldP = b->ldP; v->num_filters = b->F; v->max_features = b->nmax; v->d_klt_last = b->d_klt_last;
    return 0;
}

int ekfvio_batch_get_state_range(ekfvio_batch* b, int first, int count, double* h_mu, double* h_feat, double* h_P, int* h_nfeat, double* h_cache,
                                 uint8_t* h_flags, double* h_klt_last, int* h_status, int* h_route) {
    if (!b || first < 0 || count < 0 || first + count > b->F) return fail_msg("get_state_range: filter range outside the batch");
    CU(cudaSetDevice(b->device));
    if (h_P && ensure_full_sigma(b, b->last_stream)) return 1;
    CU(cudaDeviceSynchronize());
    if (count == 0) return 0;
    size_t F = count, o = first, nm = b->nmax;
    if (h_mu) CU(cudaMemcpy(h_mu, b->d_mu + o * BASE, F * BASE * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_feat && nm) CU(cudaMemcpy(h_feat, b->d_feat + o * nm * 3, F * nm * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_nfeat) CU(cudaMemcpy(h_nfeat, b->d_nfeat + o, F * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_cache) CU(cudaMemcpy(h_cache, b->d_cache + o * 7, F * 7 * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_flags && nm) CU(cudaMemcpy(h_flags, b->d_flags + o * nm, F * nm, cudaMemcpyDeviceToHost));
    if (h_klt_last && nm) CU(cudaMemcpy(h_klt_last, b->d_klt_last + o * nm * 2, F * nm * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_status) CU(cudaMemcpy(h_status, b->d_status + o, F * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_route) CU(cudaMemcpy(h_route, b->d_route + o, F * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_P) {
        double* tmp = nullptr;
        size_t bytes = F * (size_t)b->Nmax * b->Nmax * sizeof(double);
        CU(cudaMalloc((void**)&tmp, bytes));
        cudaError_t e = launch_pack_P(b->d_P[b->cur] + o * b->ldP * b->ldP, tmp, b->ldP, b->Nmax, count, 1, nullptr);
        b->launches += 1;
        if (e == cudaSuccess) e = cudaMemcpy(h_P, tmp, bytes, cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        if (e != cudaSuccess) return ekfvio::fail("get_state P", e);
    }
    return 0;
}

int ekfvio_batch_get_state(ekfvio_batch* b, double* h_mu, double* h_feat, double* h_P, int* h_nfeat, double* h_cache, uint8_t* h_flags,
                           double* h_klt_last, int* h_status) {
    return ekfvio_batch_get_state_range(b, 0, b ? b->F : 0, h_mu, h_feat, h_P, h_nfeat, h_cache, h_flags, h_klt_last, h_status, nullptr);
}

int ekfvio_batch_set_state(ekfvio_batch* b, const double* h_mu, const double* h_feat, const double* h_P, const int* h_nfeat,
                           const double* h_cache, const uint8_t* h_flags, const double* h_klt_last) {
    CU(cudaSetDevice(b->device));
    b->state_ev_valid = false;
    if (ensure_full_sigma(b, b->last_stream)) return 1;
    CU(cudaDeviceSynchronize());
    size_t F = b->F, nm = b->nmax;
    if (h_nfeat) {
        for (size_t i = 0; i < F; ++i) if (h_nfeat[i] < 0 || h_nfeat[i] > b->nmax) return fail_msg("set_state: nfeat out of range");
        CU(cudaMemcpy(b->d_nfeat, h_nfeat, F * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (h_mu) CU(cudaMemcpy(b->d_mu, h_mu, F * BASE * sizeof(double), cudaMemcpyHostToDevice));
    if (h_feat && nm) CU(cudaMemcpy(b->d_feat, h_feat, F * nm * 3 * sizeof(double), cudaMemcpyHostToDevice));
    if (h_cache) CU(cudaMemcpy(b->d_cache, h_cache, F * 7 * sizeof(double), cudaMemcpyHostToDevice));
    if (h_flags && nm) CU(cudaMemcpy(b->d_flags, h_flags, F * nm, cudaMemcpyHostToDevice));
    if (h_klt_last && nm) CU(cudaMemcpy(b->d_klt_last, h_klt_last, F * nm * 2 * sizeof(double), cudaMemcpyHostToDevice));
    if (h_P) {
        // filters whose Sigma is not symmetric (beyond rounding) are marked so that the update
        // takes the kernels that make no symmetry assumption
        {
            std::vector<int> asym(F, 0), nf(F, 0);
            CU(cudaMemcpy(nf.data(), b->d_nfeat, F * sizeof(int), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(asym.data(), b->d_asym, F * sizeof(int), cudaMemcpyDeviceToHost));
            const size_t Nm = b->Nmax;
            for (size_t f = 0; f < F; ++f) {
                const double* P = h_P + f * Nm * Nm;
                const int N = BASE + 3 * nf[f];
                double mx = 0, ma = 0;
                for (int i = 0; i < N; ++i)
                    for (int j = 0; j <= i; ++j) {
                        double a = P[i * Nm + j], c = P[j * Nm + i];
                        mx = std::max(mx, std::max(std::fabs(a), std::fabs(c)));
                        ma = std::max(ma, std::fabs(a - c));
                    }
                if (ma > 1e-13 * mx) asym[f] = 1;
            }
            CU(cudaMemcpy(b->d_asym, asym.data(), F * sizeof(int), cudaMemcpyHostToDevice));
        }
        double* tmp = nullptr;
        size_t bytes = F * (size_t)b->Nmax * b->Nmax * sizeof(double);
        CU(cudaMalloc((void**)&tmp, bytes));
        cudaError_t e = cudaMemcpy(tmp, h_P, bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = launch_pack_P(b->d_P[b->cur], tmp, b->ldP, b->Nmax, b->F, 0, nullptr);
        b->launches += 1;
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(tmp);
        if (e != cudaSuccess) return ekfvio::fail("set_state P", e);
    }
    return 0;
}

int ekfvio_batch_add_features_h(ekfvio_batch* b, const int* h_k, const double* h_uv, int kmax, void* stream) {
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (kmax > b->nmax) return fail_msg("add_features_h: kmax exceeds max_features");
    size_t F = b->F;
    int* d_k = nullptr; double* d_uv = nullptr;
    CU(cudaMalloc((void**)&d_k, F * sizeof(int)));
    if (cudaMalloc((void**)&d_uv, F * (size_t)(kmax > 0 ? kmax : 1) * 2 * sizeof(double)) != cudaSuccess) { cudaFree(d_k); return fail_msg("add_features_h: cudaMalloc"); }
    cudaError_t e = cudaMemcpyAsync(d_k, h_k, F * sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && kmax > 0) e = cudaMemcpyAsync(d_uv, h_uv, F * (size_t)kmax * 2 * sizeof(double), cudaMemcpyHostToDevice, st);
    int rc = 0;
    if (e == cudaSuccess) rc = ekfvio_batch_add_features(b, d_k, d_uv, kmax, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_k); cudaFree(d_uv);
    if (e != cudaSuccess) return ekfvio::fail("add_features_h", e);
    return rc;
}

int ekfvio_batch_update_h(ekfvio_batch* b, const double* h_z, const double* h_R, const uint8_t* h_pass, void* stream) {
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t F = b->F, nm = b->nmax;
    if (nm == 0) return ekfvio_batch_update(b, b->dd_z, b->dd_R, b->dd_pass, stream);
    // page-locked caller buffers are DMA'd from where they are; pageable ones are staged through the
    // batch's pinned buffers.  The copies run on the batch's copy stream: they wait for the previous
    // update's gain kernels (the last readers of the device input buffers) and overlap whatever the
    // caller's stream is still running (typically process()).
    const void* src[3] = {h_z, h_R, h_pass};
    void* stg[3] = {b->h_z, b->h_R, b->h_pass};
    void* dst[3] = {b->dd_z, b->dd_R, b->dd_pass};
    const size_t bytes[3] = {F * nm * 2 * sizeof(double), F * nm * 4 * sizeof(double), F * nm};
    if (b->inputs_ev_valid) CU(cudaStreamWaitEvent(b->copy_st, b->ev_inputs_free, 0));
    else if (!stream_is_capturing(st)) {   // no record of the last reader of the device input buffers: order the upload behind the caller's stream
        CU(cudaEventRecord(b->ev_entry, st));
        CU(cudaStreamWaitEvent(b->copy_st, b->ev_entry, 0));
    }
    // (not ordered behind earlier work on `stream`: that is what lets the upload overlap process(); the host buffers must hold
    // their final contents when this function is called — see include/ekfvio_c.h)
    bool synced = false;
    for (int i = 0; i < 3; ++i) {
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, src[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) {
            cudaGetLastError();
            if (!synced) { CU(cudaStreamSynchronize(b->copy_st)); synced = true; }   // staging buffers free again
            memcpy(stg[i], src[i], bytes[i]);
        }
        CU(cudaMemcpyAsync(dst[i], pinned ? src[i] : stg[i], bytes[i], cudaMemcpyHostToDevice, b->copy_st));
    }
    CU(cudaEventRecord(b->ev_h2d, b->copy_st));
    CU(cudaStreamWaitEvent(st, b->ev_h2d, 0));
    return ekfvio_batch_update(b, b->dd_z, b->dd_R, b->dd_pass, stream);
}

// convolveBaseState / convolveFeature as single evaluations (TightlyCoupledEKF.cpp:328-460).  Filter `f` of the batch supplies
// and receives the dq_inv cache that convolveFeature keys on omega alone (E2); the filter's state is not touched.
static int convolve_single(ekfvio_batch* b, int f, const double* h_base, const double* h_feat3, double dt, int which, double* h_out) {
    if (!b || f < 0 || f >= b->F || !h_base || !h_out || (which == 1 && !h_feat3)) return fail_msg("ekfvio_batch_convolve: bad arguments");
    CU(cudaSetDevice(b->device));
    double io[57] = {0};
    memcpy(io, h_base, BASE * sizeof(double));
    if (which == 1) memcpy(io + 22, h_feat3, 3 * sizeof(double));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(io + 25, b->d_cache + (size_t)f * 7, 7 * sizeof(double), cudaMemcpyDeviceToHost));
    double* d_io = nullptr;
    CU(cudaMalloc((void**)&d_io, sizeof(io)));
    cudaError_t e = cudaMemcpy(d_io, io, sizeof(io), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_convolve_single(d_io, dt, which, (b->prm.flags & EKFVIO_FLAG_FRESH_DQ_CACHE) ? 1 : 0, nullptr);
    b->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpy(io, d_io, sizeof(io), cudaMemcpyDeviceToHost);
    cudaFree(d_io);
    if (e != cudaSuccess) return ekfvio::fail("ekfvio_batch_convolve", e);
    if (which == 1) CU(cudaMemcpy(b->d_cache + (size_t)f * 7, io + 25, 7 * sizeof(double), cudaMemcpyHostToDevice));
    memcpy(h_out, which == 0 ? io + 32 : io + 54, (which == 0 ? BASE : 3) * sizeof(double));
    return 0;
}
int ekfvio_batch_convolve_base_h(ekfvio_batch* b, int f, const double* h_base22, double dt, double* h_out22) {
    return convolve_single(b, f, h_base22, nullptr, dt, 0, h_out22);
}
int ekfvio_batch_convolve_feature_h(ekfvio_batch* b, int f, const double* h_base22, const double* h_feat3, double dt, double* h_out3) {
    return convolve_single(b, f, h_base22, h_feat3, dt, 1, h_out3);
}

int ekfvio_batch_linearize_h(ekfvio_batch* b, double dt, double* h_F) {
    CU(cudaSetDevice(b->device));
    double* dF = nullptr;
    const size_t bytes = (size_t)b->F * b->Nmax * b->Nmax * sizeof(double);
    CU(cudaMalloc((void**)&dF, bytes));
    cudaError_t e = launch_fill_dt(b->d_dt, dt, b->F, nullptr);
    b->launches += 1;
    int rc = 0;
    if (e == cudaSuccess) rc = ekfvio_batch_linearize(b, b->d_dt, dF, nullptr);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_F, dF, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dF);
    if (e != cudaSuccess) return ekfvio::fail("linearize_h", e);
    return rc;
}

int ekfvio_batch_check_sigma_h(ekfvio_batch* b, int* h_neg_diag, double* h_max_asym) {
    CU(cudaSetDevice(b->device));
    int* dn = nullptr; double* da = nullptr;
    CU(cudaMalloc((void**)&dn, b->F * sizeof(int)));
    if (cudaMalloc((void**)&da, b->F * sizeof(double)) != cudaSuccess) { cudaFree(dn); return fail_msg("check_sigma_h: cudaMalloc"); }
    int rc = ekfvio_batch_check_sigma(b, dn, da, nullptr);
    cudaError_t e = cudaSuccess;
    if (rc == 0) e = cudaMemcpy(h_neg_diag, dn, b->F * sizeof(int), cudaMemcpyDeviceToHost);
    if (rc == 0 && e == cudaSuccess) e = cudaMemcpy(h_max_asym, da, b->F * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dn); cudaFree(da);
    if (e != cudaSuccess) return ekfvio::fail("check_sigma_h", e);
    return rc;
}

int ekfvio_batch_read_mu_h(ekfvio_batch* b, double* h_mu, double* h_feat, void* stream) {
    CU(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t F = b->F, nm = b->nmax;
    // the download waits only for the last kernel that wrote the state, not for the covariance update
    // behind it; page-locked caller buffers receive the DMA directly
    cudaStream_t cs = st;
    if (b->state_ev_valid) { cs = b->copy_st; CU(cudaStreamWaitEvent(cs, b->ev_state, 0)); }
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes attr;
        const bool ok = p && cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!ok) cudaGetLastError();
        return ok;
    };
    const bool want_feat = h_feat && nm;
    const bool pin_mu = is_pinned(h_mu), pin_feat = want_feat && is_pinned(h_feat);
    if (h_mu) CU(cudaMemcpyAsync(pin_mu ? h_mu : b->h_out, b->d_mu, F * BASE * sizeof(double), cudaMemcpyDeviceToHost, cs));
    if (want_feat) CU(cudaMemcpyAsync(pin_feat ? h_feat : b->h_out + F * BASE, b->d_feat, F * nm * 3 * sizeof(double), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    if (h_mu && !pin_mu) memcpy(h_mu, b->h_out, F * BASE * sizeof(double));
    if (want_feat && !pin_feat) memcpy(h_feat, b->h_out + F * BASE, F * nm * 3 * sizeof(double));
    return 0;
}

}  // extern "C"
