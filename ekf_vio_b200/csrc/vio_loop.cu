// EKFVIO::addFrame (EKFVIO.cpp:139-196) as a device-side frame loop over S independent sequences:
// a composition of the library's own C-ABI components (batched EKF, pyramidal KLT, FAST replenishment)
// plus the small conversion kernels the reference does on the host between them
// (KLTTracker.cpp:53-59 point set-up, EKFVIO.cpp:201-217 measurement hand-over, :224-311 replenishment).
#include <new>
#include <string>

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ekfvio_c.h"

namespace ekfvio {
int fail(const char* what, cudaError_t e);
int fail_msg(const std::string& msg);
}  // namespace ekfvio
using ekfvio::fail_msg;

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return ekfvio::fail(#x, e_); } while (0)
#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

struct ekfvio_vio {
    int device = 0, S = 0, width = 0, height = 0, nmax = 0;
    ekfvio_vio_params prm{};
    ekfvio_klt_params klt_prm{};
    ekfvio_batch* ekf = nullptr;
    ekfvio_klt* klt = nullptr;
    ekfvio_fast* fast = nullptr;
    ekfvio_batch_view view{};              // state pointers of the filters (stable for the batch's lifetime; d_P is not used here)
    int frames = 0, cur_slot = 0;          // pyramid slot of the most recent frame
    long long launches = 0;
    float* d_K_prev = nullptr;             // [S][9] K of the previous frame (metric2Pixel(lf, .))
    float* d_prev_pts = nullptr; float* d_next_pts = nullptr; uint8_t* d_status = nullptr; float* d_err = nullptr; int* d_npts = nullptr;
    float* d_measured = nullptr; float* d_cov = nullptr; uint8_t* d_passed = nullptr;
    double* d_z = nullptr; double* d_R = nullptr; uint8_t* d_pass = nullptr;
    short* d_kp = nullptr; int* d_count = nullptr; float* d_exist = nullptr; int* d_needed = nullptr;
    short* d_new_px = nullptr; float* d_new_metric = nullptr; int* d_nnew = nullptr; double* d_uv = nullptr; int* d_k = nullptr;
    // CUDA graph replay: fixed-address copies of the per-call inputs and one executable graph per pyramid-slot parity
    uint8_t* d_frames_in = nullptr; float* d_K_in = nullptr; double* d_dt_in = nullptr;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    long long graph_launches[2] = {0, 0};    // kernel launches one replay stands for
    int graph_batch_state[2] = {-1, -1};     // ekfvio_batch_graph_state at capture: a graph is only replayed in that state
    int graph_state_after[2] = {-1, -1};     // ... and the state the recorded frame leaves the batch's bookkeeping in
    int parity_seen[2] = {0, 0};
    cudaStream_t own_st = nullptr;           // capture is not allowed on the legacy default stream: graph frames of such callers run here
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
};

namespace {

// KLTTracker.cpp:53-59: prev = metric2Pixel(lf, last KLT result), initial = Feature::getPixel(cf), both in
// float with the reference's linear indices into the column-major K (E1); also Feature::getPixel of every
// feature for the replenishment check image (EKFVIO.cpp:258-260) and needed = NUM_FEATURES - size (:246).
__global__ void vio_points_kernel(const double* __restrict__ feat, const double* __restrict__ klt_last, const int* __restrict__ nfeat, int nmax,
                                  const float* __restrict__ K_prev, const float* __restrict__ K_cur, float* __restrict__ prev_pts,
                                  float* __restrict__ next_pts, int* __restrict__ npts, int* __restrict__ needed, int num_features) {
    const int s = blockIdx.x, n = nfeat[s];
    const float* Kp = K_prev + (size_t)s * 9;
    const float* Kc = K_cur + (size_t)s * 9;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t o = (size_t)s * nmax + i;
        if (prev_pts) {
            const float lx = (float)klt_last[o * 2], ly = (float)klt_last[o * 2 + 1];
            prev_pts[o * 2] = __fadd_rn(__fmul_rn(lx, Kp[0]), Kp[2]);
            prev_pts[o * 2 + 1] = __fadd_rn(__fmul_rn(ly, Kp[4]), Kp[5]);
        }
        const float u = (float)feat[o * 3], v = (float)feat[o * 3 + 1];
        next_pts[o * 2] = __fadd_rn(__fmul_rn(Kc[0], u), Kc[2]);
        next_pts[o * 2 + 1] = __fadd_rn(__fmul_rn(Kc[4], v), Kc[5]);
    }
    if (threadIdx.x == 0) {
        if (npts) npts[s] = n;
        if (needed) needed[s] = n < num_features ? num_features - n : 0;
    }
}

// EKFVIO.cpp:215-217: the tracker's float outputs become the filter's measurement, covariance and pass vectors
__global__ void vio_measurement_kernel(const float* __restrict__ measured, const float* __restrict__ cov, const uint8_t* __restrict__ passed,
                                       const int* __restrict__ nfeat, int nmax, double* __restrict__ z, double* __restrict__ R,
                                       uint8_t* __restrict__ pass) {
    const int s = blockIdx.x, n = nfeat[s];
    for (int i = threadIdx.x; i < nmax; i += blockDim.x) {
        const size_t o = (size_t)s * nmax + i;
        const bool ok = i < n && passed[o];
        pass[o] = ok ? 1 : 0;
        z[o * 2] = ok ? (double)measured[o * 2] : 0.0; z[o * 2 + 1] = ok ? (double)measured[o * 2 + 1] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) R[o * 4 + q] = ok ? (double)cov[o * 4 + q] : 0.0;
    }
}

// EKFVIO.cpp:301-308: the accepted keypoints' metric coordinates go to addNewFeatures
__global__ void vio_new_features_kernel(const float* __restrict__ new_metric, const int* __restrict__ n_new, const int* __restrict__ nfeat, int nmax,
                                        int max_new, double* __restrict__ uv, int* __restrict__ k) {
    const int s = blockIdx.x;
    int kk = n_new[s];
    if (kk > nmax - nfeat[s]) kk = nmax - nfeat[s];
    for (int i = threadIdx.x; i < kk; i += blockDim.x) {
        uv[((size_t)s * max_new + i) * 2] = (double)new_metric[((size_t)s * max_new + i) * 2];
        uv[((size_t)s * max_new + i) * 2 + 1] = (double)new_metric[((size_t)s * max_new + i) * 2 + 1];
    }
    if (threadIdx.x == 0) k[s] = kk;
}

}  // namespace

extern "C" {

void ekfvio_vio_default_params(ekfvio_vio_params* p) {
    p->num_features = 100;           // Params.h:46
    p->fast_threshold = 50;          // Params.h:24
    p->min_new_feature_dist = 30;    // Params.h:43
    p->remove_lost_features = 0;     // the reference never removes features
    p->use_cuda_graph = 1;
}

int ekfvio_vio_destroy(ekfvio_vio* v) {
    if (!v) return 0;
    cudaSetDevice(v->device);
    ekfvio_batch_destroy(v->ekf); ekfvio_klt_destroy(v->klt); ekfvio_fast_destroy(v->fast);
    cudaFree(v->d_K_prev); cudaFree(v->d_prev_pts); cudaFree(v->d_next_pts); cudaFree(v->d_status); cudaFree(v->d_err); cudaFree(v->d_npts);
    cudaFree(v->d_measured); cudaFree(v->d_cov); cudaFree(v->d_passed); cudaFree(v->d_z); cudaFree(v->d_R); cudaFree(v->d_pass);
    cudaFree(v->d_kp); cudaFree(v->d_count); cudaFree(v->d_exist); cudaFree(v->d_needed); cudaFree(v->d_new_px); cudaFree(v->d_new_metric);
    cudaFree(v->d_nnew); cudaFree(v->d_uv); cudaFree(v->d_k);
    cudaFree(v->d_frames_in); cudaFree(v->d_K_in); cudaFree(v->d_dt_in);
    for (int i = 0; i < 2; ++i) if (v->graph[i]) cudaGraphExecDestroy(v->graph[i]);
    if (v->own_st) cudaStreamDestroy(v->own_st);
    if (v->ev_in) cudaEventDestroy(v->ev_in);
    if (v->ev_out) cudaEventDestroy(v->ev_out);
    delete v;
    return 0;
}

int ekfvio_vio_create(ekfvio_vio** out, int device, int num_sequences, int width, int height, const ekfvio_params* ekf_params,
                      const ekfvio_klt_params* klt_params, const ekfvio_vio_params* vio_params) {
    if (!out || num_sequences <= 0 || width <= 0 || height <= 0) return fail_msg("ekfvio_vio_create: bad arguments");
    ekfvio_vio* v = new (std::nothrow) ekfvio_vio;
    if (!v) return fail_msg("out of host memory");
    v->device = device; v->S = num_sequences; v->width = width; v->height = height;
    if (vio_params) v->prm = *vio_params; else ekfvio_vio_default_params(&v->prm);
    if (klt_params) v->klt_prm = *klt_params; else ekfvio_klt_default_params(&v->klt_prm);
    if (v->prm.num_features <= 0 || v->prm.num_features > 512) { delete v; return fail_msg("ekfvio_vio_create: num_features must be within 1..512"); }
    v->nmax = v->prm.num_features;
    ekfvio_params ep;
    if (ekf_params) ep = *ekf_params; else ekfvio_default_params(&ep);
    int rc = ekfvio_batch_create(&v->ekf, device, num_sequences, v->nmax, &ep);
    if (!rc) rc = ekfvio_klt_create(&v->klt, device, width, height, num_sequences, v->nmax, 2, &v->klt_prm);
    if (!rc) rc = ekfvio_fast_create(&v->fast, device, width, height, num_sequences, 4096);
    if (!rc) rc = ekfvio_batch_get_view(v->ekf, &v->view);
    if (rc) { ekfvio_vio_destroy(v); return rc; }
    const size_t S = num_sequences, np = S * v->nmax;
    cudaError_t e = cudaSetDevice(device);
#define VALLOC(ptr, bytes) if (e == cudaSuccess) { e = cudaMalloc((void**)&(ptr), (bytes)); if (e == cudaSuccess) e = cudaMemset((ptr), 0, (bytes)); }
    VALLOC(v->d_K_prev, S * 9 * sizeof(float));
    VALLOC(v->d_prev_pts, np * 2 * sizeof(float)); VALLOC(v->d_next_pts, np * 2 * sizeof(float)); VALLOC(v->d_status, np); VALLOC(v->d_err, np * sizeof(float));
    VALLOC(v->d_npts, S * sizeof(int));
    VALLOC(v->d_measured, np * 2 * sizeof(float)); VALLOC(v->d_cov, np * 4 * sizeof(float)); VALLOC(v->d_passed, np);
    VALLOC(v->d_z, np * 2 * sizeof(double)); VALLOC(v->d_R, np * 4 * sizeof(double)); VALLOC(v->d_pass, np);
    VALLOC(v->d_kp, S * 4096 * 2 * sizeof(short)); VALLOC(v->d_count, S * sizeof(int)); VALLOC(v->d_exist, np * 2 * sizeof(float));
    VALLOC(v->d_needed, S * sizeof(int)); VALLOC(v->d_new_px, np * 2 * sizeof(short)); VALLOC(v->d_new_metric, np * 2 * sizeof(float));
    VALLOC(v->d_nnew, S * sizeof(int)); VALLOC(v->d_uv, np * 2 * sizeof(double)); VALLOC(v->d_k, S * sizeof(int));
    VALLOC(v->d_frames_in, S * (size_t)width * height); VALLOC(v->d_K_in, S * 9 * sizeof(float)); VALLOC(v->d_dt_in, S * sizeof(double));
#undef VALLOC
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->own_st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->ev_out, cudaEventDisableTiming);
    if (e != cudaSuccess) { ekfvio_vio_destroy(v); return ekfvio::fail("ekfvio_vio_create", e); }
    *out = v;
    return 0;
}

// The launch sequence of one frame (run eagerly, or recorded into a graph).
static int enqueue_frame(ekfvio_vio* v, const uint8_t* d_frames, int pitch, const float* d_K9, const double* d_dt, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int S = v->S, nmax = v->nmax;
    const ekfvio_batch_view& view = v->view;
    const int new_slot = v->frames == 0 ? 0 : (v->cur_slot ^ 1);
    // this frame's pyramid, with derivatives: it is the "previous" frame of the next call
    RC(ekfvio_klt_build_pyramid(v->klt, new_slot, d_frames, pitch, S, 1, stream));
    if (v->frames > 0) {
        RC(ekfvio_batch_process(v->ekf, d_dt, stream));                                              // EKFVIO.cpp:163
        vio_points_kernel<<<S, 128, 0, st>>>(view.d_feat, view.d_klt_last, view.d_nfeat, nmax, v->d_K_prev, d_K9, v->d_prev_pts, v->d_next_pts,
                                             v->d_npts, nullptr, v->prm.num_features);
        CU(cudaGetLastError());
        RC(ekfvio_klt_track(v->klt, v->cur_slot, new_slot, v->d_prev_pts, v->d_next_pts, v->d_status, v->d_err, v->d_npts, S, stream));
        RC(ekfvio_klt_postprocess(v->klt, v->d_next_pts, v->d_status, v->d_npts, d_K9, S, v->d_measured, v->d_cov, v->d_passed, stream));
        vio_measurement_kernel<<<S, 128, 0, st>>>(v->d_measured, v->d_cov, v->d_passed, view.d_nfeat, nmax, v->d_z, v->d_R, v->d_pass);
        CU(cudaGetLastError());
        RC(ekfvio_batch_update(v->ekf, v->d_z, v->d_R, v->d_pass, stream));                          // EKFVIO.cpp:217
        if (v->prm.remove_lost_features) RC(ekfvio_batch_remove_features(v->ekf, nullptr, stream));   // (not in the reference)
        v->launches += 2;
    }
    // replenishFeatures (EKFVIO.cpp:172 / :153)
    vio_points_kernel<<<S, 128, 0, st>>>(view.d_feat, view.d_klt_last, view.d_nfeat, nmax, d_K9, d_K9, nullptr, v->d_exist, v->d_npts, v->d_needed,
                                         v->prm.num_features);
    CU(cudaGetLastError());
    RC(ekfvio_fast_detect(v->fast, d_frames, pitch, S, v->prm.fast_threshold, 1, v->d_kp, nullptr, v->d_count, stream));
    RC(ekfvio_fast_select(v->fast, v->d_kp, v->d_count, v->d_exist, v->d_npts, nmax, v->d_needed, v->prm.min_new_feature_dist, v->klt_prm.kill_pad,
                          d_K9, v->d_new_px, v->d_new_metric, v->d_nnew, nmax, S, stream));
    vio_new_features_kernel<<<S, 128, 0, st>>>(v->d_new_metric, v->d_nnew, view.d_nfeat, nmax, nmax, v->d_uv, v->d_k);
    CU(cudaGetLastError());
    RC(ekfvio_batch_add_features(v->ekf, v->d_k, v->d_uv, nmax, stream));                            // EKFVIO.cpp:308
    CU(cudaMemcpyAsync(v->d_K_prev, d_K9, (size_t)S * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    v->launches += 2;
    v->cur_slot = new_slot;
    v->frames += 1;
    return 0;
}

int ekfvio_vio_add_frame(ekfvio_vio* v, const uint8_t* d_frames, int pitch, const float* d_K9, const double* d_dt, void* stream) {
    if (!v || !d_frames || !d_K9 || (v->frames > 0 && !d_dt)) return fail_msg("ekfvio_vio_add_frame: null argument");
    if (pitch < v->width) return fail_msg("ekfvio_vio_add_frame: pitch smaller than the frame width");
    CU(cudaSetDevice(v->device));
    cudaStream_t st = (cudaStream_t)stream;
    // The first frame and the first frame of each pyramid-slot parity run eagerly (they also configure the kernels);
    // the second frame of a parity is recorded into a graph, later ones replay it.  Replays read their inputs from
    // fixed buffers, so the caller's frame / K / dt arrays are copied there first.
    const int parity = v->cur_slot;
    if (!v->prm.use_cuda_graph || v->frames == 0 || !v->parity_seen[parity]) {
        if (v->frames > 0) v->parity_seen[parity] = 1;
        return enqueue_frame(v, d_frames, pitch, d_K9, d_dt, stream);
    }
    const bool legacy = st == nullptr || st == cudaStreamLegacy;   // cannot be captured: hop onto the loop's own stream
    cudaStream_t caller = st;
    if (legacy) {
        CU(cudaEventRecord(v->ev_in, caller));
        st = v->own_st; stream = (void*)st;
        CU(cudaStreamWaitEvent(st, v->ev_in, 0));
    }
    CU(cudaMemcpy2DAsync(v->d_frames_in, v->width, d_frames, pitch, v->width, (size_t)v->height * v->S, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(v->d_K_in, d_K9, (size_t)v->S * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(v->d_dt_in, d_dt, (size_t)v->S * sizeof(double), cudaMemcpyDeviceToDevice, st));
    bool captured_now = false;                                   // (the capture pass already did the host-side bookkeeping)
    // The graph has the batch's Sigma ping-pong buffer baked in: if somebody stepped the batch behind the loop's back
    // (ekfvio_vio_filters() hands it out), the recording is stale — drop it and record again from the present state.
    if (v->graph[parity] && v->graph_batch_state[parity] != ekfvio_batch_graph_state(v->ekf)) {
        cudaGraphExecDestroy(v->graph[parity]);
        v->graph[parity] = nullptr;
    }
    if (!v->graph[parity]) {
        captured_now = true;
        const long long before = ekfvio_vio_launch_count(v);
        const int frames = v->frames, slot = v->cur_slot;
        const int batch_state = ekfvio_batch_graph_state(v->ekf);
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_frame(v, v->d_frames_in, v->width, v->d_K_in, v->d_dt_in, stream);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &g);
        v->frames = frames; v->cur_slot = slot;              // the recorded frame has not run yet
        if (rc || e != cudaSuccess) ekfvio_batch_graph_state_restore(v->ekf, batch_state);   // nothing of the recording ran: undo the batch's bookkeeping too
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return ekfvio::fail("cudaStreamEndCapture", e);
        v->graph_batch_state[parity] = batch_state;
        v->graph_state_after[parity] = ekfvio_batch_graph_state(v->ekf);
        v->graph_launches[parity] = ekfvio_vio_launch_count(v) - before;
        v->launches -= v->graph_launches[parity];            // counted again by the launch below
        e = cudaGraphInstantiate(&v->graph[parity], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return ekfvio::fail("cudaGraphInstantiate", e);
    }
    CU(cudaGraphLaunch(v->graph[parity], st));
    // a replay leaves the batch's host-side bookkeeping (Sigma ping-pong buffer, lower form, pending covariance pass) where the recording did
    if (!captured_now) RC(ekfvio_batch_graph_state_restore(v->ekf, v->graph_state_after[parity]));
    if (legacy) { CU(cudaEventRecord(v->ev_out, st)); CU(cudaStreamWaitEvent(caller, v->ev_out, 0)); }
    v->launches += v->graph_launches[parity];
    v->cur_slot ^= 1;
    v->frames += 1;
    return 0;
}

ekfvio_batch* ekfvio_vio_filters(ekfvio_vio* v) { return v ? v->ekf : nullptr; }
int ekfvio_vio_frame_count(const ekfvio_vio* v) { return v ? v->frames : 0; }
long long ekfvio_vio_launch_count(const ekfvio_vio* v) {
    return v ? v->launches + ekfvio_batch_launch_count(v->ekf) + ekfvio_klt_launch_count(v->klt) + ekfvio_fast_launch_count(v->fast) : 0;
}

}  // extern "C"
