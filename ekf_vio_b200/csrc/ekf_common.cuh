// Shared device code for the batched FP64 EKF kernels (sm_100a).
// Process model of TightlyCoupledEKF (reference: include/ekf_vio/TightlyCoupledEKF.cpp:328-460)
// written for one thread per evaluation; Eigen quaternion semantics are spelled out
// (Quaternion*Vector3 without normalisation, inverse = conjugate/squaredNorm).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ekfvio_c.h"
#include "timing.h"

namespace ekfvio {

constexpr int BASE = EKFVIO_BASE_STATE_SIZE;
constexpr double PRUNE_LIMIT = 1e-8 * 1e-5;  // SPARSE_THRESH * SPARSE_EPS (TightlyCoupledEKF.h:13-14)
constexpr double DELTA_SHIFT = 1e-3;         // TightlyCoupledEKF.cpp:182

struct Q4 { double w, x, y, z; };
struct V3 { double x, y, z; };

__device__ __forceinline__ V3 cross3(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen::QuaternionBase::_transformVector
__device__ __forceinline__ V3 qrot(const Q4& q, const V3& v) {
    V3 qv{q.x, q.y, q.z};
    V3 uv = cross3(qv, v);
    uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
    V3 c = cross3(qv, uv);
    return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ Q4 qinv(const Q4& q) {
    double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    if (n2 > 0.0) return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
    return {0.0, 0.0, 0.0, 0.0};
}
__device__ __forceinline__ Q4 qnormalized(const Q4& q) {
    double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}

// dq of convolveBaseState (sign = +1, TightlyCoupledEKF.cpp:340-355) and the directly built
// dq_inv of convolveFeature (sign = -1, :425-440).
__device__ __forceinline__ Q4 delta_quat(double ox, double oy, double oz, double dt, double sign) {
    double on = sqrt(ox * ox + oy * oy + oz * oz);
    if (on < 1e-10) return qnormalized(Q4{1.0, sign * ox * dt, sign * oy * dt, sign * oz * dt});
    double theta = dt * on;
    double hx = ox / on, hy = oy / on, hz = oz / on;
    double s, c;
    sincos(theta / 2, &s, &c);
    return {c, sign * hx * s, sign * hy * s, sign * hz * s};
}

// convolveBaseState (TightlyCoupledEKF.cpp:328-395).  in/out: 22 doubles.
__device__ __forceinline__ void convolve_base(const double* in, double dt, double* out) {
    V3 pos{in[0], in[1], in[2]};
    Q4 quat{in[3], in[4], in[5], in[6]};
    V3 vel{in[7], in[8], in[9]};
    V3 accel{in[13], in[14], in[15]};
    double hdt2 = 0.5 * dt * dt;
    V3 tr{dt * vel.x + hdt2 * accel.x, dt * vel.y + hdt2 * accel.y, dt * vel.z + hdt2 * accel.z};
    V3 r = qrot(quat, tr);
    Q4 dq = delta_quat(in[10], in[11], in[12], dt, 1.0);
    Q4 dqi = qinv(dq);
    V3 va{vel.x + dt * accel.x, vel.y + dt * accel.y, vel.z + dt * accel.z};
    V3 v2 = qrot(dqi, va);
    V3 a2 = qrot(dqi, accel);
    Q4 q2 = qmul(quat, dq);
    out[0] = pos.x + r.x; out[1] = pos.y + r.y; out[2] = pos.z + r.z;
    out[3] = q2.w; out[4] = q2.x; out[5] = q2.y; out[6] = q2.z;
    out[7] = v2.x; out[8] = v2.y; out[9] = v2.z;
    out[10] = in[10]; out[11] = in[11]; out[12] = in[12];
    out[13] = a2.x; out[14] = a2.y; out[15] = a2.z;
#pragma unroll
    for (int i = 16; i < 22; ++i) out[i] = in[i];
}

// convolveFeature (TightlyCoupledEKF.cpp:397-460) given the dq_inv the reference would use.
// vel/accel are base_state(7..9)/(13..15).
__device__ __forceinline__ void convolve_feature(const Q4& dqi, const V3& vel, const V3& accel, double dt, double u, double v,
                                                 double rho, double* out3) {
    V3 fp;
    fp.z = 1.0 / rho;
    fp.x = u * fp.z;
    fp.y = v * fp.z;
    double hdt2 = 0.5 * dt * dt;
    V3 tr{dt * vel.x + hdt2 * accel.x, dt * vel.y + hdt2 * accel.y, dt * vel.z + hdt2 * accel.z};
    V3 a = qrot(dqi, fp);
    V3 b = qrot(dqi, tr);
    double x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    out3[0] = x / z;
    out3[1] = y / z;
    out3[2] = 1.0 / z;
}

// Gain (K) and Joseph residual (W) panels are stored chunk-major: element (row, k) of a filter sits
// at (k/16) * ldP*16 + row*16 + k%16, i.e. each 16-column k-chunk of all rows is one contiguous
// block — the unit the covariance update streams through shared memory.  ldK = 16 * #chunks.
__host__ __device__ __forceinline__ size_t kw_at(int ldP, int row, int k) {
    return (size_t)(k >> 4) * ldP * 16 + (size_t)row * 16 + (k & 15);
}

__device__ __forceinline__ double prune(double v) { return (fabs(v) > PRUNE_LIMIT) ? v : 0.0; }

// Diagonal of Q*dt (generateProcessNoise, TightlyCoupledEKF.cpp:123-174).
__device__ __forceinline__ double process_noise_diag(int i, double dt) {
    if (i <= 6) return 0.0001 * dt;
    if (i <= 9) return 0.01 * dt;
    if (i <= 15) return 5 * dt;
    if (i <= 21) return 0.001 * dt;
    return 0.0001 * dt;
}

// FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8).  Fragments (PTX ISA, mma.m8n8k4.f64):
//   a : A[row = lane/4][k = lane%4]          b : B[k = lane%4][col = lane/4]
//   c0,c1 : C[row = lane/4][col = 2*(lane%4) + {0,1}]
// (not volatile: the scheduler is free to interleave fragment loads of the next tile with these)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma884_volatile(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

}  // namespace ekfvio

// ---- batch object shared by the EKF translation units ---------------------------------------
struct ekfvio_batch {
    int device = 0;
    int F = 0, nmax = 0, Nmax = 0, ldP = 0, mmax = 0, ldK = 0;
    ekfvio_params prm{};
    double illcond = 0.0;          // see ekf_kernels.h ILLCOND_RATIO
    // state
    double* d_mu = nullptr;        // [F][22]
    double* d_feat = nullptr;      // [F][nmax][3]
    double* d_P[2] = {nullptr, nullptr};  // ping-pong [F][ldP][ldP]
    int cur = 0;
    int* d_nfeat = nullptr;        // [F]
    double* d_cache = nullptr;     // [F][7]
    uint8_t* d_flags = nullptr;    // [F][nmax]
    double* d_klt_last = nullptr;  // [F][nmax][2]
    int* d_status = nullptr;       // [F]
    // scratch
    double* d_dt = nullptr;        // [F]
    double* d_K = nullptr;         // [F][ldK/16][ldP][16]   gain, chunk-major (kw_at)
    double* d_W = nullptr;         // [F][ldK/16][ldP][16]   Joseph residual panel
    double* d_S = nullptr;         // [F][mmax][mmax] (general path only; lazily allocated)
    double* d_L = nullptr;         // [F][tiles] Cholesky factor + inverse diagonal tiles (tiled path)
    double* d_LS = nullptr;        // large path: S [F][mp][mp]
    double* d_LL = nullptr;        // large path: L [F][mp][mp]
    double* d_LT = nullptr;        // large path: diagonal-block tiles + inverses
    bool large = false;
    double* d_y = nullptr;         // [F][mmax]
    int* d_idx = nullptr;          // [F][mmax]
    int* d_m = nullptr;            // [F]
    int* d_asym = nullptr;         // [F] sticky: Sigma or an R block of this filter is not symmetric
    int* d_route = nullptr;        // [F] update path of the current update (ekf_kernels.h ROUTE_*)
    int* d_fb = nullptr;           // [1 + F] fused mode: the filters ekf_update_fused left to the tiled kernels (count, indices)
    double* d_fjac = nullptr;      // [F][(22*22 + nmax*27 + nmax*9)] A | B | D
    // pinned staging + device input buffers for the *_h entry points
    double* h_z = nullptr; double* h_R = nullptr; uint8_t* h_pass = nullptr;
    double* dd_z = nullptr; double* dd_R = nullptr; uint8_t* dd_pass = nullptr;
    double* h_out = nullptr;
    long long launches = 0;
    KernelTimer timer;
    // host-buffer entry points: copies run on their own stream, ordered against the kernels by events, so
    // the upload of a step's measurements overlaps process() and the state download overlaps the covariance update
    cudaStream_t copy_st = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_inputs_free = nullptr, ev_state = nullptr, ev_entry = nullptr;
    bool inputs_ev_valid = false, state_ev_valid = false;
    // lower mode: process() of a batch on the reduced tiled update path leaves the feature rows of symmetric filters
    // complete only up to their diagonal blocks; the update that follows restores the full matrix, any other reader
    // of Sigma gets it mirrored first (ensure_full_sigma)
    bool lower_ok = false, upper_stale = false;
    // fused mode: ekf_update_fused (Sigma in registers, block-sequential update) runs first and leaves Sigma' in the same
    // lower form; the tiled kernels behind it serve the filters it leaves alone
    bool fused_ok = false;
    cudaStream_t last_stream = nullptr;       // stream of the last process()
};
