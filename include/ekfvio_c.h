/* ekfvio_c.h — C ABI of the B200-native (sm_100a) EKF + KLT hot paths of k-sheridan/ekf_vio.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types, no exceptions.
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).  Host code (the C++ facade in include/ekf_vio/, the Python binding in
 * ekf_vio_b200/capi.py, or a maintainer's own binding — see INTEGRATION.md) calls only this.
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on failure; ekfvio_last_error() gives the
 *     message of the last failure on the calling thread.  (The reference's own convention is
 *     void + ROS_ASSERT/ROS_ERROR; contract violations are reported as errors.)
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous
 *     on that stream unless stated otherwise.
 *   - Pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *   - There is no CPU fallback: if no CUDA device is usable the create calls fail.
 *
 * Batch layout (F filters, capacity nmax features each, Nmax = 22 + 3*nmax):
 *   mu    [F][22]            base state: p(0-2) q wxyz(3-6) body vel(7-9) omega(10-12)
 *                            body accel(13-15) acc bias(16-18) gyro bias(19-21)
 *                            (include/ekf_vio/TightlyCoupledEKF.h:11,29)
 *   feat  [F][nmax][3]       Feature::mu = [u, v, 1/depth]          (Feature.h:41)
 *   P     [F][Nmax][Nmax]    Sigma, dense row-major                  (TightlyCoupledEKF.h:34)
 *   nfeat [F]                features.size()
 *   cache [F][7]             convolveFeature's static cache (omega xyz, dq_inv wxyz), one
 *                            private copy per filter                 (TightlyCoupledEKF.cpp:400-403)
 *   flags [F][nmax]          Feature::delete_flag                    (Feature.h:46)
 *   klt_last [F][nmax][2]    Feature::last_result_from_klt_tracker   (Feature.h:43)
 *   status[F]                bit0: zero pivot in the S factorisation (TightlyCoupledEKF.cpp:579, Eigen::NumericalIssue)
 *                            bit1: non-finite state after an update
 *                            bit2: addNewFeatures beyond the batch's capacity (nothing appended)
 *                            bit3: informational — S was not positive definite in some update; as in the reference, whose
 *                                  unpivoted LDL^T carries on with such an S, the update was evaluated by the LDL^T kernels
 */
#ifndef EKFVIO_C_H_
#define EKFVIO_C_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EKFVIO_BASE_STATE_SIZE 22 /* TightlyCoupledEKF.h:12 */

/* Flags in ekfvio_params.flags.  0 = reference behaviour, including the reference's errors
 * (SURVEY.md §8a E1..E6).  Each fix is opt-in. */
#define EKFVIO_FLAG_FORCE_GENERAL_PATH 0x1u /* always use the general (non-tiled) kernels */
#define EKFVIO_FLAG_FRESH_DQ_CACHE 0x2u     /* fix E2: never reuse a dq_inv computed for another dt */
/* Diagnostic bits (tools/diag_parity.py, tools/process_probe.py; no effect on results other than the path taken):
 * 0x100 general gain kernel, 0x200 general covariance kernel, 0x400 tiled covariance kernel without the symmetric
 * variant, 0x800 process() stops after linearisation + state propagation, 0x1000 process() keeps the row-block covariance kernel
 * for every filter (instead of the DMMA tile kernel for symmetric filters; tools/process_tiles_probe.py, tests). */
#define EKFVIO_FLAG_LITERAL_JOSEPH 0x4u     /* evaluate (I-KH) Sigma (I-KH)' + K R K' term by term even where Sigma and R are
                                             * symmetric; by default such filters use the algebraically identical
                                             * Sigma - Z Z', Z = Sigma(:,idx) inv(L)', S = L L'  */

typedef struct ekfvio_params {
    double default_point_depth;               /* Params.h:83   DEFAULT_POINT_DEPTH = 0.5 */
    double default_point_depth_variance;      /* Params.h:84   DEFAULT_POINT_DEPTH_VARIANCE = 100 */
    double default_point_homogenous_variance; /* Params.h:86   DEFAULT_POINT_HOMOGENOUS_VARIANCE = 1e-5 */
    uint32_t flags;
} ekfvio_params;

typedef struct ekfvio_batch ekfvio_batch;

const char* ekfvio_last_error(void);
/* Fills *p with the reference defaults (Params.h:83-86). */
void ekfvio_default_params(ekfvio_params* p);

/* ---- batched filter: TightlyCoupledEKF (include/ekf_vio/TightlyCoupledEKF.h:25-70) ---------- */

/* TightlyCoupledEKF::TightlyCoupledEKF + initializeBaseState (TightlyCoupledEKF.cpp:10-56) for
 * num_filters independent filters with room for max_features features each. */
int ekfvio_batch_create(ekfvio_batch** out, int device, int num_filters, int max_features, const ekfvio_params* params);
int ekfvio_batch_destroy(ekfvio_batch* b);
int ekfvio_batch_num_filters(const ekfvio_batch* b);
int ekfvio_batch_max_features(const ekfvio_batch* b);

/* initializeBaseState (TightlyCoupledEKF.cpp:23-56) on every filter; drops all features. */
int ekfvio_batch_reset(ekfvio_batch* b, void* stream);

/* addNewFeatures (TightlyCoupledEKF.cpp:58-94): filter f appends d_k[f] features whose metric
 * (u,v) are d_uv[f][0..k-1][0..1] (row stride kmax*2).  Exceeding max_features is an error
 * reported through status bit2 for that filter (nothing is appended to it). */
int ekfvio_batch_add_features(ekfvio_batch* b, const int* d_k, const double* d_uv, int kmax, void* stream);

/* For callers that capture these entry points into a CUDA graph (as ekfvio_vio_add_frame does) and replay it: a replay runs the
 * kernels but none of the host-side bookkeeping of the calls it stands for.  Tell the batch how many Sigma buffer flips the
 * replayed sequence contains (process, update and remove_features flip once each); its copy-stream events are invalidated. */
int ekfvio_batch_graph_replayed(ekfvio_batch* b, int sigma_buffer_flips);
/* A captured graph has the batch's Sigma ping-pong buffer baked into its kernel arguments: it may only be replayed while the batch
 * is in the state it was captured in.  ekfvio_batch_graph_state returns that state as an opaque token (compare tokens for equality
 * before a replay; re-capture on mismatch); _restore puts back a token taken before a capture that failed half way. */
int ekfvio_batch_graph_state(const ekfvio_batch* b);
int ekfvio_batch_graph_state_restore(ekfvio_batch* b, int token);

/* Feature removal.  The reference only flags lost features (TightlyCoupledEKF.cpp:524-528, Feature::delete_flag)
 * and never deletes them; this marginalises features out of the state: mean entries and rows/columns of Sigma are deleted, the
 * survivors keep their order.  d_remove[F][nmax] non-zero = remove; NULL = remove the features whose delete flag the updates
 * have set.  Not called by anything in the library (the frame loop keeps the reference's behaviour); parity unpinned. */
int ekfvio_batch_remove_features(ekfvio_batch* b, const uint8_t* d_remove, void* stream);

/* process(dt) (TightlyCoupledEKF.cpp:96-121): numericallyLinearizeProcess, convolveFeature on
 * every feature, convolveBaseState, Sigma = F Sigma F' + Q, prune.  d_dt[F]. */
int ekfvio_batch_process(ekfvio_batch* b, const double* d_dt, void* stream);
/* Same with one dt for all filters. */
int ekfvio_batch_process_dt(ekfvio_batch* b, double dt, void* stream);

/* updateWithFeaturePositions (TightlyCoupledEKF.cpp:475-628).  d_z[F][nmax][2] metric
 * measurements, d_R[F][nmax][4] row-major 2x2 covariances, d_pass[F][nmax] (0/1).  Entries of
 * features >= nfeat[f] are ignored. */
int ekfvio_batch_update(ekfvio_batch* b, const double* d_z, const double* d_R, const uint8_t* d_pass, void* stream);

/* numericallyLinearizeProcess (TightlyCoupledEKF.cpp:176-325): writes the dense N x N Jacobian
 * of filter f to d_F[f][Nmax][Nmax] (row-major, leading dimension Nmax) without changing the
 * state — except the convolveFeature cache, exactly as in the reference. */
int ekfvio_batch_linearize(ekfvio_batch* b, const double* d_dt, double* d_F, void* stream);
/* convolveBaseState / convolveFeature (TightlyCoupledEKF.cpp:328-395, 397-460) as single evaluations for arbitrary arguments
 * (test_ekf.cpp:156-204 calls them that way).  HOST pointers, synchronous.  Filter `f` of the batch lends its dq_inv cache to
 * convolveFeature, which — like the reference's function-static cache — is keyed on omega only: a call with an unchanged omega
 * reuses the rotation of an earlier dt (E2; EKFVIO_FLAG_FRESH_DQ_CACHE recomputes it).  The filter's state is not modified. */
int ekfvio_batch_convolve_base_h(ekfvio_batch* b, int f, const double* h_base22, double dt, double* h_out22);
int ekfvio_batch_convolve_feature_h(ekfvio_batch* b, int f, const double* h_base22, const double* h_feat3, double dt, double* h_out3);

/* checkSigma (TightlyCoupledEKF.cpp:699-714): number of negative diagonal entries and
 * max |Sigma_ij - Sigma_ji| per filter.  Device outputs. */
int ekfvio_batch_check_sigma(ekfvio_batch* b, int* d_neg_diag, double* d_max_asym, void* stream);

/* State access (also the checkpoint/restore hook).  HOST pointers, any may be NULL.  Synchronous.
 * Layouts as in the header comment; P has leading dimension Nmax = 22 + 3*max_features. */
int ekfvio_batch_get_state(ekfvio_batch* b, double* h_mu, double* h_feat, double* h_P, int* h_nfeat, double* h_cache,
                           uint8_t* h_flags, double* h_klt_last, int* h_status);
/* The same for filters [first, first + count) only (outputs sized for `count` filters), plus the update path each filter took
 * in the last update (h_route, may be NULL: 0 symmetric fast path, 1 tiled Joseph form, 2 general LDL^T kernels). */
int ekfvio_batch_get_state_range(ekfvio_batch* b, int first, int count, double* h_mu, double* h_feat, double* h_P, int* h_nfeat,
                                 double* h_cache, uint8_t* h_flags, double* h_klt_last, int* h_status, int* h_route);
int ekfvio_batch_set_state(ekfvio_batch* b, const double* h_mu, const double* h_feat, const double* h_P, const int* h_nfeat,
                           const double* h_cache, const uint8_t* h_flags, const double* h_klt_last);

/* Host-buffer convenience wrappers (the calls the C++ facade and the e2e benchmark use): the kernels are
 * enqueued on `stream` and the call returns after the work is enqueued.  The host->device copies run on a
 * private copy stream that the kernels wait for — NOT behind earlier work of `stream`, so that the upload of a
 * step's measurements overlaps the process() still running there.  Host buffers are therefore read as they are
 * at call time: they must already hold their final contents when the function is called (as the reference's
 * by-value std::vector arguments do), and page-locked ones (cudaHostAlloc / cudaHostRegister), which are DMA'd
 * from where they are, must stay unchanged until the stream has passed this call.  Pageable inputs are staged
 * through the batch's pinned buffers (consumed before return). */
int ekfvio_batch_add_features_h(ekfvio_batch* b, const int* h_k, const double* h_uv, int kmax, void* stream);
int ekfvio_batch_update_h(ekfvio_batch* b, const double* h_z, const double* h_R, const uint8_t* h_pass, void* stream);
/* numericallyLinearizeProcess with one dt for all filters, Jacobians to HOST memory
 * h_F[F][Nmax][Nmax]; checkSigma with HOST outputs.  Synchronous (facade and test use). */
int ekfvio_batch_linearize_h(ekfvio_batch* b, double dt, double* h_F);
int ekfvio_batch_check_sigma_h(ekfvio_batch* b, int* h_neg_diag, double* h_max_asym);
/* Copies mu (F x 22) and feat (F x nmax x 3) to host after the enqueued work; synchronises. */
int ekfvio_batch_read_mu_h(ekfvio_batch* b, double* h_mu, double* h_feat, void* stream);

/* Device pointers of the live state, for zero-copy consumers (valid until destroy; P points at
 * the current buffer and may change after each process/update/remove_features — query again.  Between a process() and the
 * update() that follows, a batch on the reduced update path holds the feature rows of Sigma only up to the diagonal; this
 * call (like get_state and check_sigma) completes the matrix first, on the stream of that process()). */
typedef struct ekfvio_batch_view {
    double* d_mu;       /* [F][22] */
    double* d_feat;     /* [F][nmax][3] */
    double* d_P;        /* [F][ldP][ldP] */
    int* d_nfeat;       /* [F] */
    int* d_status;      /* [F] */
    int ldP;            /* leading dimension of P (>= Nmax, multiple of 8) */
    int num_filters, max_features;
    double* d_klt_last; /* [F][nmax][2] last KLT result per feature (Feature::last_result_from_klt_tracker) */
} ekfvio_batch_view;
int ekfvio_batch_get_view(ekfvio_batch* b, ekfvio_batch_view* view);

/* Number of kernel launches this library has issued on behalf of `b` since creation. */
long long ekfvio_batch_launch_count(const ekfvio_batch* b);

/* Per-kernel device timing (CUDA events on the launching stream) for the roofline report.
 * Slots: 0 process, 1 gain part 1 (measurement map, S, factorisation; the whole gain on the
 * general path), 2 covariance (Joseph) update, 3 gain part 2 (K, W, state update; tiled path).
 * get_timing synchronises, then returns accumulated milliseconds and launch counts (8 slots). */
int ekfvio_batch_enable_timing(ekfvio_batch* b, int on);
int ekfvio_batch_get_timing(ekfvio_batch* b, double* ms8, long long* count8);

/* Measures this GPU's FP64 peak with register-resident loops (about 50 ms each): the
 * DMMA.8x8x4 rate and the DFMA rate, in TFLOP/s.  The roofline denominator for the EKF. */
int ekfvio_measure_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops);

/* Monte-Carlo error statistics (no reference counterpart; north_star): per-filter squared
 * position / velocity error against ground truth accumulated on device into d_acc[8] =
 * {sum_e2_pos, sum_e2_vel, sum_e2_quat, count, max_e2_pos, 0, 0, 0}; the caller reduces d_acc
 * across ranks with one NCCL all-reduce (sum for 0-3). */
int ekfvio_batch_accumulate_errors(ekfvio_batch* b, const double* d_truth_mu /*[F][22]*/, double* d_acc, void* stream);

/* ---- pyramidal KLT tracker: KLTTracker (include/ekf_vio/KLTTracker.h:25-99) ----------------- */

typedef struct ekfvio_klt_params {
    int window_size;      /* Params.h:104  WINDOW_SIZE = 21  (odd, <= 31) */
    int max_pyramid_level;/* Params.h:103  MAX_PYRAMID_LEVEL = 3 */
    int max_iterations;   /* KLTTracker.cpp:63   30 */
    double epsilon;       /* KLTTracker.cpp:63   0.01 */
    double min_eigen;     /* Params.h:36   KLT_MIN_EIGEN = 1e-4 */
    int kill_pad;         /* Params.h:33   KILL_PAD = 11 */
    int use_initial_flow; /* KLTTracker.cpp:64   OPTFLOW_USE_INITIAL_FLOW -> 1 */
} ekfvio_klt_params;

typedef struct ekfvio_klt ekfvio_klt;

void ekfvio_klt_default_params(ekfvio_klt_params* p);

/* A tracker for up to max_batch image pairs of width x height 8-bit pixels and up to max_points
 * points per pair.  It owns `num_slots` pyramid slots, each holding max_batch pyramids. */
int ekfvio_klt_create(ekfvio_klt** out, int device, int width, int height, int max_batch, int max_points, int num_slots,
                      const ekfvio_klt_params* params);
int ekfvio_klt_destroy(ekfvio_klt* k);
/* Number of pyramid levels actually used (OpenCV stops when the next level would be <= window). */
int ekfvio_klt_num_levels(const ekfvio_klt* k);

/* cv::buildOpticalFlowPyramid (+ calcScharrDeriv when with_derivs) as called inside
 * cv::calcOpticalFlowPyrLK (KLTTracker.cpp:61): d_imgs[batch][height][pitch] 8-bit gray.
 * Fills pyramid slot `slot`. */
int ekfvio_klt_build_pyramid(ekfvio_klt* k, int slot, const uint8_t* d_imgs, int pitch, int batch, int with_derivs, void* stream);

/* Both pyramids of a frame pair in one pass, as cv::calcOpticalFlowPyrLK builds them per call
 * (KLTTracker.cpp:61): prev_slot gets intensities + derivatives, next_slot intensities (and
 * derivatives if next_with_derivs, so it can serve as the previous frame of the next call).
 * The two slots share each level's kernel launch.  NULL image pointers mean "level 0 is already
 * in the slot". */
int ekfvio_klt_build_pyramid_pair(ekfvio_klt* k, int prev_slot, const uint8_t* d_prev, int next_slot, const uint8_t* d_next, int pitch,
                                  int batch, int next_with_derivs, void* stream);
/* The same without the pass-through copy of level 0: the slots reference the caller's image batches, which must stay alive and
 * unmodified until the last ekfvio_klt_track on these slots (cv::calcOpticalFlowPyrLK's own contract: the images belong to the
 * caller for the duration of the call, KLTTracker.cpp:61-64).  Saves 2 x width x height bytes of HBM writes per pair. */
int ekfvio_klt_build_pyramid_pair_ref(ekfvio_klt* k, int prev_slot, const uint8_t* d_prev, int next_slot, const uint8_t* d_next, int pitch,
                                      int batch, int next_with_derivs, void* stream);

/* LKTrackerInvoker over all levels (cv::calcOpticalFlowPyrLK, KLTTracker.cpp:61-64) between
 * pyramid slots prev_slot (needs derivatives) and next_slot.  d_prev_pts[batch][max_points][2]
 * pixel coordinates, d_next_pts same shape: in = initial flow (if use_initial_flow), out =
 * result.  d_status[batch][max_points] (1 = tracked), d_err[batch][max_points] (may be NULL),
 * d_npts[batch] points per image pair. */
int ekfvio_klt_track(ekfvio_klt* k, int prev_slot, int next_slot, const float* d_prev_pts, float* d_next_pts, uint8_t* d_status,
                     float* d_err, const int* d_npts, int batch, void* stream);

/* KLTTracker.cpp:72-92 epilogue + Feature::pixel2Metric (Feature.h:60-62, with the reference's
 * linear-index use of K — E1): passed = status==1 && inside kill-pad; cov = 1e-5*I scaled by
 * 1/fx^2, 1/fy^2; measured = ((x-K(2))/K(0), (y-K(5))/K(4)).  d_K9[batch][9] column-major float
 * 3x3.  d_measured entries of failed points are left untouched, as in the reference. */
int ekfvio_klt_postprocess(ekfvio_klt* k, const float* d_next_pts, const uint8_t* d_status, const int* d_npts, const float* d_K9,
                           int batch, float* d_measured, float* d_cov, uint8_t* d_passed, void* stream);

/* KLTTracker::estimateUncertaintySampleBased (KLTTracker.cpp:111-175; dead code in the reference, provided for completeness):
 * per feature a 2x2 covariance in pixel units from a 5x5 reference patch (cv::getRectSubPix) compared with 25 shifted patches of
 * the current frame.  d_cov[batch][max_points][4] row-major.  Parity: OpenCV's getRectSubPix (cv2 4.13) + the reference's loop
 * restated in numpy (tests/test_gpu_klt.py); the reference has no test for it. */
int ekfvio_klt_sample_uncertainty(int device, const uint8_t* d_ref_imgs, const uint8_t* d_cur_imgs, int width, int height, int pitch, int batch,
                                  const float* d_ref_pts, const float* d_pts, const int* d_npts, int max_points, float* d_cov, void* stream);
int ekfvio_klt_sample_uncertainty_h(int device, const uint8_t* h_ref_img, const uint8_t* h_cur_img, int width, int height, int pitch,
                                    const float* h_ref_pts, const float* h_pts, int n, float* h_cov);

/* Host-buffer convenience: build both pyramids from host images, track, post-process, copy the
 * results back; synchronous.  The KLTTracker facade's findNewFeaturePositions is this call. */
int ekfvio_klt_track_pair_h(ekfvio_klt* k, const uint8_t* h_prev, const uint8_t* h_next, int pitch, int batch, const float* h_prev_pts,
                            float* h_next_pts, uint8_t* h_status, float* h_err, const int* h_npts, void* stream);
/* The same inside a sequence: the previous frame's pyramid is still on the device from the last ekfvio_klt_track_pair_h /
 * _track_next_h call, so only the new frame is uploaded (half the host->device traffic) and built — with derivatives, since it is
 * the previous frame of the next call.  This is KLTTracker::findNewFeaturePositions as EKFVIO::addFrame drives it frame after
 * frame (EKFVIO.cpp:201-217).  Arguments as ekfvio_klt_track_pair_h without h_prev. */
int ekfvio_klt_track_next_h(ekfvio_klt* k, const uint8_t* h_next, int pitch, int batch, const float* h_prev_pts, float* h_next_pts,
                            uint8_t* h_status, float* h_err, const int* h_npts, void* stream);

/* Reads back level `level` of image `img` in slot `slot` (tests): h_img[h_l][w_l] and, if not
 * NULL and the slot has derivatives, h_deriv[h_l][w_l][2] int16.  Synchronous. */
int ekfvio_klt_read_level(ekfvio_klt* k, int slot, int img, int level, uint8_t* h_img, int16_t* h_deriv, int* w_out, int* h_out);

long long ekfvio_klt_launch_count(const ekfvio_klt* k);
/* Timing slots: 0..3 pyramid/Scharr kernel of level 0..3 (slot 3 also collects deeper levels),
 * 4 track kernel, 5 post-process. */
int ekfvio_klt_enable_timing(ekfvio_klt* k, int on);
int ekfvio_klt_get_timing(ekfvio_klt* k, double* ms8, long long* count8);

/* ------------------------------------------------------------------------------------------------
 * Feature replenishment — EKFVIO::replenishFeatures (EKFVIO.cpp:224-311), the caller of
 * TightlyCoupledEKF::addNewFeatures: cv::FAST(img, kp, FAST_THRESHOLD, true) (:242), the check image
 * with cv::circle(.., MIN_NEW_FEATURE_DIST, 255, -1) around every feature already in the state
 * (:255-260) and the greedy scan over the keypoints in detector order (:262-305).  Bit-identical to
 * OpenCV: keypoint set, order (rows top to bottom, x ascending) and response, filled-circle raster.
 * FAST_BLUR_SIGMA is 0 in the reference's defaults (Params.h:26), so no blur stage.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ekfvio_fast ekfvio_fast;

/* max_keypoints: capacity of the keypoint list per image (more are counted but not stored). */
int ekfvio_fast_create(ekfvio_fast** out, int device, int width, int height, int max_batch, int max_keypoints);
int ekfvio_fast_destroy(ekfvio_fast* f);

/* cv::FAST with the 9/16 pattern for a batch of 8-bit images d_imgs[batch][height][pitch].
 * d_kp_xy[batch][max_keypoints][2] int16 (x, y); d_response[batch][max_keypoints] (cv::KeyPoint::response,
 * may be NULL); d_count[batch] = number of keypoints found (the first max_keypoints are stored). */
int ekfvio_fast_detect(ekfvio_fast* f, const uint8_t* d_imgs, int pitch, int batch, int threshold, int nonmax, short* d_kp_xy, int* d_response,
                       int* d_count, void* stream);

/* The greedy scan: d_existing_px[batch][max_existing][2] float pixel positions of the features in the
 * state (Feature::getPixel, rounded like cv::Point), d_n_existing[batch] (both may be NULL),
 * d_needed[batch] = NUM_FEATURES - features.size().  Outputs d_new_px[batch][max_new][2] int16,
 * d_new_metric[batch][max_new][2] = Feature::pixel2Metric with d_K9[batch][9] column-major (E1
 * semantics; both may be NULL), d_n_new[batch]. */
int ekfvio_fast_select(ekfvio_fast* f, const short* d_kp_xy, const int* d_count, const float* d_existing_px, const int* d_n_existing,
                       int max_existing, const int* d_needed, int min_dist, int kill_pad, const float* d_K9, short* d_new_px,
                       float* d_new_metric, int* d_n_new, int max_new, int batch, void* stream);

/* Host-buffer convenience for the whole of replenishFeatures: upload, detect (non-max suppression
 * on), select, download; synchronous.  h_kp_xy / h_count (may be NULL) return the detector output. */
int ekfvio_fast_replenish_h(ekfvio_fast* f, const uint8_t* h_imgs, int pitch, int batch, int threshold, const float* h_existing_px,
                            const int* h_n_existing, int max_existing, const int* h_needed, int min_dist, int kill_pad, const float* h_K9,
                            short* h_new_px, float* h_new_metric, int* h_n_new, int max_new, short* h_kp_xy, int* h_count, void* stream);

long long ekfvio_fast_launch_count(const ekfvio_fast* f);

/* ------------------------------------------------------------------------------------------------
 * Frame loop — EKFVIO::addFrame (EKFVIO.cpp:139-196) for one new frame of each of S independent
 * sequences, entirely on the device: first frame -> replenishFeatures; afterwards process(dt) (:163),
 * updateStateWithNewImage (:169 = KLTTracker::findNewFeaturePositions with prev = metric2Pixel(last KLT
 * result), initial flow = Feature::getPixel of the predicted state, :201-217, then
 * updateWithFeaturePositions), replenishFeatures (:172).  Every frame's pyramid is built once (with
 * derivatives) and serves as "previous" for the next frame.  The ROS glue around it (publishers,
 * tf, frame buffer bookkeeping) stays with the caller.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ekfvio_vio ekfvio_vio;
typedef struct ekfvio_vio_params {
    int num_features;          /* NUM_FEATURES, Params.h:46 (100) — also the filters' feature capacity */
    int fast_threshold;        /* FAST_THRESHOLD, Params.h:24 (50) */
    int min_new_feature_dist;  /* MIN_NEW_FEATURE_DIST, Params.h:43 (30) */
    int remove_lost_features;  /* 0 (reference behaviour: lost features stay in the state for ever); 1: features whose track was lost in
                                * this frame's update are marginalised out (ekfvio_batch_remove_features) before replenishment */
    int use_cuda_graph;        /* 1: after the first frames the per-frame launch sequence is captured once per pyramid-slot
                                * parity and replayed as a CUDA graph (one launch advances all sequences by a frame) */
} ekfvio_vio_params;
void ekfvio_vio_default_params(ekfvio_vio_params* p);
int ekfvio_vio_create(ekfvio_vio** out, int device, int num_sequences, int width, int height, const ekfvio_params* ekf_params,
                      const ekfvio_klt_params* klt_params, const ekfvio_vio_params* vio_params);
int ekfvio_vio_destroy(ekfvio_vio* v);
/* d_frames[S][height][pitch] 8-bit frames at the working resolution (Frame::Frame's resize: ekfvio_frame_resize),
 * d_K9[S][9] the frames' scaled K as column-major float 3x3, d_dt[S] seconds since the previous frame
 * (ignored for the first frame).  Asynchronous on `stream`. */
int ekfvio_vio_add_frame(ekfvio_vio* v, const uint8_t* d_frames, int pitch, const float* d_K9, const double* d_dt, void* stream);
/* The filters (state read-out, checkpointing) and the number of frames added so far. */
ekfvio_batch* ekfvio_vio_filters(ekfvio_vio* v);
int ekfvio_vio_frame_count(const ekfvio_vio* v);
long long ekfvio_vio_launch_count(const ekfvio_vio* v);

/* Frame::Frame (Frame.cpp:15-21): cv::resize(img, scaled, Size(cols / inv_scale, rows / inv_scale)), default
 * INTER_LINEAR, for a batch of 8-bit images d_src[batch][src_height][src_pitch] ->
 * d_dst[batch][src_height / inv_scale][dst_pitch]; bit-identical to OpenCV (area-fast path at exactly 2x,
 * 11-bit fixed-point bilinear otherwise).  The K scaling of Frame.cpp:26-30 is four host divisions and stays
 * with the caller.  Asynchronous on `stream`. */
int ekfvio_frame_resize(const uint8_t* d_src, int src_pitch, int src_width, int src_height, int batch, int inv_scale, uint8_t* d_dst,
                        int dst_pitch, void* stream);
/* The same with host buffers (upload, resize, download; synchronous) — what the Frame facade's resizing constructor calls. */
int ekfvio_frame_resize_h(const uint8_t* h_src, int src_pitch, int src_width, int src_height, int batch, int inv_scale, uint8_t* h_dst,
                          int dst_pitch);

/* ---- Monte-Carlo statistics over the GPUs of a box (the only collective) ------------------------------------
 * Independent filter / sequence shards never exchange data; the error accumulators that ekfvio_batch_accumulate_errors fills
 * (the reference has no counterpart: analyzeEKFSimulation.cpp:86-99 inspects one filter by eye) are summed across ranks with
 * one ncclAllReduce(ncclDouble, ncclSum) over NVLink.  One process per GPU: rank 0 obtains an id, the host application hands
 * its 128 bytes to the other ranks by any means (bench.py: torch.distributed), every rank creates its communicator. */
#define EKFVIO_COMM_ID_BYTES 128
typedef struct ekfvio_comm ekfvio_comm;
int ekfvio_comm_unique_id(unsigned char* id128);
int ekfvio_comm_create(ekfvio_comm** out, int device, int nranks, int rank, const unsigned char* id128);
int ekfvio_comm_destroy(ekfvio_comm* c);
int ekfvio_comm_size(const ekfvio_comm* c);   /* ranks the communicator really spans (ncclCommCount) */
/* In-place sum of d_buf[0..n) (device, FP64) over all ranks, enqueued on `stream`. */
int ekfvio_stats_allreduce(ekfvio_comm* c, double* d_buf, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EKFVIO_C_H_ */
