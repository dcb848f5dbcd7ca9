/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the arithmetic behind KLTTracker::findNewFeaturePositionsOpenCV
 * (reference: include/ekf_vio/KLTTracker.cpp:40-95).  The tracker there is a single call to
 * cv::calcOpticalFlowPyrLK (KLTTracker.cpp:61-64); OpenCV is a third-party dependency that is
 * NOT vendored under /root/reference and whose version the reference does not pin
 * (CMakeLists.txt:31, find_package(OpenCV REQUIRED)).  This file restates OpenCV's published
 * algorithm (modules/video/src/lkpyramid.cpp: buildOpticalFlowPyramid, calcScharrDeriv,
 * LKTrackerInvoker; modules/imgproc/src/pyramids.cpp: pyrDown for 8-bit) and is pinned against
 * the OpenCV 4.13.0 Python wheel of this image (tests/test_oracle_klt.py and the golden vectors
 * in tests/golden/ produced by tests/golden/make_klt_golden.py with cv2 itself).
 * The reference holds no test vectors for this path (test/klt_test.cpp never tracks).
 *
 * Differences to OpenCV that remain by construction: OpenCV accumulates the window sums
 * (A11, A12, A22, b1, b2) in float SIMD lanes; here they are accumulated exactly in int64 and
 * converted once.  Measured effect: status identical, positions within 2e-4 px.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KLT_MAX_LEVELS 16

static int reflect101(int p, int len) {
    /* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* cv::pyrDown on CV_8UC1: separable [1 4 6 4 1], integer, (sum + 128) >> 8, REFLECT_101. */
void klt_oracle_pyrdown_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dstride) {
    static const int k[5] = {1, 4, 6, 4, 1};
    int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    for (int y = 0; y < dh; ++y) {
        for (int x = 0; x < dw; ++x) {
            int sum = 0;
            for (int j = 0; j < 5; ++j) {
                int sy = reflect101(2 * y + j - 2, sh);
                int row = 0;
                for (int i = 0; i < 5; ++i) row += k[i] * src[(size_t)sy * sstride + reflect101(2 * x + i - 2, sw)];
                sum += k[j] * row;
            }
            dst[(size_t)y * dstride + x] = (uint8_t)((sum + 128) >> 8);
        }
    }
}

/* cv::detail::calcScharrDeriv: interleaved int16 (Ix, Iy), 3-10-3 Scharr, REFLECT_101. */
void klt_oracle_scharr_s16(const uint8_t* src, int w, int h, int stride, int16_t* dst /* h x w x 2 */) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * stride;
        const uint8_t* r1 = src + (size_t)y * stride;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * stride;
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int s_m = 3 * r0[xm] + 10 * r1[xm] + 3 * r2[xm];
            int s_p = 3 * r0[xp] + 10 * r1[xp] + 3 * r2[xp];
            int d_m = r2[xm] - r0[xm], d_c = r2[x] - r0[x], d_p = r2[xp] - r0[xp];
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(s_p - s_m);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)(3 * d_m + 10 * d_c + 3 * d_p);
        }
    }
}

/* Level sizes as cv::buildOpticalFlowPyramid computes them; returns the effective max level. */
int klt_oracle_level_sizes(int w, int h, int win, int max_level, int* lw, int* lh) {
    int level = 0;
    lw[0] = w; lh[0] = h;
    for (level = 0; level < max_level; ++level) {
        int nw = (lw[level] + 1) / 2, nh = (lh[level] + 1) / 2;
        if (nw <= win || nh <= win) return level;
        lw[level + 1] = nw; lh[level + 1] = nh;
    }
    return max_level;
}

typedef struct {
    int levels;                       /* effective max level + 1 */
    int w[KLT_MAX_LEVELS], h[KLT_MAX_LEVELS];
    uint8_t* img[KLT_MAX_LEVELS];     /* unpadded, stride = w */
    int16_t* deriv[KLT_MAX_LEVELS];   /* interleaved, may be NULL */
} klt_pyr;

static void pyr_build(klt_pyr* p, const uint8_t* img, int w, int h, int stride, int win, int max_level, int with_deriv) {
    int ml = klt_oracle_level_sizes(w, h, win, max_level, p->w, p->h);
    p->levels = ml + 1;
    for (int l = 0; l <= ml; ++l) {
        p->img[l] = (uint8_t*)malloc((size_t)p->w[l] * p->h[l]);
        if (l == 0) for (int y = 0; y < h; ++y) memcpy(p->img[0] + (size_t)y * w, img + (size_t)y * stride, (size_t)w);
        else klt_oracle_pyrdown_u8(p->img[l - 1], p->w[l - 1], p->h[l - 1], p->w[l - 1], p->img[l], p->w[l]);
        p->deriv[l] = NULL;
        if (with_deriv) {
            p->deriv[l] = (int16_t*)malloc((size_t)p->w[l] * p->h[l] * 2 * sizeof(int16_t));
            klt_oracle_scharr_s16(p->img[l], p->w[l], p->h[l], p->w[l], p->deriv[l]);
        }
    }
}
static void pyr_free(klt_pyr* p) {
    for (int l = 0; l < p->levels; ++l) { free(p->img[l]); free(p->deriv[l]); }
}

/* Pixel fetches with the borders OpenCV's pyramid carries: REFLECT_101 for intensities
 * (copyMakeBorder(..., pyrBorder)), constant zero for derivatives (BORDER_CONSTANT). */
static inline int pixI(const uint8_t* im, int w, int h, int x, int y) { return im[(size_t)reflect101(y, h) * w + reflect101(x, w)]; }
static inline int pixD(const int16_t* d, int w, int h, int x, int y, int c) {
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0;
    return d[((size_t)y * w + x) * 2 + c];
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }   /* cvRound: round half to even */
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

void klt_oracle_weights(float a, float b, int* w) {
    w[0] = cv_round_f((1.f - a) * (1.f - b) * (1 << 14));
    w[1] = cv_round_f(a * (1.f - b) * (1 << 14));
    w[2] = cv_round_f((1.f - a) * b * (1 << 14));
    w[3] = (1 << 14) - w[0] - w[1] - w[2];
}

/* One point on one level: LKTrackerInvoker::operator() body.  next is in/out (level coords). */
static void track_point_level(const klt_pyr* I, const klt_pyr* J, int level, int max_level_eff, int use_initial,
                              float prev_x, float prev_y, float* next_xy /* stored nextPts */, uint8_t* status, float* err,
                              int win, int max_count, double epsilon, double min_eig_thr, int get_min_eig, int* iters_out) {
    const float half = (win - 1) * 0.5f;
    const int w = I->w[level], h = I->h[level];
    const float scale = (float)(1. / (1 << level));
    float px = prev_x * scale, py = prev_y * scale;
    float nx, ny;
    if (level == max_level_eff) {
        if (use_initial) { nx = next_xy[0] * scale; ny = next_xy[1] * scale; }
        else { nx = px; ny = py; }
    } else { nx = next_xy[0] * 2.f; ny = next_xy[1] * 2.f; }
    next_xy[0] = nx; next_xy[1] = ny;

    px -= half; py -= half;
    int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) {
        if (level == 0) { *status = 0; if (err) *err = 0; }
        return;
    }
    float a = px - ipx, b = py - ipy;
    int iw[4];
    klt_oracle_weights(a, b, iw);
    const float FLT_SCALE = 1.f / (1 << 20);
    short* Iw = (short*)malloc(sizeof(short) * win * win * 3);
    short* dIw = Iw + win * win;
    int64_t iA11 = 0, iA12 = 0, iA22 = 0;
    for (int y = 0; y < win; ++y)
        for (int x = 0; x < win; ++x) {
            int X = ipx + x, Y = ipy + y;
            int ival = DESCALE(pixI(I->img[level], w, h, X, Y) * iw[0] + pixI(I->img[level], w, h, X + 1, Y) * iw[1] +
                               pixI(I->img[level], w, h, X, Y + 1) * iw[2] + pixI(I->img[level], w, h, X + 1, Y + 1) * iw[3], 14 - 5);
            int ixval = DESCALE(pixD(I->deriv[level], w, h, X, Y, 0) * iw[0] + pixD(I->deriv[level], w, h, X + 1, Y, 0) * iw[1] +
                                pixD(I->deriv[level], w, h, X, Y + 1, 0) * iw[2] + pixD(I->deriv[level], w, h, X + 1, Y + 1, 0) * iw[3], 14);
            int iyval = DESCALE(pixD(I->deriv[level], w, h, X, Y, 1) * iw[0] + pixD(I->deriv[level], w, h, X + 1, Y, 1) * iw[1] +
                                pixD(I->deriv[level], w, h, X, Y + 1, 1) * iw[2] + pixD(I->deriv[level], w, h, X + 1, Y + 1, 1) * iw[3], 14);
            Iw[y * win + x] = (short)ival;
            dIw[(y * win + x) * 2] = (short)ixval;
            dIw[(y * win + x) * 2 + 1] = (short)iyval;
            iA11 += (int64_t)ixval * ixval; iA12 += (int64_t)ixval * iyval; iA22 += (int64_t)iyval * iyval;
        }
    float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (err && get_min_eig) *err = minEig;
    if ((double)minEig < min_eig_thr || D < FLT_EPSILON) {
        if (level == 0) *status = 0;
        free(Iw);
        return;
    }
    D = 1.f / D;
    nx -= half; ny -= half;
    float pdx = 0, pdy = 0;
    const int jw = J->w[level], jh = J->h[level];
    int j;
    for (j = 0; j < max_count; ++j) {
        int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= jw || iny < -win || iny >= jh) {
            if (level == 0) *status = 0;
            break;
        }
        a = nx - inx; b = ny - iny;
        klt_oracle_weights(a, b, iw);
        int64_t ib1 = 0, ib2 = 0;
        for (int y = 0; y < win; ++y)
            for (int x = 0; x < win; ++x) {
                int X = inx + x, Y = iny + y;
                int diff = DESCALE(pixI(J->img[level], jw, jh, X, Y) * iw[0] + pixI(J->img[level], jw, jh, X + 1, Y) * iw[1] +
                                   pixI(J->img[level], jw, jh, X, Y + 1) * iw[2] + pixI(J->img[level], jw, jh, X + 1, Y + 1) * iw[3], 14 - 5) -
                           Iw[y * win + x];
                ib1 += (int64_t)diff * dIw[(y * win + x) * 2];
                ib2 += (int64_t)diff * dIw[(y * win + x) * 2 + 1];
            }
        float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
        float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
        nx += dx; ny += dy;
        next_xy[0] = nx + half; next_xy[1] = ny + half;
        if (iters_out) ++*iters_out;
        if ((double)dx * dx + (double)dy * dy <= epsilon) break;
        if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
            next_xy[0] -= dx * 0.5f; next_xy[1] -= dy * 0.5f;
            break;
        }
        pdx = dx; pdy = dy;
    }
    if (*status && err && level == 0 && !get_min_eig) {
        float fx = next_xy[0] - half, fy = next_xy[1] - half;
        int inx = (int)floorf(fx), iny = (int)floorf(fy);
        if (inx < -win || inx >= jw || iny < -win || iny >= jh) {
            *status = 0;
            free(Iw);
            return;
        }
        float aa = fx - inx, bb = fy - iny;
        klt_oracle_weights(aa, bb, iw);
        float errval = 0.f;
        for (int y = 0; y < win; ++y)
            for (int x = 0; x < win; ++x) {
                int X = inx + x, Y = iny + y;
                int diff = DESCALE(pixI(J->img[level], jw, jh, X, Y) * iw[0] + pixI(J->img[level], jw, jh, X + 1, Y) * iw[1] +
                                   pixI(J->img[level], jw, jh, X, Y + 1) * iw[2] + pixI(J->img[level], jw, jh, X + 1, Y + 1) * iw[3], 14 - 5) -
                           Iw[y * win + x];
                errval += fabsf((float)diff);
            }
        *err = errval * 1.f / (32 * win * win);
    }
    free(Iw);
}

/* cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), max_level,
 * TermCriteria(COUNT+EPS, max_count, eps), flags, min_eig).  next_pts is in/out when
 * use_initial_flow != 0.  iters (may be NULL) receives the LK iteration count per point. */
void klt_oracle_calc_optical_flow(const uint8_t* prev, const uint8_t* next, int w, int h, int stride, int npts,
                                  const float* prev_pts, float* next_pts, uint8_t* status, float* err, int win, int max_level,
                                  int max_count, double eps, int use_initial_flow, int get_min_eig, double min_eig, int* iters) {
    klt_pyr I, J;
    /* The reference passes a cv::Mat for err (KLTTracker.cpp:46,61), so OpenCV's final level-0
     * bounds check (which can clear status) is always active: keep err non-NULL internally. */
    float* err_local = NULL;
    if (!err) { err_local = (float*)malloc(sizeof(float) * (npts > 0 ? npts : 1)); err = err_local; }
    pyr_build(&I, prev, w, h, stride, win, max_level, 1);
    pyr_build(&J, next, w, h, stride, win, max_level, 0);
    int ml = I.levels - 1;
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    double epsilon = eps * eps;
#pragma omp parallel for schedule(dynamic, 8)
    for (int p = 0; p < npts; ++p) {
        status[p] = 1;
        if (err) err[p] = 0;
        if (iters) iters[p] = 0;
        if (!use_initial_flow) { next_pts[2 * p] = prev_pts[2 * p]; next_pts[2 * p + 1] = prev_pts[2 * p + 1]; }
        for (int level = ml; level >= 0; --level)
            track_point_level(&I, &J, level, ml, use_initial_flow, prev_pts[2 * p], prev_pts[2 * p + 1], next_pts + 2 * p, status + p,
                              err ? err + p : NULL, win, max_count, epsilon, min_eig, get_min_eig, iters ? iters + p : NULL);
    }
    pyr_free(&I);
    pyr_free(&J);
    free(err_local);
}

/* Pyramid accessors for unit tests: writes level l of the pyramid of img (and its Scharr
 * derivatives if deriv != NULL) into caller buffers sized by klt_oracle_level_sizes. */
void klt_oracle_build_level(const uint8_t* img, int w, int h, int stride, int win, int max_level, int level, uint8_t* out,
                            int16_t* deriv) {
    klt_pyr P;
    pyr_build(&P, img, w, h, stride, win, max_level, deriv != NULL);
    if (level < P.levels) {
        memcpy(out, P.img[level], (size_t)P.w[level] * P.h[level]);
        if (deriv) memcpy(deriv, P.deriv[level], (size_t)P.w[level] * P.h[level] * 2 * sizeof(int16_t));
    }
    pyr_free(&P);
}

/* KLTTracker.cpp:72-92 + Feature.h:60-66 (E1: K is indexed linearly into a column-major 3x3, so
 * "K(2)" and "K(5)" are K(2,0) and K(2,1)).  K9 is the column-major 3x3 float matrix. */
void klt_oracle_postprocess(int npts, const float* next_pts, const uint8_t* status, int cols, int rows, const float* K9,
                            int kill_pad, float* measured /* n x 2 */, float* cov /* n x 4 */, uint8_t* passed) {
    for (int i = 0; i < npts; ++i) {
        float x = next_pts[2 * i], y = next_pts[2 * i + 1];
        if (status[i] == 1 && !(x < kill_pad || y < kill_pad || cols - x < kill_pad || rows - y < kill_pad)) {
            passed[i] = 1;
            float c[4] = {0.00001f, 0, 0, 0.00001f};
            float scale = (float)pow(1.0 / K9[0], 2);       /* K(0,0) */
            c[0] *= scale; c[1] *= scale;
            scale = (float)pow(1.0 / K9[4], 2);             /* K(1,1) */
            c[3] *= scale; c[2] *= scale;
            memcpy(cov + 4 * i, c, sizeof(c));
            measured[2 * i] = (x - K9[2]) / K9[0];
            measured[2 * i + 1] = (y - K9[5]) / K9[4];
        } else {
            passed[i] = 0;
            memset(cov + 4 * i, 0, 4 * sizeof(float));
            /* measured_positions[i] is left untouched by the reference */
        }
    }
}
