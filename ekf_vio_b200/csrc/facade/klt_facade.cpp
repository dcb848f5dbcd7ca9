// Host-side C++ facade: the reference's KLTTracker interface over the C ABI (ekfvio_c.h).
// findNewFeaturePositionsOpenCV follows KLTTracker.cpp:40-95 line by line with
// cv::calcOpticalFlowPyrLK replaced by the CUDA tracker.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../../include/ekf_vio/KLTTracker.h"
#include "../../../include/ekfvio_c.h"

#define KLT_ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ASSERTION FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::abort(); } } while (0)
#define KLT_CALL(x) do { if ((x) != 0) { std::fprintf(stderr, "ekfvio: %s failed: %s\n", #x, ekfvio_last_error()); std::abort(); } } while (0)

KLTTracker::KLTTracker() {}
KLTTracker::~KLTTracker() { if (dev_) ekfvio_klt_destroy(dev_); }

void KLTTracker::findNewFeaturePositions(const Frame& lf, const Frame& cf, const std::vector<Eigen::Vector2f>& previous_feature_positions,
                                         const std::list<Feature>& estimated_new_feature_positions, std::vector<Eigen::Vector2f>& measured_positions,
                                         std::vector<Eigen::Matrix2f>& estimated_uncertainty, std::vector<bool>& passed) {
    this->findNewFeaturePositionsOpenCV(lf, cf, previous_feature_positions, estimated_new_feature_positions, measured_positions, estimated_uncertainty, passed);
}

void KLTTracker::findNewFeaturePositionsOpenCV(const Frame& lf, const Frame& cf, const std::vector<Eigen::Vector2f>& previous_feature_positions,
                                               const std::list<Feature>& estimated_new_feature_positions, std::vector<Eigen::Vector2f>& measured_positions,
                                               std::vector<Eigen::Matrix2f>& estimated_uncertainty, std::vector<bool>& passed) {
    KLT_ASSERT(previous_feature_positions.size() == estimated_new_feature_positions.size());   // KLTTracker.cpp:51
    KLT_ASSERT(lf.img.rows == cf.img.rows && lf.img.cols == cf.img.cols && lf.img.data && cf.img.data);
    std::vector<cv::Point2f> prev_fts, new_fts;
    for (auto e : previous_feature_positions) prev_fts.push_back(Feature::metric2Pixel(lf, e));   // :53-55
    for (auto e : estimated_new_feature_positions) new_fts.push_back(e.getPixel(cf));             // :57-59

    const int n = (int)prev_fts.size();
    if (!dev_ || w_ != cf.img.cols || h_ != cf.img.rows || n > max_points_) {
        if (dev_) ekfvio_klt_destroy(dev_);
        ekfvio_klt_params p;
        ekfvio_klt_default_params(&p);
        p.window_size = WINDOW_SIZE; p.max_pyramid_level = MAX_PYRAMID_LEVEL; p.min_eigen = KLT_MIN_EIGEN; p.kill_pad = KILL_PAD;
        w_ = cf.img.cols; h_ = cf.img.rows; max_points_ = n > 256 ? n : 256;
        KLT_CALL(ekfvio_klt_create(&dev_, 0, w_, h_, 1, max_points_, 2, &p));
    }
    std::vector<float> pp((size_t)max_points_ * 2, 0.f), nn((size_t)max_points_ * 2, 0.f), err(max_points_, 0.f);
    std::vector<unsigned char> status(max_points_, 0);
    for (int i = 0; i < n; ++i) { pp[2 * i] = prev_fts[i].x; pp[2 * i + 1] = prev_fts[i].y; nn[2 * i] = new_fts[i].x; nn[2 * i + 1] = new_fts[i].y; }
    KLT_ASSERT(lf.img.step == cf.img.step);
    // cv::calcOpticalFlowPyrLK(lf.img, cf.img, prev_fts, new_fts, status, error, Size(WINDOW_SIZE, WINDOW_SIZE), MAX_PYRAMID_LEVEL,
    //                          TermCriteria(COUNT + EPS, 30, 0.01), OPTFLOW_USE_INITIAL_FLOW, KLT_MIN_EIGEN)      (:61-64)
    if (n > 0) KLT_CALL(ekfvio_klt_track_pair_h(dev_, lf.img.data, cf.img.data, (int)cf.img.step, 1, pp.data(), nn.data(), status.data(), err.data(), &n, nullptr));
    for (int i = 0; i < n; ++i) { new_fts[i].x = nn[2 * i]; new_fts[i].y = nn[2 * i + 1]; }
    last_new_fts = new_fts;
    last_status.assign(status.begin(), status.begin() + n);

    passed.resize(new_fts.size());                 // :68-70
    estimated_uncertainty.resize(new_fts.size());
    measured_positions.resize(new_fts.size());
    for (size_t i = 0; i < new_fts.size(); i++) {   // :72-92
        if (status.at(i) == 1 && !(new_fts[i].x < KILL_PAD || new_fts[i].y < KILL_PAD || cf.img.cols - new_fts[i].x < KILL_PAD || cf.img.rows - new_fts[i].y < KILL_PAD)) {
            passed[i] = true;
            estimated_uncertainty[i] = this->estimateUncertainty(cf, new_fts.at(i));
            float scale = (float)std::pow(1.0 / cf.K(0, 0), 2);
            estimated_uncertainty[i](0, 0) *= scale;
            estimated_uncertainty[i](0, 1) *= scale;
            scale = (float)std::pow(1.0 / cf.K(1, 1), 2);
            estimated_uncertainty[i](1, 1) *= scale;
            estimated_uncertainty[i](1, 0) *= scale;
            measured_positions[i] = Feature::pixel2Metric(cf, new_fts.at(i));
        } else {
            passed[i] = false;
            estimated_uncertainty[i] = Eigen::Matrix2f::Zero();
        }
    }
}

Eigen::Matrix2f KLTTracker::estimateUncertainty(const Frame& cf, cv::Point2f mu) {   // :100-106
    KLT_ASSERT(cf.img.rows || mu.x);
    Eigen::Matrix2f A;
    A(0, 0) = 0.00001f; A(0, 1) = 0; A(1, 0) = 0; A(1, 1) = 0.00001f;
    return A;
}

// KLTTracker.cpp:111-175 (dead code in the reference): evaluated on the device for this one feature
Eigen::Matrix2f KLTTracker::estimateUncertaintySampleBased(const Frame& lf, cv::Point2f mu_ref, const Frame& cf, cv::Point2f mu) {
    KLT_ASSERT(lf.img.data && cf.img.data && lf.img.cols == cf.img.cols && lf.img.rows == cf.img.rows && lf.img.step == cf.img.step);
    const float rp[2] = {mu_ref.x, mu_ref.y}, p[2] = {mu.x, mu.y};
    float c[4] = {0, 0, 0, 0};
    if (ekfvio_klt_sample_uncertainty_h(0, lf.img.data, cf.img.data, lf.img.cols, lf.img.rows, (int)lf.img.step, rp, p, 1, c)) {
        std::fprintf(stderr, "estimateUncertaintySampleBased: %s\n", ekfvio_last_error());
        std::abort();
    }
    Eigen::Matrix2f A;
    A(0, 0) = c[0]; A(0, 1) = c[1]; A(1, 0) = c[2]; A(1, 1) = c[3];
    return A;
}

// ---- Frame::Frame (Frame.cpp:15-41): resize on the device, scale K ---------------------------------------------------
Frame::Frame(int inv_scale, const cv::Mat& full_img, const double k[9], const std::vector<double>& d, ros::Time _t) : t(_t) {
    if (inv_scale <= 0 || !full_img.data || full_img.cols / inv_scale <= 0 || full_img.rows / inv_scale <= 0) {
        std::fprintf(stderr, "Frame: bad image or inverse scale\n");
        std::abort();
    }
    const int dw = full_img.cols / inv_scale, dh = full_img.rows / inv_scale;
    std::vector<uint8_t> scaled((size_t)dw * dh);
    if (ekfvio_frame_resize_h(full_img.data, (int)full_img.step, full_img.cols, full_img.rows, 1, inv_scale, scaled.data(), dw)) {
        std::fprintf(stderr, "Frame: %s\n", ekfvio_last_error());
        std::abort();
    }
    img = cv::Mat::copyOf(dh, dw, scaled.data(), (size_t)dw);
    K.setZero();
    K(0, 0) = (float)(k[0] / inv_scale);
    K(0, 2) = (float)(k[2] / inv_scale);
    K(1, 1) = (float)(k[4] / inv_scale);
    K(1, 2) = (float)(k[5] / inv_scale);
    K(2, 2) = 1.0f;
    for (int i = 0; i < 5 && i < (int)d.size(); ++i) D(0, i) = (float)d[i];
}
