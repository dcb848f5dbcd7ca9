/* Feature.h — one tracked point of the filter state (reference include/ekf_vio/Feature.h:37-95,
 * Feature.cpp:14-36), including E1: K is read with LINEAR indices into a column-major 3x3, so
 * "K(2)" and "K(5)" are K(2,0) and K(2,1) — zero — and the principal point is dropped. */
#ifndef EKFVIO_FEATURE_H_
#define EKFVIO_FEATURE_H_

#include "Frame.h"

class Feature {
private:
    Eigen::Vector3f mu;                            /* [u, v, 1/depth] */
    Eigen::Vector2f last_result_from_klt_tracker;
    bool delete_flag;

public:
    Feature() : delete_flag(false) {}
    Feature(Eigen::Vector2f homogenous, float depth) {   /* Feature.cpp:14-20 */
        last_result_from_klt_tracker = homogenous;
        mu(0) = homogenous.x();
        mu(1) = homogenous.y();
        mu(2) = (float)(1.0 / depth);
        delete_flag = false;
    }
    virtual ~Feature() {}

    Eigen::Vector2f getNormalizedPixel() { return Eigen::Vector2f(mu(0), mu(1)); }
    float getDepth() { return mu(2); }              /* the INVERSE depth, as in the reference (Feature.cpp:30-32) */
    cv::Point2f getPixel(const Frame& f) const { return cv::Point2f(f.K(0) * mu(0) + f.K(2), f.K(4) * mu(1) + f.K(5)); }

    static inline Eigen::Vector2f pixel2Metric(const Frame& f, const cv::Point2f px) {
        return Eigen::Vector2f((px.x - f.K(2)) / f.K(0), (px.y - f.K(5)) / f.K(4));
    }
    static inline cv::Point2f metric2Pixel(const Frame& f, const Eigen::Vector2f pos) {
        return cv::Point2f(pos.x() * f.K(0) + f.K(2), pos.y() * f.K(4) + f.K(5));
    }

    Eigen::Vector2f getLastResultFromKLTTracker() const { return last_result_from_klt_tracker; }
    void setLastResultFromKLTTracker(Eigen::Vector2f in) { last_result_from_klt_tracker = in; }
    bool flaggedForDeletion() const { return delete_flag; }
    void setDeleteFlag(bool in) { delete_flag = in; }
    void setNormalizedPixel(Eigen::Vector2f in) { mu(0) = in(0); mu(1) = in(1); }
    void setDepth(float in) { mu(2) = in; }
    void setMu(Eigen::Vector3f in) { mu = in; }
    Eigen::Vector3f& getMu() { return mu; }
    const Eigen::Vector3f& getMu() const { return mu; }
};

#endif
