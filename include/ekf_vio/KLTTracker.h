/* KLTTracker.h — the reference's tracker interface (include/ekf_vio/KLTTracker.h:25-99) backed by
 * the CUDA pyramidal Lucas-Kanade path through the C ABI.  findNewFeaturePositionsOpenCV does what
 * the reference's does (KLTTracker.cpp:40-95) with cv::calcOpticalFlowPyrLK replaced by
 * ekfvio_klt_track_pair_h (bit-compatible pyramid/derivatives, status bit-exact, positions
 * within 0.01 px). */
#ifndef EKFVIO_KLTTRACKER_H_
#define EKFVIO_KLTTRACKER_H_

#include <list>
#include <vector>

#include "Feature.h"
#include "Frame.h"
#include "Params.h"

struct ekfvio_klt;

class KLTTracker {
public:
    /* the reference's allocate-only scaffold (KLTTracker.h:29-83), kept for source compatibility */
    struct Pyramid {
        struct PyramidLevel {
            struct Pixel { bool set; float value; };
            std::vector<Pixel> image;
            int rows = 0, cols = 0;
            PyramidLevel() {}
            PyramidLevel(const int _rows, const int _cols) : image((size_t)_rows * _cols), rows(_rows), cols(_cols) {}
        };
        cv::Mat base_img;
        int level_count = 0;
        std::vector<PyramidLevel> levels;
        Pyramid(const int num_levels, cv::Mat image_ptr) : base_img(image_ptr), level_count(num_levels), levels(num_levels) {
            for (int i = 1; i <= level_count; i++) {
                if (i == 1) levels[i - 1] = PyramidLevel(base_img.rows / 2, base_img.cols / 2);
                else levels[i - 1] = PyramidLevel(levels[i - 2].rows / 2, levels[i - 2].cols / 2);
            }
        }
    };

    KLTTracker();
    virtual ~KLTTracker();
    KLTTracker(const KLTTracker&) = delete;
    KLTTracker& operator=(const KLTTracker&) = delete;

    void findNewFeaturePositions(const Frame& lf, const Frame& cf, const std::vector<Eigen::Vector2f>& previous_feature_positions,
                                 const std::list<Feature>& estimated_new_feature_positions, std::vector<Eigen::Vector2f>& measured_positions,
                                 std::vector<Eigen::Matrix2f>& estimated_uncertainty, std::vector<bool>& passed);
    void findNewFeaturePositionsOpenCV(const Frame& lf, const Frame& cf, const std::vector<Eigen::Vector2f>& previous_feature_positions,
                                       const std::list<Feature>& estimated_new_feature_positions, std::vector<Eigen::Vector2f>& measured_positions,
                                       std::vector<Eigen::Matrix2f>& estimated_uncertainty, std::vector<bool>& passed);
    Eigen::Matrix2f estimateUncertainty(const Frame& cf, cv::Point2f mu);
    /* dead code in the reference (KLTTracker.cpp:111-175, never called); not provided on the device */
    Eigen::Matrix2f estimateUncertaintySampleBased(const Frame& lf, cv::Point2f mu_ref, const Frame& cf, cv::Point2f mu);

    /* raw tracker outputs of the last call (pixels / OpenCV status), for tests */
    std::vector<cv::Point2f> last_new_fts;
    std::vector<unsigned char> last_status;

private:
    ekfvio_klt* dev_ = nullptr;
    int w_ = 0, h_ = 0, max_points_ = 0;
};

#endif
