/* TightlyCoupledEKF.h — the reference's filter interface (include/ekf_vio/TightlyCoupledEKF.h:25-70)
 * kept name for name, backed by the batched FP64 CUDA path through the C ABI (ekfvio_c.h) with a
 * batch of one filter.  Public members mirror the reference's and are refreshed after every
 * mutating call; values a caller writes into base_mu / features / Sigma (as test_ekf.cpp:170-199 and
 * jacobian_test.cpp:39-43 do) are pushed to the device before the next call.
 *
 * Differences a caller can see: arithmetic is FP64 on the device (the reference is float); Sigma is
 * a dense matrix type (the reference's SparseMatrix<float> is ~90 % dense after one step).
 */
#ifndef EKFVIO_TIGHTLYCOUPLEDEKF_H_
#define EKFVIO_TIGHTLYCOUPLEDEKF_H_

#define BASE_STATE_SIZE 22
#define SPARSE_THRESH 1e-8
#define SPARSE_EPS 1e-5

#include <list>
#include <vector>

#include "Feature.h"

struct ekfvio_batch;

class TightlyCoupledEKF {
public:
    TightlyCoupledEKF();
    TightlyCoupledEKF(const TightlyCoupledEKF& o);
    TightlyCoupledEKF& operator=(const TightlyCoupledEKF& o);
    ~TightlyCoupledEKF();

    Eigen::Matrix<float, BASE_STATE_SIZE, 1> base_mu;
    std::list<Feature> features;
    Eigen::SparseMatrix<float> Sigma;
    ros::Time t;

    void initializeBaseState();
    void addNewFeatures(std::vector<Eigen::Vector2f> new_homogenous_features);
    std::vector<Eigen::Vector2f> previousFeaturePositionVector();
    void process(float dt);
    Eigen::SparseMatrix<float> generateProcessNoise(float dt);
    Eigen::Matrix<float, BASE_STATE_SIZE, 1> convolveBaseState(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& last, float dt);
    Eigen::Vector3f convolveFeature(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& base_state, Eigen::Vector3f& feature_state, float dt);
    Eigen::SparseMatrix<float> numericallyLinearizeProcess(Eigen::Matrix<float, BASE_STATE_SIZE, 1>& base_mu, std::list<Feature>& features, float dt);
    void updateWithFeaturePositions(std::vector<Eigen::Vector2f> measured_positions, std::vector<Eigen::Matrix2f> estimated_covariance,
                                    std::vector<bool> pass);
    Eigen::SparseMatrix<float> formFeatureMeasurementMap(std::vector<bool> measured);
    Eigen::Matrix2f getFeatureHomogenousCovariance(int index);
    float getFeatureDepthVariance(int index);
    void setFeatureHomogenousCovariance(int index, Eigen::Matrix2f cov);
    void checkSigma();
    void fixSigma();
    Eigen::SparseMatrix<float> getMetric2PixelMap(Eigen::Matrix3f& K);
    Eigen::SparseMatrix<float> getPixel2MetricMap(Eigen::Matrix3f& K);

    /* not in the reference: results of the last checkSigma() (the reference only logs them) and the
     * device status word (bit0 zero pivot, bit1 non-finite state, bit2 capacity) */
    int last_check_negative_diagonals = 0;
    double last_check_max_asymmetry = 0.0;
    int deviceStatus();

private:
    ekfvio_batch* dev_ = nullptr;
    int capacity_ = 0;
    /* what the device state looked like in float the last time the members were refreshed */
    std::vector<float> snap_mu_, snap_feat_, snap_sigma_;
    std::vector<double> dmu_, dfeat_, dP_;
    void ensureCapacity(int n_features);
    void pushIfEdited();
    void pull();
    void recreate(int capacity);
};

#endif
