// Drives the C++ facade (include/ekf_vio/*.h) the way the reference's own test programs drive the
// reference classes: test/test_ekf.cpp (assertions + update/process calls), test/jacobian_test.cpp
// (numericallyLinearizeProcess after editing base_mu) and test/analyzeEKFSimulation.cpp (closed
// loop).  Prints "key: values" lines that tests/test_gpu_facade.py compares with the FP64 oracle.
// Usage: facade_test <scenario.bin> [klt.bin]
#include <cstdio>
#include <array>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/ekf_vio/KLTTracker.h"
#include "../../include/ekf_vio/TightlyCoupledEKF.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

static void dump(const char* key, TightlyCoupledEKF& f) {
    std::printf("%s_mu:", key);
    for (int i = 0; i < 22; ++i) std::printf(" %.9g", f.base_mu(i));
    std::printf("\n%s_feat:", key);
    for (auto& e : f.features) std::printf(" %.9g %.9g %.9g", e.getMu()(0), e.getMu()(1), e.getMu()(2));
    std::printf("\n%s_sigma_diag:", key);
    for (int i = 0; i < f.Sigma.rows(); ++i) std::printf(" %.9g", f.Sigma(i, i));
    double s = 0; for (int i = 0; i < f.Sigma.rows(); ++i) for (int j = 0; j < f.Sigma.cols(); ++j) s += (double)f.Sigma(i, j) * (1 + ((i * 31 + j * 17) % 7));
    std::printf("\n%s_sigma_checksum: %.12g\n", key, s);
}

int main(int argc, char** argv) {
    DEFAULT_POINT_DEPTH = D_DEFAULT_POINT_DEPTH; DEFAULT_POINT_DEPTH_VARIANCE = D_DEFAULT_POINT_DEPTH_VARIANCE;
    DEFAULT_POINT_HOMOGENOUS_VARIANCE = D_DEFAULT_POINT_HOMOGENOUS_VARIANCE;

    // ---- test/test_ekf.cpp:27-37: conservativeResize keeps the top-left block
    Eigen::MatrixXf A(2, 2); A(0, 0) = 1; A(1, 1) = 1; A(0, 1) = 2; A(1, 0) = 3;
    Eigen::MatrixXf B = A; B.conservativeResize(4, 4);
    CHECK(B(0, 0) == 1 && B(0, 1) == 2 && B(1, 0) == 3 && B(1, 1) == 1 && B(3, 3) == 0);

    // ---- test/test_ekf.cpp:41-63: H for measured {T,F,T}
    TightlyCoupledEKF tc_ekf;
    std::vector<Eigen::Vector2f> features;
    features.push_back(Eigen::Vector2f(0.1f, 0.1f)); features.push_back(Eigen::Vector2f(-0.1f, -0.1f)); features.push_back(Eigen::Vector2f(0.1f, -0.1f));
    tc_ekf.addNewFeatures(features);
    std::vector<bool> measured = {true, false, true};
    Eigen::SparseMatrix<float> H = tc_ekf.formFeatureMeasurementMap(measured);
    CHECK(H.rows() == 4 && H.cols() == 31);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 31; ++j) {
        bool one = (i == 0 && j == 22) || (i == 1 && j == 23) || (i == 2 && j == 28) || (i == 3 && j == 29);
        CHECK(H(i, j) == (one ? 1.0f : 0.0f));
    }
    // ---- test/test_ekf.cpp:66-82: update with cov 1e-3 I
    std::vector<Eigen::Matrix2f> covs;
    Eigen::Matrix2f cov; cov(0, 0) = 0.001f; cov(1, 1) = 0.001f;
    covs.assign(3, cov);
    tc_ekf.updateWithFeaturePositions(features, covs, measured);
    dump("update3", tc_ekf);
    CHECK(!tc_ekf.features.front().flaggedForDeletion());
    CHECK((++tc_ekf.features.begin())->flaggedForDeletion());

    // ---- test/test_ekf.cpp:97-111: value semantics + 103 features, all measured
    tc_ekf = TightlyCoupledEKF();
    CHECK(tc_ekf.features.size() == 0 && tc_ekf.Sigma.rows() == 22);
    for (int i = 0; i < 100; i++) { features.push_back(Eigen::Vector2f(0.1f, 0.1f)); covs.push_back(cov); measured.push_back(true); }
    tc_ekf.addNewFeatures(features);
    tc_ekf.updateWithFeaturePositions(features, covs, measured);
    dump("update103", tc_ekf);

    // ---- test/test_ekf.cpp:156-204: process-model prints with hand-edited state
    tc_ekf = TightlyCoupledEKF();
    features.resize(3);
    tc_ekf.addNewFeatures(features);
    tc_ekf.base_mu(9) = 1;
    tc_ekf.base_mu(10) = 3.14f;
    {
        auto b = tc_ekf.convolveBaseState(tc_ekf.base_mu, 0.1f);
        auto f = tc_ekf.convolveFeature(tc_ekf.base_mu, tc_ekf.features.front().getMu(), 0.1f);
        std::printf("convolve_base:"); for (int i = 0; i < 22; ++i) std::printf(" %.9g", b(i));
        std::printf("\nconvolve_feat: %.9g %.9g %.9g\n", f(0), f(1), f(2));
    }
    // ---- test/jacobian_test.cpp:34-47: F after editing base_mu through the public member
    tc_ekf.base_mu(10) = 3.1415f; tc_ekf.base_mu(9) = 0; tc_ekf.base_mu(7) = 1;
    {
        Eigen::SparseMatrix<float> F = tc_ekf.numericallyLinearizeProcess(tc_ekf.base_mu, tc_ekf.features, 0.1f);
        CHECK(F.rows() == 31);
        std::printf("jacobian:");
        for (int i = 0; i < 31; ++i) for (int j = 0; j < 31; ++j) std::printf(" %.9g", F(i, j));
        std::printf("\n");
    }
    {
        Eigen::SparseMatrix<float> Q = tc_ekf.generateProcessNoise(0.1f);
        std::printf("noise_diag:"); for (int i = 0; i < Q.rows(); ++i) std::printf(" %.9g", Q(i, i)); std::printf("\n");
    }

    // ---- test/analyzeEKFSimulation.cpp:31-99 closed loop on a scenario written by the Python test
    if (argc > 1) {
        FILE* fp = std::fopen(argv[1], "rb");
        CHECK(fp);
        int n = 0, steps = 0; float dt = 0;
        CHECK(std::fread(&n, 4, 1, fp) == 1 && std::fread(&steps, 4, 1, fp) == 1 && std::fread(&dt, 4, 1, fp) == 1);
        std::vector<float> uv((size_t)n * 2), meas((size_t)steps * n * 2);
        CHECK(std::fread(uv.data(), 4, uv.size(), fp) == uv.size() && std::fread(meas.data(), 4, meas.size(), fp) == meas.size());
        std::fclose(fp);
        TightlyCoupledEKF sim;
        std::vector<Eigen::Vector2f> init;
        for (int i = 0; i < n; ++i) init.push_back(Eigen::Vector2f(uv[2 * i], uv[2 * i + 1]));
        sim.addNewFeatures(init);
        Eigen::Matrix2f c2; c2(0, 0) = 0.00001f; c2(1, 1) = 0.00001f;
        for (int s = 0; s < steps; ++s) {
            sim.process(dt);
            sim.checkSigma();
            CHECK(sim.last_check_negative_diagonals == 0 && sim.last_check_max_asymmetry <= 0.001);
            std::vector<Eigen::Vector2f> z; std::vector<Eigen::Matrix2f> cv2; std::vector<bool> ms;
            for (int i = 0; i < n; ++i) { z.push_back(Eigen::Vector2f(meas[((size_t)s * n + i) * 2], meas[((size_t)s * n + i) * 2 + 1])); cv2.push_back(c2); ms.push_back(true); }
            sim.updateWithFeaturePositions(z, cv2, ms);
            sim.checkSigma();
            CHECK(sim.last_check_negative_diagonals == 0 && sim.last_check_max_asymmetry <= 0.001);
        }
        CHECK(sim.deviceStatus() == 0);
        dump("sim", sim);
        std::printf("sim_depth_var0: %.9g\n", sim.getFeatureDepthVariance(0));
    }

    // ---- KLTTracker::findNewFeaturePositions on a frame pair written by the Python test
    if (argc > 2) {
        FILE* fp = std::fopen(argv[2], "rb");
        CHECK(fp);
        int w = 0, h = 0, n = 0; float fx = 0, fy = 0;
        CHECK(std::fread(&w, 4, 1, fp) == 1 && std::fread(&h, 4, 1, fp) == 1 && std::fread(&n, 4, 1, fp) == 1 && std::fread(&fx, 4, 1, fp) == 1 && std::fread(&fy, 4, 1, fp) == 1);
        std::vector<uint8_t> i0((size_t)w * h), i1((size_t)w * h);
        std::vector<float> prev_metric((size_t)n * 2);
        CHECK(std::fread(i0.data(), 1, i0.size(), fp) == i0.size() && std::fread(i1.data(), 1, i1.size(), fp) == i1.size());
        CHECK(std::fread(prev_metric.data(), 4, prev_metric.size(), fp) == prev_metric.size());
        std::fclose(fp);
        const double k[9] = {fx, 0, w / 2.0, 0, fy, h / 2.0, 0, 0, 1};
        Frame lf(1, cv::Mat(h, w, i0.data(), (size_t)w), k, ros::Time(0.0));
        Frame cf(1, cv::Mat(h, w, i1.data(), (size_t)w), k, ros::Time(0.05));
        std::vector<Eigen::Vector2f> prev;
        std::list<Feature> est;
        for (int i = 0; i < n; ++i) { Eigen::Vector2f m(prev_metric[2 * i], prev_metric[2 * i + 1]); prev.push_back(m); est.push_back(Feature(m, 0.5f)); }
        KLTTracker tracker;
        std::vector<Eigen::Vector2f> measured_positions; std::vector<Eigen::Matrix2f> unc; std::vector<bool> passed;
        tracker.findNewFeaturePositions(lf, cf, prev, est, measured_positions, unc, passed);
        CHECK((int)passed.size() == n);
        std::printf("klt_passed:"); for (int i = 0; i < n; ++i) std::printf(" %d", passed[i] ? 1 : 0);
        std::printf("\nklt_status:"); for (int i = 0; i < n; ++i) std::printf(" %d", (int)tracker.last_status[i]);
        std::printf("\nklt_px:"); for (int i = 0; i < n; ++i) std::printf(" %.9g %.9g", tracker.last_new_fts[i].x, tracker.last_new_fts[i].y);
        std::printf("\nklt_metric:"); for (int i = 0; i < n; ++i) std::printf(" %.9g %.9g", passed[i] ? measured_positions[i].x() : 0.f, passed[i] ? measured_positions[i].y() : 0.f);
        std::printf("\nklt_cov00:"); for (int i = 0; i < n; ++i) std::printf(" %.9g", unc[i](0, 0));
        std::printf("\n");
    }
    {   // accessors (TightlyCoupledEKF.cpp:663-697) and the E2 behaviour of convolveFeature through the class interface
        TightlyCoupledEKF f;
        std::vector<Eigen::Vector2f> fs; fs.push_back(Eigen::Vector2f(0.1f, 0.1f)); fs.push_back(Eigen::Vector2f(-0.2f, 0.05f));
        f.addNewFeatures(fs);
        f.base_mu(7) = 0.2f; f.base_mu(11) = 0.3f;
        f.process(0.05f);
        std::vector<Eigen::Matrix2f> cs; Eigen::Matrix2f c; c(0, 0) = 1e-5f; c(1, 1) = 1e-5f; cs.assign(2, c);
        f.updateWithFeaturePositions(fs, cs, std::vector<bool>{true, true});
        // getFeatureHomogenousCovariance = Sigma.block(start, start, 2, 2), getFeatureDepthVariance = Sigma(start + 2, start + 2)
        for (int i = 0; i < 2; ++i) {
            const int st = 22 + 3 * i;
            Eigen::Matrix2f hc = f.getFeatureHomogenousCovariance(i);
            CHECK(hc(0, 0) == f.Sigma(st, st) && hc(0, 1) == f.Sigma(st, st + 1) && hc(1, 0) == f.Sigma(st + 1, st) && hc(1, 1) == f.Sigma(st + 1, st + 1));
            CHECK(f.getFeatureDepthVariance(i) == f.Sigma(st + 2, st + 2));
            CHECK(hc(0, 0) > 0 && hc(0, 0) < 1e-4f && f.getFeatureDepthVariance(i) > 0);
        }
        // setFeatureHomogenousCovariance writes the four coefficients (and, like the reference, complains on stderr) and the
        // change reaches the device: the next update sees it
        Eigen::Matrix2f nc; nc(0, 0) = 3e-4f; nc(0, 1) = 1e-5f; nc(1, 0) = 1e-5f; nc(1, 1) = 4e-4f;
        f.setFeatureHomogenousCovariance(1, nc);
        Eigen::Matrix2f rb = f.getFeatureHomogenousCovariance(1);
        CHECK(rb(0, 0) == 3e-4f && rb(0, 1) == 1e-5f && rb(1, 0) == 1e-5f && rb(1, 1) == 4e-4f);
        f.process(0.05f);
        CHECK(f.getFeatureHomogenousCovariance(1)(0, 0) > 3e-4f);          // F Sigma F' + Q on top of the value that was set
        // getMetric2PixelMap / getPixel2MetricMap: diag(fx, fy) and its inverse (:680-697)
        Eigen::Matrix3f K; K.setZero(); K(0, 0) = 400.f; K(1, 1) = 410.f; K(0, 2) = 320.f; K(1, 2) = 240.f; K(2, 2) = 1.f;
        Eigen::SparseMatrix<float> m2p = f.getMetric2PixelMap(K), p2m = f.getPixel2MetricMap(K);
        CHECK(m2p.rows() == 2 && m2p.cols() == 2 && m2p(0, 0) == 400.f && m2p(1, 1) == 410.f && m2p(0, 1) == 0.f && m2p(1, 0) == 0.f);
        CHECK(p2m(0, 0) == 1.0f / 400.f && p2m(1, 1) == 1.0f / 410.f && p2m(0, 1) == 0.f);
        // E2 through the class: two convolveFeature calls with the same omega and different dt — the second reuses the rotation
        // cached for the first dt (function-static cache keyed on omega only, TightlyCoupledEKF.cpp:400-446); a filter that
        // never saw the first call rotates by the second dt
        TightlyCoupledEKF a, b;
        Eigen::Matrix<float, BASE_STATE_SIZE, 1> mu; mu.setZero(); mu(3) = 1.f; mu(10) = 0.5f; mu(12) = -0.2f;
        Eigen::Vector3f x(0.3f, -0.1f, 2.0f);
        Eigen::Vector3f a1 = a.convolveFeature(mu, x, 0.1f), a2 = a.convolveFeature(mu, x, 0.3f), b2 = b.convolveFeature(mu, x, 0.3f);
        CHECK(a2(0) == a1(0) && a2(1) == a1(1) && a2(2) == a1(2));            // no velocity: only the (stale) rotation acts
        CHECK(std::fabs(b2(0) - a2(0)) + std::fabs(b2(1) - a2(1)) + std::fabs(b2(2) - a2(2)) > 1e-2f);   // the fresh rotation by 0.3 s differs
        std::printf("convolve_stale: %.9g %.9g %.9g fresh: %.9g %.9g %.9g\n", a2(0), a2(1), a2(2), b2(0), b2(1), b2(2));
        // numericallyLinearizeProcess with a foreign state leaves the filter's own mean alone
        Eigen::Matrix<float, BASE_STATE_SIZE, 1> before = f.base_mu, other = f.base_mu; other(7) = 5.f;
        Eigen::SparseMatrix<float> Fo = f.numericallyLinearizeProcess(other, f.features, 0.05f);
        CHECK(Fo.rows() == 28);
        for (int k2 = 0; k2 < BASE_STATE_SIZE; ++k2) CHECK(f.base_mu(k2) == before(k2));
    }
    {   // Frame with the reference's array type for K (Frame.h:39: boost::array<double, 9>; EKFVIO.cpp:126)
        const int w = 32, h = 16;
        std::vector<uint8_t> im((size_t)w * h, 7);
        std::array<double, 9> ka = {300, 0, 16, 0, 310, 8, 0, 0, 1};
        Frame fa(2, cv::Mat(h, w, im.data(), (size_t)w), ka, std::vector<double>{0, 0, 0, 0, 0}, ros::Time(2.0));
        CHECK(fa.img.cols == 16 && fa.img.rows == 8 && fa.K(0, 0) == 150.f && fa.K(1, 2) == 4.f);
    }
    {   // Frame::Frame with resizing (Frame.cpp:15-41): 2x takes OpenCV's area-fast path (a+b+c+d+2)>>2, 4x the 11-bit bilinear
        const int w = 64, h = 48;
        std::vector<uint8_t> im((size_t)w * h);
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) im[(size_t)y * w + x] = (uint8_t)((x * 7 + y * 13 + (x * y) % 11) & 255);
        const double k[9] = {400, 0, 32, 0, 410, 24, 0, 0, 1};
        Frame f2(2, cv::Mat(h, w, im.data(), (size_t)w), k, std::vector<double>{0.1, 0.2}, ros::Time(1.0));
        CHECK(f2.img.cols == 32 && f2.img.rows == 24);
        CHECK(f2.K(0, 0) == 200.f && f2.K(0, 2) == 16.f && f2.K(1, 1) == 205.f && f2.K(1, 2) == 12.f && f2.K(2, 2) == 1.f);
        CHECK(f2.D(0, 0) == 0.1f && f2.D(0, 1) == 0.2f && f2.D(0, 4) == 0.f);
        for (int y = 0; y < 24; ++y)
            for (int x = 0; x < 32; ++x) {
                const int a = im[(size_t)(2 * y) * w + 2 * x], b = im[(size_t)(2 * y) * w + 2 * x + 1], c = im[(size_t)(2 * y + 1) * w + 2 * x],
                          d = im[(size_t)(2 * y + 1) * w + 2 * x + 1];
                CHECK(f2.img.data[(size_t)y * f2.img.step + x] == (uint8_t)((a + b + c + d + 2) >> 2));
            }
        Frame f4(4, cv::Mat(h, w, im.data(), (size_t)w), k, std::vector<double>(), ros::Time(1.0));
        CHECK(f4.img.cols == 16 && f4.img.rows == 12 && f4.K(0, 0) == 100.f);
        std::printf("frame_resize4:"); for (int i = 0; i < 16 * 12; ++i) std::printf(" %d", (int)f4.img.data[i]);
        std::printf("\n");
    }
    std::printf("facade_test: OK\n");
    return 0;
}
