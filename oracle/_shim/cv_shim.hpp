// TEST INFRASTRUCTURE ONLY — stand-in for the OpenCV types that Feature.h / Frame.h name (cv::Point2f, cv::Mat, cv::Size), so
// the unmodified reference EKF sources compile here.  No OpenCV arithmetic is provided: the reference's KLT path is OpenCV
// itself and is pinned through the cv2 wheel instead (tests/golden/make_klt_golden.py).
#pragma once
namespace cv {
struct Point2f {
    float x = 0.f, y = 0.f;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Mat {
    int rows = 0, cols = 0;
};
}  // namespace cv
