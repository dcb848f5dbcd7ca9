/* Frame.h — image + intrinsics, as the tracker and Feature see it (reference
 * include/ekf_vio/Frame.h:24-43, Frame.cpp:15-55).  Two constructors: the reference's resizing one
 * (Frame.cpp:15-41: cv::resize by 1/inv_scale on the device, bit-identical to OpenCV's INTER_LINEAR, K divided by
 * inv_scale), and one that takes an already scaled 8-bit image. */
#ifndef EKFVIO_FRAME_H_
#define EKFVIO_FRAME_H_

#include <utility>
#include <vector>

#include "compat.h"
#include "Params.h"

class Frame {
public:
    Eigen::Matrix<float, 3, 3> K;
    Eigen::Matrix<float, 1, 5> D;   /* distortion coefficients; unused by the hot path */
    cv::Mat img;
    ros::Time t;

    Frame() {}
    /* K9 row-major fx 0 cx 0 fy cy 0 0 1, as sensor_msgs/CameraInfo::K (Frame.cpp:26-33) */
    Frame(int inv_scale, cv::Mat scaled_img, const double k[9], ros::Time _t) : img(scaled_img), t(_t) {
        K.setZero();
        K(0, 0) = (float)(k[0] / inv_scale);
        K(0, 2) = (float)(k[2] / inv_scale);
        K(1, 1) = (float)(k[4] / inv_scale);
        K(1, 2) = (float)(k[5] / inv_scale);
        K(2, 2) = 1.0f;
    }
    /* Frame.cpp:15-41: Frame(inv_scale, full-resolution image, CameraInfo::K, CameraInfo::D, stamp); defined in the facade
     * library (ekfvio_frame_resize_h).  d may hold fewer than five coefficients (missing ones stay 0). */
    Frame(int inv_scale, const cv::Mat& full_img, const double k[9], const std::vector<double>& d, ros::Time _t);
    /* The reference's signature takes sensor_msgs::CameraInfo::K, a boost::array<double, 9> (Frame.h:39; EKFVIO.cpp:126 passes
     * cam->K): any array type with operator[] and contiguous storage — boost::array, std::array — goes through here. */
    template <class Array9, class = decltype(std::declval<const Array9&>().size())>
    Frame(int inv_scale, const cv::Mat& full_img, const Array9& k, const std::vector<double>& d, ros::Time _t) : Frame(inv_scale, full_img, &k[0], d, _t) {}
    /* Frame.cpp:44-55 */
    bool isPixelInBox(cv::Point2f px) const {
        return !(px.x < KILL_PAD || px.y < KILL_PAD || this->img.cols - px.x < KILL_PAD || this->img.rows - px.y < KILL_PAD);
    }
};

#endif
