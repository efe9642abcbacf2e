// erp_rotation.hpp -- drop-in for the reference's src/erp_rotation.hpp:9-19.  eular2rot / rot2eular
// are host arithmetic (src/erp_rotation.cpp:14-63); rotate_pixel / rotate_image run on the B200
// through the C ABI (erp_rotate_pixels, erp_rotate_image; src/erp_rotation.cpp:66-122).
#pragma once
#include <cmath>

#include "debug_print.h"
#include <opencv2/opencv.hpp>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define RAD(x) M_PI*(x)/180.0
#define DEGREE(x) 180.0*(x)/M_PI

class erp_rotation
{
public:
    cv::Mat eular2rot(cv::Vec3d theta);      // R = Rx * Ry * Rz, XYZ Euler
    cv::Vec3d rot2eular(cv::Mat R);
    // in_vec = (row, col).  One pixel per call costs a device round trip: batch through erp_rotate_pixels.
    cv::Vec2i rotate_pixel(const cv::Vec2i& in_vec, cv::Mat& rot_mat, int width, int height);
    cv::Mat rotate_image(const cv::Mat& im, cv::Mat& rot_mat);
};
