// erp_rotation.hpp -- the two pure functions of the reference's erp_rotation class that the hot
// path needs on the host (src/erp_rotation.hpp:13-14, src/erp_rotation.cpp:14-63).  rotate_pixel /
// rotate_image are image warps outside the hot path ("next").
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
};
