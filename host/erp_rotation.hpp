// erp_rotation.hpp -- drop-in for the reference's src/erp_rotation.hpp:9-19 (same class, same four members).
//
//   eular2rot, rot2eular    host arithmetic, src/erp_rotation.cpp:14-63
//   rotate_pixel            src/erp_rotation.cpp:66-92   host arithmetic as well: the callers invoke it per pixel from
//                           OpenMP loops; thread safe, no CUDA (batches: erp_rotate_pixels / erp_crop_rotated_image)
//   rotate_image            src/erp_rotation.cpp:94-122  -> erp_rotate_image (inverse-mapped nearest-neighbour warp)
#ifndef ERP_B200_HOST_ERP_ROTATION_HPP
#define ERP_B200_HOST_ERP_ROTATION_HPP
#include <cmath>

#include <opencv2/opencv.hpp>          // the real OpenCV, or the type shim under host/compat
#include "debug_print.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
// degrees <-> radians, as the callers spell them (src/erp_rotation.hpp:6-7)
// (token for token the reference's expansions: unparenthesised, callers rely on that)
#define RAD(x)    M_PI * (x) / 180.0
#define DEGREE(x) 180.0 * (x) / M_PI

class erp_rotation {
public:
    // XYZ Euler angles (radians) -> 3x3 CV_64F rotation, R = Rx * Ry * Rz
    cv::Mat
    eular2rot(cv::Vec3d theta);

    // the inverse; x = 0 in the singular case sqrt(R22^2 + R12^2) < 1e-6
    cv::Vec3d
    rot2eular(cv::Mat R);

    // (row, col) of an ERP pixel after rotating its bearing by rot_mat
    cv::Vec2i rotate_pixel(const cv::Vec2i& in_vec,
                           cv::Mat& rot_mat,
                           int width,
                           int height);

    // the whole image: every output pixel fetches its source through rot_mat^-1
    cv::Mat rotate_image(const cv::Mat& im,
                         cv::Mat& rot_mat);
};

#endif  // ERP_B200_HOST_ERP_ROTATION_HPP
