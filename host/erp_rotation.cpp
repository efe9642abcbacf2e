// XYZ-Euler helpers (reference: src/erp_rotation.cpp:14-63), written against plain arrays.
#include "erp_rotation.hpp"

#include "erp_host_context.hpp"

cv::Mat erp_rotation::eular2rot(cv::Vec3d theta)
{
    const double cx = std::cos(theta[0]), sx = std::sin(theta[0]);
    const double cy = std::cos(theta[1]), sy = std::sin(theta[1]);
    const double cz = std::cos(theta[2]), sz = std::sin(theta[2]);
    // Rx*Ry*Rz expanded
    const double r[9] = {cy * cz, -cy * sz, sy,
                         sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy,
                         -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy};
    cv::Mat R(3, 3, CV_64FC1);
    for (int i = 0; i < 9; i++) R.at<double>(i / 3, i % 3) = r[i];
    return R;
}

cv::Vec3d erp_rotation::rot2eular(cv::Mat R)
{
    const double r22 = R.at<double>(2, 2), r12 = R.at<double>(1, 2);
    const double sy = std::sqrt(r22 * r22 + r12 * r12);
    const double x = sy < 1e-6 ? 0.0 : std::atan2(-r12, r22);
    return cv::Vec3d(x, std::atan2(R.at<double>(0, 2), sy), std::atan2(-R.at<double>(0, 1), R.at<double>(0, 0)));
}

static void mat9(const cv::Mat& R, double* m)
{
    if (R.rows != 3 || R.cols != 3 || R.type() != CV_64FC1) CV_Error(cv::Error::StsBadArg, "erp_rotation: rotation matrix must be 3x3 CV_64F");
    for (int i = 0; i < 9; i++) m[i] = R.at<double>(i / 3, i % 3);
}

// One pixel is fifteen lines of trigonometry: it stays on the host (the reference's callers invoke it per pixel from
// OpenMP loops, src/spherical_surf.cpp:27-45,53-62 -- a device round trip each would be absurd).  Same operation order
// as src/erp_rotation.cpp:66-92 and as rotate_pixels_kernel / the oracle, so the integer results agree:
//   (row, col) -> colatitude, longitude -> bearing (OMAF axes: -sin cos, sin sin, cos) -> R * bearing -> back, truncated.
// Whole images go through rotate_image (device).
cv::Vec2i erp_rotation::rotate_pixel(const cv::Vec2i& in_vec, cv::Mat& rot_mat, int width, int height)
{
    double m[9];
    mat9(rot_mat, m);
    const double colat = M_PI * in_vec[0] / height, lon = 2 * M_PI * in_vec[1] / width;
    const double b[3] = {-std::sin(colat) * std::cos(lon), std::sin(colat) * std::sin(lon), std::cos(colat)};
    double r[3];
    for (int i = 0; i < 3; i++) r[i] = m[3 * i] * b[0] + m[3 * i + 1] * b[1] + m[3 * i + 2] * b[2];
    const double colat_r = std::acos(r[2]);
    double lon_r = std::atan2(r[1], -r[0]);
    if (lon_r < 0) lon_r += M_PI * 2;
    cv::Vec2i px;
    px[0] = (int)(height * colat_r / M_PI);
    px[1] = (int)(width * lon_r / (2 * M_PI));
    return px;
}

cv::Mat erp_rotation::rotate_image(const cv::Mat& im, cv::Mat& rot_mat)
{
    if (im.type() != CV_8UC3) CV_Error(cv::Error::StsBadArg, "erp_rotation::rotate_image: CV_8UC3 image expected");
    double m[9];
    mat9(rot_mat, m);
    cv::Mat out(im.rows, im.cols, im.type());
    erp_host::Lock lock;
    erp_host::check(erp_rotate_image(erp_host::context(), im.data, im.cols, im.rows, im.step, m, out.data, out.step),
                    "erp_rotation::rotate_image");
    return out;
}
