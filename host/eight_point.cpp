// eight_point.cpp -- host side of the eight-point boundary.
#include "eight_point.hpp"

#include "erp_host_context.hpp"

using namespace std;
using namespace cv;

static_assert(sizeof(Point3d) == 24, "cv::Point3d must be three packed doubles");

double eight_point::max_vec(Vec3f& vec)      // src/eight_point.cpp:6-14
{
    if ((vec[0] > vec[1]) && (vec[0] > vec[2])) return vec[0];
    else if (vec[1] > vec[2]) return vec[1];
    else return vec[2];
}

void eight_point::find(int im_width, int im_height
                       , vector<KeyPoint>& key_left, vector<KeyPoint>& key_right
                       , Vec3f& R_vec_out, Vec3f& T_vec_out
                       , int match_size)
{
    if (match_size > (int)key_left.size() || match_size > (int)key_right.size())
        CV_Error(cv::Error::StsBadArg, "eight_point::find: match_size exceeds the keypoint vectors");
    // KeyPoint.pt is the first member: stride sizeof(KeyPoint)
    erp_host::Lock lock;
    erp_host::check(erp_find(erp_host::context(), im_width, im_height, key_left.data(), key_right.data(), sizeof(KeyPoint),
                             match_size, nullptr, 0, 0, R_vec_out.val, T_vec_out.val), "eight_point::find");
}

void eight_point::eight_point_estimation(int im_width, int im_height
                            , vector<Point3d>& key_point_left_rect, vector<Point3d>& key_point_right_rect
                            , Vec3f& R1_vec, Vec3f& R2_vec, Vec3f& T_vec
                            , bool& R1_valid, bool& R2_valid
                            , int match_size)
{
    (void)im_width; (void)im_height;     // unused by the reference as well
    int v1 = 0, v2 = 0;
    erp_host::Lock lock;
    erp_host::check(erp_eight_point_estimation(erp_host::context(), &key_point_left_rect[0].x, &key_point_right_rect[0].x, match_size,
                                               nullptr, R1_vec.val, R2_vec.val, T_vec.val, &v1, &v2), "eight_point::eight_point_estimation");
    R1_valid = v1 != 0;
    R2_valid = v2 != 0;
}

void eight_point::initial_guess(int im_width, int im_height
                    , vector<Point3d>& key_point_left_rect, vector<Point3d>& key_point_right_rect
                    , Vec3f& R_vec_out, Vec3f& T_vec_out
                    , int match_size)
{
    (void)im_width; (void)im_height;
    const int H = 80, S = (int)(match_size * 0.25);      // src/eight_point.cpp:99,102
    erp_host::Lock lock;
    erp_host::check(erp_initial_guess(erp_host::context(), &key_point_left_rect[0].x, &key_point_right_rect[0].x, match_size,
                                      nullptr, H, S, R_vec_out.val, T_vec_out.val, nullptr, nullptr, nullptr, nullptr),
                    "eight_point::initial_guess");
}

int eight_point::ransac(vector<Point3d>& l, vector<Point3d>& r, int match_size, int hypotheses, unsigned long long seed,
                        double E_out[9], Vec3f& R1_vec, Vec3f& R2_vec, Vec3f& T_vec, vector<unsigned char>* inlier_mask)
{
    erp_ransac_result res;
    if (inlier_mask) inlier_mask->resize(match_size);
    erp_host::Lock lock;
    erp_host::check(erp_ransac(erp_host::context(), &l[0].x, &r[0].x, match_size, seed, 0, hypotheses, 8, ERP_METRIC_ALGEBRAIC,
                               0.002f, &res, inlier_mask ? inlier_mask->data() : nullptr), "eight_point::ransac");
    for (int i = 0; i < 9; i++) E_out[i] = res.E_refit[i];
    for (int i = 0; i < 3; i++) { R1_vec[i] = res.pose[i]; R2_vec[i] = res.pose[3 + i]; T_vec[i] = res.pose[6 + i]; }
    return res.count;
}

int eight_point::ransac(int im_width, int im_height, vector<KeyPoint>& kl, vector<KeyPoint>& kr, int match_size, int hypotheses,
                        unsigned long long seed, double E_out[9], Vec3f& R1_vec, Vec3f& R2_vec, Vec3f& T_vec,
                        vector<unsigned char>* inlier_mask)
{
    erp_ransac_result res;
    if (inlier_mask) inlier_mask->resize(match_size);
    erp_host::Lock lock;
    erp_host::check(erp_ransac_pixels(erp_host::context(), im_width, im_height, &kl[0].pt, &kr[0].pt, sizeof(KeyPoint), match_size,
                                      seed, 0, hypotheses, 8, ERP_METRIC_ALGEBRAIC, 0.002f, &res,
                                      inlier_mask ? inlier_mask->data() : nullptr), "eight_point::ransac");
    for (int i = 0; i < 9; i++) E_out[i] = res.E_refit[i];
    for (int i = 0; i < 3; i++) { R1_vec[i] = res.pose[i]; R2_vec[i] = res.pose[3 + i]; T_vec[i] = res.pose[6 + i]; }
    return res.count;
}

random_array::random_array(int size) : length_(size), order_(size > 0 ? size : 0), next_(0)
{
    if (size > 0) erp_host::check(erp_libstdcxx_sample_table(size, 1, size, 1, order_.data()), "random_array");
}
