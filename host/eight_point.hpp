// eight_point.hpp -- drop-in for the reference's src/eight_point.hpp:8-59: classes eight_point and random_array with the
// reference's public member signatures.  The arithmetic runs on the B200 through the C ABI (include/erp_b200.h):
//
//   eight_point::find                    src/eight_point.cpp:152-192  -> erp_find
//   eight_point::eight_point_estimation  src/eight_point.cpp:16-85    -> erp_eight_point_estimation
//   eight_point::initial_guess           src/eight_point.cpp:87-150   -> erp_initial_guess
//   random_array                         src/eight_point.hpp:30-59    -> erp_libstdcxx_sample_table
//
// plus the RANSAC entry points north_star asks for (extensions: the reference has no inlier-counting RANSAC).
#pragma once

#include "debug_print.h"
#include "erp_rotation.hpp"
#include <opencv2/opencv.hpp>
#include <algorithm>
#include <numeric>      // the reference header pulls it in (src/eight_point.hpp:6): callers may lean on it
#include <vector>

class eight_point
{
public:
    // Matched keypoints (pixels of a width x height ERP image, first match_size entries) -> relative pose: XYZ Euler
    // angles in radians and a unit translation.  Pixels become bearings, then initial_guess runs.
    void find(int im_width,
              int im_height,
              std::vector<cv::KeyPoint>& key_left,
              std::vector<cv::KeyPoint>& key_right,
              cv::Vec3f& R_vec_out,
              cv::Vec3f& T_vec_out,
              int match_size);

    // N-point least-squares essential matrix on the first match_size bearing pairs, rank-2 projected and decomposed:
    // both rotations as Euler vectors, the translation, and "valid" = every |angle| < 1.57.  The image size is unused,
    // as in the reference.
    void eight_point_estimation(int im_width,
                                int im_height,
                                std::vector<cv::Point3d>& key_point_left_rect,
                                std::vector<cv::Point3d>& key_point_right_rect,
                                cv::Vec3f& R1_vec,
                                cv::Vec3f& R2_vec,
                                cv::Vec3f& T_vec,
                                bool& R1_valid,
                                bool& R2_valid,
                                int match_size);

    // The reference's estimator: 80 rounds on random quarters of the correspondences, every valid (R, t) collected,
    // the candidate with the smallest trimmed-mean distance to the others wins.  The sample sequence is the one a
    // fresh process draws from glibc rand() through libstdc++'s random_shuffle.
    void initial_guess(int im_width,
                       int im_height,
                       std::vector<cv::Point3d>& key_point_left_rect,
                       std::vector<cv::Point3d>& key_point_right_rect,
                       cv::Vec3f& R_vec_out,
                       cv::Vec3f& T_vec_out,
                       int match_size);

    // --- extensions ---------------------------------------------------------------------------------------------
    // Minimal-sample RANSAC (8 points per hypothesis, Philox sampling, residual |l^T E r| < 0.002, refit on the inliers)
    // on bearings; returns the inlier count of the winner, E_out = refitted essential matrix (row major).
    int ransac(std::vector<cv::Point3d>& key_point_left_rect,
               std::vector<cv::Point3d>& key_point_right_rect,
               int match_size,
               int hypotheses,
               unsigned long long seed,
               double E_out[9],
               cv::Vec3f& R1_vec,
               cv::Vec3f& R2_vec,
               cv::Vec3f& T_vec,
               std::vector<unsigned char>* inlier_mask = nullptr);

    // The same on matched keypoints (find's argument list): pixels -> bearings -> RANSAC in one device round trip.
    int ransac(int im_width,
               int im_height,
               std::vector<cv::KeyPoint>& key_point_left,
               std::vector<cv::KeyPoint>& key_point_right,
               int match_size,
               int hypotheses,
               unsigned long long seed,
               double E_out[9],
               cv::Vec3f& R1_vec,
               cv::Vec3f& R2_vec,
               cv::Vec3f& T_vec,
               std::vector<unsigned char>* inlier_mask = nullptr);

private:
    erp_rotation erp_rot;
    double max_vec(cv::Vec3f& vec);           // largest of three (callers pass absolute values)
};

// A random permutation of 0 .. size-1, read cyclically.  The reference shuffles with std::random_shuffle over the
// never-seeded process-wide rand(); this class replays that libstdc++ / glibc sequence from a private generator, so the
// numbers are the reference's on Linux, deterministic, and rand() is left alone.
class random_array
{
public:
    random_array(int size);

    int get_rand()
    {
        const int value = order_[next_];
        next_ = next_ + 1 == length_ ? 0 : next_ + 1;
        return value;
    }

private:
    int length_;
    std::vector<int> order_;
    int next_;
};
