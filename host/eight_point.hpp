// eight_point.hpp -- drop-in for the reference's src/eight_point.hpp:8-59.
// Same class names and public member signatures; the work runs on the B200 through the C ABI:
//   find                   -> erp_find                    (src/eight_point.cpp:152-192)
//   eight_point_estimation -> erp_eight_point_estimation  (src/eight_point.cpp:16-85)
//   initial_guess          -> erp_initial_guess           (src/eight_point.cpp:87-150)
#pragma once

#include "debug_print.h"
#include "erp_rotation.hpp"
#include <opencv2/opencv.hpp>
#include <vector>

class eight_point
{
public:
    void find(int im_width, int im_height
                       , std::vector<cv::KeyPoint>& key_left, std::vector<cv::KeyPoint>& key_right
                       , cv::Vec3f& R_vec_out, cv::Vec3f& T_vec_out
                       , int match_size);
    void eight_point_estimation(int im_width, int im_height
                            , std::vector<cv::Point3d>& key_point_left_rect, std::vector<cv::Point3d>& key_point_right_rect
                            , cv::Vec3f& R1_vec, cv::Vec3f& R2_vec, cv::Vec3f& T_vec
                            , bool& R1_valid, bool& R2_valid
                            , int match_size);
    void initial_guess(int im_width, int im_height
                    , std::vector<cv::Point3d>& key_point_left_rect, std::vector<cv::Point3d>& key_point_right_rect
                    , cv::Vec3f& R_vec_out, cv::Vec3f& T_vec_out
                    , int match_size);

    // extension: minimal-sample RANSAC on the same bearings (north_star (b)); returns the inlier count
    int ransac(std::vector<cv::Point3d>& key_point_left_rect, std::vector<cv::Point3d>& key_point_right_rect,
               int match_size, int hypotheses, unsigned long long seed, double E_out[9],
               cv::Vec3f& R1_vec, cv::Vec3f& R2_vec, cv::Vec3f& T_vec, std::vector<unsigned char>* inlier_mask = nullptr);

    // ... and on the matched keypoints, with find's argument list: pixels -> bearings -> RANSAC in one device round trip
    int ransac(int im_width, int im_height
               , std::vector<cv::KeyPoint>& key_point_left, std::vector<cv::KeyPoint>& key_point_right
               , int match_size, int hypotheses, unsigned long long seed, double E_out[9]
               , cv::Vec3f& R1_vec, cv::Vec3f& R2_vec, cv::Vec3f& T_vec, std::vector<unsigned char>* inlier_mask = nullptr);

private:
    erp_rotation erp_rot;
    double max_vec(cv::Vec3f& vec);
};

// Random permutation of 0..size-1 read cyclically (src/eight_point.hpp:30-59).  The reference
// shuffles with std::random_shuffle over the never-seeded process-wide rand(); this class replays
// the same libstdc++/glibc sequence from a private generator (erp_libstdcxx_sample_table), so it is
// deterministic and leaves rand() alone.
class random_array
{
public:
    random_array(int size);
    int get_rand()
    {
        int retval = rand_arr[count_];
        count_ = (count_ + 1) % size_;
        return retval;
    }

private:
    int size_;
    std::vector<int> rand_arr;
    int count_;
};
