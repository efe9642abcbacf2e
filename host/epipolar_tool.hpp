// epipolar_tool.hpp -- drop-in for the reference's src/epipolar_tool.hpp:7-28.  Same constructor and
// draw_epipole signature; the per-pixel residual image is rendered on the B200 (erp_draw_epipole,
// src/epipolar_tool.cpp:84-128).  The 7 test correspondences are picked with the same libstdc++ /
// glibc shuffle the reference's constructor performs (src/epipolar_tool.cpp:13-16), replayed from a
// private generator.
#pragma once

#include "debug_print.h"
#include <opencv2/opencv.hpp>
#include <vector>

class epipolar_tool
{
public:
    epipolar_tool(std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key
                     , int im_width, int im_height, int output_width, int output_height, int test_key_num);
    cv::Mat draw_epipole(cv::Mat& test_E_mat);

private:
    int match_size;
    std::vector<int> random_idx;
    std::vector<cv::KeyPoint> left_key_;      // the selected correspondences
    std::vector<cv::KeyPoint> right_key_;
    int im_width_, im_height_;
    int epipole_mat_width;
    int epipole_mat_height;
    int n_key;
};
