// epipolar_tool.hpp -- drop-in for the reference's src/epipolar_tool.hpp:7-28: same constructor, same draw_epipole.
//
// The constructor keeps test_key_num (<= 7) correspondences, chosen with the libstdc++ / glibc shuffle the reference's
// constructor performs (src/epipolar_tool.cpp:13-16) replayed from a private generator; draw_epipole renders the
// per-pixel epipolar residual |l^T E^T p| < 0.002 of those correspondences on the B200 (erp_draw_epipole,
// src/epipolar_tool.cpp:84-128) -- no per-pixel bearing tables are kept on the host.
#pragma once

#include "debug_print.h"
#include <opencv2/opencv.hpp>
#include <vector>

class epipolar_tool
{
public:
    epipolar_tool(std::vector<cv::KeyPoint>& left_key,
                  std::vector<cv::KeyPoint>& right_key,
                  int im_width,
                  int im_height,
                  int output_width,
                  int output_height,
                  int test_key_num);

    // output_height x output_width CV_8UC3: epipolar curves of the kept correspondences under test_E_mat, 11 x 11 dots
    // at the right-view keypoints
    cv::Mat draw_epipole(cv::Mat& test_E_mat);

private:
    int n_matches_;                           // correspondences offered to the constructor
    int n_kept_;                              // test_key_num
    std::vector<int> picked_;                 // indices of the kept ones
    std::vector<cv::KeyPoint> kept_left_;
    std::vector<cv::KeyPoint> kept_right_;
    int src_width_, src_height_;              // ERP size the keypoints live in
    int out_width_, out_height_;              // size of the rendered image
};
