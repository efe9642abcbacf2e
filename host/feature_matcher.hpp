// feature_matcher.hpp -- drop-in for the reference's src/feature_matcher.hpp:26-51.
// Same class name, same public member signatures; match_two_image (the hot path,
// src/feature_matcher.cpp:42-59) runs on the B200 through the C ABI (erp_knn2_match).
// SURF detection / description / drawing are outside the hot path: they forward to OpenCV when
// the build has the real library (xfeatures2d) and throw otherwise.
#pragma once

#define _USE_MATH_DEFINES
#include "debug_print.h"
#include "opencv2/opencv_modules.hpp"

#include <cmath>
#include <vector>

#include "opencv2/core/ocl.hpp"
#include "opencv2/imgproc.hpp"
#include "opencv2/imgcodecs.hpp"
#include "opencv2/highgui.hpp"
#include "opencv2/calib3d.hpp"
#include "opencv2/features2d.hpp"
#include "opencv2/xfeatures2d.hpp"

class feature_matcher
{
public:
    void init();
    void deinit();
    feature_matcher() { init(); }
    ~feature_matcher() { deinit(); }

    std::vector<cv::KeyPoint> detect_key_point(const cv::Mat &image);
    cv::Mat comput_descriptor(const cv::Mat &image, std::vector<cv::KeyPoint> &key_point);
    std::vector<cv::DMatch> match_two_image(const cv::Mat &descriptor1, const cv::Mat &descriptor2);
    cv::Mat draw_match(const cv::Mat& im_left, const cv::Mat& im_right, const std::vector<cv::KeyPoint>& key_left, const std::vector<cv::KeyPoint>& key_right);

    void do_all(const cv::Mat &im_left, const cv::Mat &im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num);

    // extensions: Lowe ratio (reference: 0.3f) and mutual-nearest cross-check
    float ratio_thresh = 0.3f;
    bool cross_check = false;
    // SURF-128: re-creates the detector / extractor with extended descriptors (SURF::create(100, 4, 3, true));
    // match_two_image itself takes any descriptor width that is a multiple of 4 (64 and 128 run on the tensor cores)
    void set_extended(bool extended_descriptors);
    bool extended() const { return extended_; }

private:
#ifndef ERP_OPENCV_COMPAT
    cv::Ptr<cv::Feature2D> detector;
    cv::Ptr<cv::Feature2D> descriptor_extractor;
#endif
    bool extended_ = false;
    std::vector<cv::DMatch> matches;     // last result (draw_match reads it, as in the reference)
};
