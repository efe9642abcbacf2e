#include "epipolar_tool.hpp"

#include "erp_host_context.hpp"

epipolar_tool::epipolar_tool(std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int im_width, int im_height,
                             int output_width, int output_height, int test_key_num)
    : n_matches_((int)left_key.size()), n_kept_(test_key_num), src_width_(im_width), src_height_(im_height),
      out_width_(output_width), out_height_(output_height)
{
    if (n_kept_ > 7) CV_Error(cv::Error::StsBadArg, "epipolar_tool: at most 7 test keys (the colour set has 7 entries)");
    if (n_kept_ > n_matches_ || right_key.size() < left_key.size()) CV_Error(cv::Error::StsBadArg, "epipolar_tool: not enough correspondences");
    // iota + random_shuffle, first n_kept_ entries (src/epipolar_tool.cpp:13-16)
    picked_.resize(n_kept_);
    if (n_kept_ > 0) erp_host::check(erp_libstdcxx_sample_table(n_matches_, 1, n_kept_, 1, picked_.data()), "epipolar_tool");
    for (int k = 0; k < n_kept_; k++) {
        kept_left_.push_back(left_key[picked_[k]]);
        kept_right_.push_back(right_key[picked_[k]]);
    }
}

cv::Mat epipolar_tool::draw_epipole(cv::Mat& test_E_mat)
{
    if (test_E_mat.rows != 3 || test_E_mat.cols != 3 || test_E_mat.type() != CV_64FC1)
        CV_Error(cv::Error::StsBadArg, "epipolar_tool::draw_epipole: E must be 3x3 CV_64F");
    double e[9];
    for (int i = 0; i < 9; i++) e[i] = test_E_mat.at<double>(i / 3, i % 3);
    cv::Mat out = cv::Mat::zeros(out_height_, out_width_, CV_8UC3);
    erp_host::Lock lock;
    erp_host::check(erp_draw_epipole(erp_host::context(), e, n_kept_ ? &kept_left_[0].pt.x : nullptr, n_kept_ ? &kept_right_[0].pt.x : nullptr,
                                     sizeof(cv::KeyPoint), n_kept_, src_width_, src_height_, out_width_, out_height_,
                                     out.data, out.step), "epipolar_tool::draw_epipole");
    return out;
}
