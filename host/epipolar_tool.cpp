#include "epipolar_tool.hpp"

#include "erp_host_context.hpp"

epipolar_tool::epipolar_tool(std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key
                            , int im_width, int im_height, int output_width, int output_height, int test_key_num)
    : match_size((int)left_key.size()), im_width_(im_width), im_height_(im_height),
      epipole_mat_width(output_width), epipole_mat_height(output_height), n_key(test_key_num)
{
    if (n_key > 7) throw cv::Exception("epipolar_tool: at most 7 test keys (the colour set has 7 entries)");
    if (n_key > match_size || right_key.size() < left_key.size()) throw cv::Exception("epipolar_tool: not enough correspondences");
    // iota + random_shuffle, first n_key entries (src/epipolar_tool.cpp:13-16)
    random_idx.resize(n_key);
    if (n_key > 0) erp_host::check(erp_libstdcxx_sample_table(match_size, 1, n_key, 1, random_idx.data()), "epipolar_tool");
    for (int k = 0; k < n_key; k++) {
        left_key_.push_back(left_key[random_idx[k]]);
        right_key_.push_back(right_key[random_idx[k]]);
    }
}

cv::Mat epipolar_tool::draw_epipole(cv::Mat& test_E_mat)
{
    if (test_E_mat.rows != 3 || test_E_mat.cols != 3 || test_E_mat.type() != CV_64FC1)
        throw cv::Exception("epipolar_tool::draw_epipole: E must be 3x3 CV_64F");
    double e[9];
    for (int i = 0; i < 9; i++) e[i] = test_E_mat.at<double>(i / 3, i % 3);
    cv::Mat out = cv::Mat::zeros(epipole_mat_height, epipole_mat_width, CV_8UC3);
    erp_host::check(erp_draw_epipole(erp_host::context(), e, n_key ? &left_key_[0].pt.x : nullptr, n_key ? &right_key_[0].pt.x : nullptr,
                                     sizeof(cv::KeyPoint), n_key, im_width_, im_height_, epipole_mat_width, epipole_mat_height,
                                     out.data, out.step), "epipolar_tool::draw_epipole");
    return out;
}
