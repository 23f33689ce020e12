// Drop-in for the descriptor-only functions of include/LineMatcher.h:57-69 (namespace ORB_SLAM2) and for the two
// stereo-association members of Frame that sit on the hot path; they take plain data so that src/Frame.cc only needs
// the two call sites shown in INTEGRATION.md.
#pragma once
#include <vector>
#include <utility>
#include "cv_min.h"
#include "../../include/olf_abi.h"

namespace ORB_SLAM2 {
class ORBextractor;
struct OlfConfig {                          // the Config:: values the line matchers read (src/Config.cpp:42-87)
    static olf_line_match_params line_match;
    static int device;
};
int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
int distance(const cv::Mat& a, const cv::Mat& b);
// Frame::ComputeStereoMatches (src/Frame.cc:702-876): fills mvuRight / mvDepth
void ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc,
                          const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight, float bf, float fx,
                          std::vector<float>& uRight, std::vector<float>& depth);
// Frame::ComputeStereoMatches_Lines (src/Frame.cc:878-1000): fills mvDisparity_l / mvle_l (doNotDropMonoLines == true)
void ComputeStereoMatches_Lines(const std::vector<cv::line_descriptor::KeyLine>& keys, const cv::Mat& desc,
                                const std::vector<cv::line_descriptor::KeyLine>& keysRight, const cv::Mat& descRight,
                                int img_width, int img_height, std::vector<std::pair<float, float>>& disparity, std::vector<double>& le /*3 per line*/);
}  // namespace ORB_SLAM2
