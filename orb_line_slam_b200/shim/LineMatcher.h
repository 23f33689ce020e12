// Drop-in for the descriptor-only functions of include/LineMatcher.h:57-69 (namespace ORB_SLAM2) and for the two
// stereo-association members of Frame that sit on the hot path; they take plain data so that src/Frame.cc only needs
// the two call sites shown in INTEGRATION.md.
#pragma once
#include <cmath>
#include <vector>
#include <utility>
#include "cv_min.h"
#ifdef OLF_IN_REFERENCE_TREE
#include "gridStructure.h"          // the reference's own (include/gridStructure.h)
#else
#include "standin/gridStructure.h"
#endif
#include "../../include/olf_abi.h"

namespace ORB_SLAM2 {
class ORBextractor; class Frame; class MapLine;
typedef std::pair<int, int> point_2d;                    // include/LineMatcher.h:44-45
typedef std::pair<point_2d, point_2d> line_2d;
inline double dot(const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first * b.first + a.second * b.second; }
inline void normalize(std::pair<double, double>& v) { const double m = std::sqrt(dot(v, v)); v.first /= m; v.second /= m; }
struct OlfConfig {                          // the Config:: values the line matchers read (src/Config.cpp:42-87)
    static olf_line_match_params line_match;
    static int device;
    static double min_ratio_12_p;
};
int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12);
int distance(const cv::Mat& a, const cv::Mat& b);
// match(mvpLocalMapLines, CurrentFrame, nnr, matches_12) (include/LineMatcher.h:59, src/LineMatcher.cpp:64-73: matchNNR of the map
// lines' descriptors against CurrentFrame.mDescriptors_Line)
int match(const std::vector<MapLine*>& mvpLocalMapLines, Frame& CurrentFrame, float nnr, std::vector<int>& matches_12);
// matchGrid, lines (include/LineMatcher.h:69, src/LineMatcher.cpp:220-299) on the device; points (:152-218) on the host (not on the hot path)
int matchGrid(const std::vector<line_2d>& lines1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2,
              const std::vector<std::pair<double, double>>& directions2, const GridWindow& w, std::vector<int>& matches_12);
int matchGrid(const std::vector<point_2d>& points1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2, const GridWindow& w, std::vector<int>& matches_12);
// Frame::ComputeStereoMatches (src/Frame.cc:702-876): fills mvuRight / mvDepth
void ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc,
                          const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight, float bf, float fx,
                          std::vector<float>& uRight, std::vector<float>& depth);
// Frame::ComputeStereoMatches_Lines (src/Frame.cc:878-1000): fills mvDisparity_l / mvle_l (doNotDropMonoLines == true)
void ComputeStereoMatches_Lines(const std::vector<cv::line_descriptor::KeyLine>& keys, const cv::Mat& desc,
                                const std::vector<cv::line_descriptor::KeyLine>& keysRight, const cv::Mat& descRight,
                                int img_width, int img_height, std::vector<std::pair<float, float>>& disparity, std::vector<double>& le /*3 per line*/);
}  // namespace ORB_SLAM2
