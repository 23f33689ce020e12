// The data classes the hot-path matchers read and write: ORB_SLAM2::Frame, MapPoint, MapLine, KeyFrame.
// Inside the reference tree (-DOLF_IN_REFERENCE_TREE) these ARE the reference's classes (include/Frame.h, MapPoint.h,
// MapLine.h, KeyFrame.h) and the shim sources compile against them unchanged.  In this repository (no OpenCV C++, Eigen,
// DBoW2 or g2o in the image) a stand-in with exactly the members SURVEY.md section 8(b) lists -- same names, same types --
// lets the very same shim sources be compiled and tested (tests/shim/test_shim.cpp).
#pragma once
#ifdef OLF_IN_REFERENCE_TREE
#include "Frame.h"
#include "MapPoint.h"
#include "MapLine.h"
#include "KeyFrame.h"
namespace ORB_SLAM2 { inline void olf_set_le(Frame& F, size_t i, double a, double b, double c) { F.mvle_l[i] = Vector3d(a, b, c); } }
#else
#include <array>
#include <map>
#include <utility>
#include <vector>
#include "cv_min.h"
#include "ORBextractor.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace DBoW2 {             // Thirdparty/DBoW2/DBoW2/FeatureVector.h: std::map<NodeId, std::vector<unsigned int>>
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
}
namespace ORB_SLAM2 {
class MapPoint {              // include/MapPoint.h: the members the hot matchers touch
public:
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0; bool mbTrackInView = false; int mnTrackScaleLevel = 0; float mTrackViewCos = 0;
    long unsigned int mnLastFrameSeen = 0;
    bool isBad() { return bad; }
    int Observations() { return nobs; }
    cv::Mat GetDescriptor() { return desc.clone(); }          // 1 x 32 CV_8U
    cv::Mat GetWorldPos() { return pos.clone(); }             // 3 x 1 CV_32F
    bool bad = false; int nobs = 1; cv::Mat desc, pos;
};
class MapLine {               // include/MapLine.h
public:
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat desc;
};
class Frame {                 // include/Frame.h:83-259
public:
    void ComputeStereoMatches();
    void ComputeStereoMatches_Lines(bool initial = false);
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    int N = 0, N_l = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<cv::line_descriptor::KeyLine> mvKeys_Line, mvKeysRight_Line;
    std::vector<float> mvuRight, mvDepth;
    std::vector<std::pair<float, float>> mvDisparity_l;
    std::vector<std::array<double, 3>> mvle_l;               // Eigen::Vector3d in the reference
    cv::Mat mDescriptors, mDescriptorsRight, mDescriptors_Line, mDescriptorsRight_Line;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mTcw;                                             // 4 x 4 CV_32F
    int mnScaleLevels = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    double inv_width = 0, inv_height = 0;
};
class KeyFrame {              // include/KeyFrame.h: what SearchByBoW(KeyFrame*, Frame&, ...) reads
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    DBoW2::FeatureVector mFeatVec;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
};
inline void olf_set_le(Frame& F, size_t i, double a, double b, double c) { F.mvle_l[i] = {a, b, c}; }
}  // namespace ORB_SLAM2
#endif
