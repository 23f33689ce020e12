// STAND-IN (builds outside the reference tree only) for the data classes the matchers read and write: ORB_SLAM2::Frame, MapPoint,
// MapLine, KeyFrame -- exactly the members SURVEY.md section 8(b) lists, same names, same types -- so that the very same shim
// sources can be compiled and tested in this repository (no OpenCV C++, Eigen, DBoW2 or g2o in the image): tests/shim/*.cpp.
// Inside the reference tree these names resolve to the reference's own include/Frame.h etc. (shim/olf_ref_classes.h).
#pragma once
#include <array>
#include <cmath>
#include <map>
#include <set>
#include <utility>
#include <vector>
#include "../cv_min.h"
#include "../ORBextractor.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace DBoW2 {             // Thirdparty/DBoW2/DBoW2/FeatureVector.h: std::map<NodeId, std::vector<unsigned int>>
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
}
namespace ORB_SLAM2 {
class KeyFrame; class Frame;
class MapPoint {              // include/MapPoint.h: the members the matchers touch
public:
    // LocalMapping / LoopClosing side: geometry of the point and the bookkeeping Fuse performs (src/MapPoint.cc:385-429 restated)
    cv::Mat GetNormal() { return normal.clone(); }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);
    bool IsInKeyFrame(KeyFrame* pKF) { return obs_kf == pKF; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return obs_kf == pKF ? obs_idx : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx) { added_kf = pKF; added_idx = (int)idx; }
    void Replace(MapPoint* pMP) { if (added_idx < 0) added_idx = pMP->added_idx; else if (pMP->added_idx < 0) pMP->added_idx = added_idx; }      // test hook: where the pair met
    float mfMinDistance = 0, mfMaxDistance = 0; cv::Mat normal; KeyFrame* obs_kf = nullptr; int obs_idx = -1;
    KeyFrame* added_kf = nullptr; int added_idx = -1;
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
    int mnScaleLevels = 0; float mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    double inv_width = 0, inv_height = 0;
};
class KeyFrame {              // include/KeyFrame.h: what the ORBmatcher overloads read and write
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p) s.insert(p); return s; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }      // src/KeyFrame.cc:788-791
    DBoW2::FeatureVector mFeatVec;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    int N = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;      // const int in the reference (include/KeyFrame.h:228-231)
    int mnScaleLevels = 0; float mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    cv::Mat Rcw, tcw, Ow;
};
// MapPoint::PredictScale (src/MapPoint.cc:397-429): log(float) resolves to the float overload in the reference's translation unit
template <class Host> inline int olf_predict_scale(float max_distance, float currentDist, Host* h) {
    const float ratio = max_distance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / h->mfLogScaleFactor);
    if (nScale < 0) nScale = 0; else if (nScale >= h->mnScaleLevels) nScale = h->mnScaleLevels - 1;
    return nScale;
}
inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) { return olf_predict_scale(mfMaxDistance, currentDist, pKF); }
inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) { return olf_predict_scale(mfMaxDistance, currentDist, pF); }
inline void olf_set_le(Frame& F, size_t i, double a, double b, double c) { F.mvle_l[i] = {a, b, c}; }
}  // namespace ORB_SLAM2
