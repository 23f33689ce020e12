// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Stand-ins for the reference classes whose real headers need Eigen / g2o / DBoW2 / Pangolin (include/Frame.h,
// MapPoint.h, MapLine.h, KeyFrame.h): exactly the members that the hot-path functions sliced out of src/Frame.cc and
// src/ORBmatcher.cc read or write, with the reference's names and types (include/Frame.h:83-259, include/MapPoint.h).
// The Makefile pre-defines the include guards of the real headers so that the reference's own ORBmatcher.h /
// LineMatcher.h (which include them) see these instead.
#pragma once
#include "cvstub.hpp"
#include "eigen_stub.hpp"
#include <line_descriptor/descriptor_custom.hpp>
#include "ORBextractor.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include <mutex>
#include <set>
#include <utility>
using namespace std;
using namespace Eigen;

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {
class Map; class KeyFrameDatabase; class Frame; class MapPoint;
// include/KeyFrame.h: the members the ORBmatcher overloads of LocalMapping / LoopClosing / relocalisation touch (same names and types;
// GetFeaturesInArea and IsInImage are the reference's own text, sliced from src/KeyFrame.cc)
class KeyFrame {
public:
    std::vector<float> mvLevelSigma2, mvInvLevelSigma2, mvScaleFactors; cv::Mat mDescriptors; bool bad_ = false; bool isBad() { return bad_; }
    cv::Mat GetRotation() { return Rcw_.clone(); }
    cv::Mat GetTranslation() { return tcw_.clone(); }
    cv::Mat GetCameraCenter() { return Ow_.clone(); }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p) s.insert(p); return s; }
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn; std::vector<float> mvuRight;
    DBoW2::FeatureVector mFeatVec;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;     // const int in the reference
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0, mfLogScaleFactor = 0; int mnScaleLevels = 0;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Rcw_, tcw_, Ow_;
};
class MapPoint {
public:
    // LocalMapping / LoopClosing side (include/MapPoint.h); PredictScale and Get*DistanceInvariance are sliced from src/MapPoint.cc
    cv::Mat GetNormal() { return normal_.clone(); }
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);
    bool IsInKeyFrame(KeyFrame* pKF) { return in_kf_ == pKF; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return in_kf_ == pKF ? kf_idx_ : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx) { fused_ = (int)idx; (void)pKF; }
    void Replace(MapPoint* pMP) { if (fused_ < 0) fused_ = pMP->fused_; else if (pMP->fused_ < 0) pMP->fused_ = fused_; }      // records where the pair met
    std::mutex mMutexPos; float mfMinDistance = 0, mfMaxDistance = 0; cv::Mat normal_; KeyFrame* in_kf_ = nullptr; int kf_idx_ = -1, fused_ = -1;
    // set by Frame::isInFrustum (src/Frame.cc:436-441)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0; bool mbTrackInView = false; int mnTrackScaleLevel = 0; float mTrackViewCos = 0;
    long unsigned int mnLastFrameSeen = 0;
    bool isBad() { return bad_; }
    int Observations() { return nobs_; }
    cv::Mat GetDescriptor() { return desc_.clone(); }
    cv::Mat GetWorldPos() { return pos_.clone(); }
    bool bad_ = false; int nobs_ = 1; cv::Mat desc_, pos_;
    // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:254-322)
    void ComputeDistinctiveDescriptors();
    std::mutex mMutexFeatures; bool mbBad = false; std::map<KeyFrame*, size_t> mObservations; cv::Mat mDescriptor;
};
class MapLine {
public:
    cv::Mat GetDescriptor() { return desc_.clone(); }
    cv::Mat desc_;
};
class Frame {
public:
    void AssignFeaturesToGrid();
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void ComputeStereoMatches();
    void ComputeStereoMatches_Lines(bool initial = false);
    double lineSegmentOverlapStereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj);
    void filterLineSegmentDisparity(Vector2d spl, Vector2d epl, Vector2d spr, Vector2d epr, double& disp_s, double& disp_e);

    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0;
    float mbf = 0, mb = 0, mThDepth = 0;
    int N = 0, N_l = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<cv::line_descriptor::KeyLine> mvKeys_Line, mvKeysRight_Line;
    std::vector<float> mvuRight, mvDepth;
    std::vector<pair<float, float>> mvDisparity_l;
    std::vector<Vector3d> mvle_l;
    cv::Mat mDescriptors, mDescriptorsRight, mDescriptors_Line, mDescriptorsRight_Line;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    int mnScaleLevels = 0; float mfScaleFactor = 0, mfLogScaleFactor = 0;
    vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    double inv_width = 0, inv_height = 0;
};
}  // namespace ORB_SLAM2
