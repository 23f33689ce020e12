// Drop-in for the hot-path part of include/ORBmatcher.h:37-103 (class ORB_SLAM2::ORBmatcher): constructor, DescriptorDistance,
// RadiusByViewingCos, the three SearchByProjection overloads Tracking calls every frame and SearchByBoW(KeyFrame*, Frame&, ...).
// Same signatures; the bodies gather the members the reference reads into the SoA blocks of include/olf_abi.h, call the
// sm_100a library and write the result back exactly where the reference writes it (Frame::mvpMapPoints, match12).
// Inside the reference tree the class declaration is the reference's own header and ORBmatcher_hot.cc replaces the four
// function bodies of src/ORBmatcher.cc.
#pragma once
#ifdef OLF_IN_REFERENCE_TREE
#include "ORBmatcher.h"
#else
#include <map>
#include <vector>
#include "Frame.h"
namespace ORB_SLAM2 {
class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, std::map<int, int>& match12);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    static int device;                       // CUDA device of the matchers (default 0)
protected:
    float RadiusByViewingCos(const float& viewCos);
    float mfNNratio;
    bool mbCheckOrientation;
};
}  // namespace ORB_SLAM2
#endif
