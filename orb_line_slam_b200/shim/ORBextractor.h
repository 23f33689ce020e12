// Drop-in for include/ORBextractor.h:52-118 of the reference: same class name, namespace, constructor and call
// signatures, getters and the public mvImagePyramid member, so src/Frame.cc and src/Tracking.cc compile unchanged.
// All computation is forwarded to the sm_100a library through the C-ABI (include/olf_abi.h).
#pragma once
#include <vector>
#include "cv_min.h"
#include "../../include/olf_abi.h"

namespace ORB_SLAM2 {
class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; the mask is ignored (as in the reference).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Frame::ComputeStereoMatches reads this in the reference; with the shim it stays empty until
    // SyncImagePyramid() is called (the stereo matcher works on the device-resident pyramid instead).
    std::vector<cv::Mat> mvImagePyramid;
    void SyncImagePyramid();

    olf_orb* handle() { return h_; }          // for olf_stereo_points
    static int device;                        // CUDA device used by new extractors (default 0)
protected:
    int nfeatures; double scaleFactor; int nlevels; int iniThFAST; int minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    olf_orb* h_;
};
}  // namespace ORB_SLAM2
