#include "ORBextractor.h"
#include <cassert>
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {
int ORBextractor::device = 0;

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    h_ = olf_orb_create(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, device);
    if (!h_) throw std::runtime_error(std::string("[ORBextractor] ") + olf_last_error());
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    olf_orb_scale_factors(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data());
    mvImagePyramid.resize(nlevels);
}
ORBextractor::~ORBextractor() { olf_orb_destroy(h_); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                  // src/ORBextractor.cc:1048
    cv::Mat image = _image.getMat();                             // :1051
    assert(image.type() == CV_8UC1);                             // :1052
    const int cap = nfeatures + nfeatures / 4 + 512;             // the quadtree may return a few more than nfeatures
    std::vector<olf_keypoint> kps(cap);
    std::vector<uint8_t> desc((size_t)cap * 32);
    int n = 0;
    const int rc = olf_orb_extract(h_, image.ptr(), image.cols, image.rows, (int)image.step, kps.data(), desc.data(), cap, &n);
    if (rc != OLF_OK) throw std::runtime_error(std::string("[ORBextractor] ") + olf_last_error());
    keypoints.clear(); keypoints.reserve(n);
    if (n == 0) { _descriptors.release(); return; }              // :1066-1067
    _descriptors.create(n, 32, CV_8U);                           // :1070
    cv::Mat descriptors = _descriptors.getMat();                 // :1071
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint kp;
        kp.pt.x = kps[i].x; kp.pt.y = kps[i].y; kp.size = kps[i].size; kp.angle = kps[i].angle;
        kp.response = kps[i].response; kp.octave = kps[i].octave; kp.class_id = -1;
        keypoints.push_back(kp);
        memcpy(descriptors.ptr(i), desc.data() + (size_t)i * 32, 32);
    }
}

void ORBextractor::SyncImagePyramid() {
    for (int l = 0; l < nlevels; ++l) {
        int w = 0, h = 0;
        if (olf_orb_level_size(h_, l, &w, &h) != OLF_OK) return;
        mvImagePyramid[l].create(h, w, CV_8UC1);
        olf_orb_get_level(h_, l, mvImagePyramid[l].ptr(), (int)mvImagePyramid[l].step);
    }
}
}  // namespace ORB_SLAM2
