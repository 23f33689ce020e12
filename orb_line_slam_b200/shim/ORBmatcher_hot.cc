// The hot overloads of ORB_SLAM2::ORBmatcher on the B200 (replaces src/ORBmatcher.cc:47-131, 161-290, 1330-1472, 1474-1618).
#include "olf_ref_classes.h"
#include "../../include/olf_abi.h"
#include <cstring>
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {
#ifndef OLF_IN_REFERENCE_TREE
const int ORBmatcher::TH_HIGH = 100;           // src/ORBmatcher.cc:39-41
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
int ORBmatcher::device = 0;
ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }     // :133-139 (applied inside the library)
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {                                 // :1795-1811, one pair: host
    const uint32_t* pa = a.ptr<uint32_t>(); const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}
void ORBmatcher::ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {      // :1749-1790 (stays in src/ORBmatcher.cc in the reference tree)
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}
#define OLF_MATCHER_DEVICE ORBmatcher::device
#else
#ifndef OLF_MATCHER_DEVICE
#define OLF_MATCHER_DEVICE 0
#endif
#endif

namespace {
std::vector<olf_keypoint> pack_kps(const std::vector<cv::KeyPoint>& k) {
    std::vector<olf_keypoint> v(k.size());
    for (size_t i = 0; i < k.size(); ++i) v[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
    return v;
}
std::vector<uint8_t> pack_rows(const cv::Mat& d) {
    std::vector<uint8_t> v((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; ++i) memcpy(v.data() + (size_t)i * 32, d.ptr(i), 32);
    return v;
}
olf_camera camera_of(const Frame& F) { return olf_camera{F.fx, F.fy, F.cx, F.cy, F.mbf, F.mnMinX, F.mnMaxX, F.mnMinY, F.mnMaxY}; }
void pose_of(const cv::Mat& Tcw, float* R, float* t) {
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = Tcw.at<float>(r, c); t[r] = Tcw.at<float>(r, 3); }
}
std::vector<uint8_t> occupied_of(Frame& F) {                  // F.mvpMapPoints[idx] && Observations() > 0
    std::vector<uint8_t> o(F.N, 0);
    for (int j = 0; j < F.N && j < (int)F.mvpMapPoints.size(); ++j) if (F.mvpMapPoints[j] && F.mvpMapPoints[j]->Observations() > 0) o[j] = 1;
    return o;
}
}  // namespace

// src/ORBmatcher.cc:47-131
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    std::vector<int> src; std::vector<float> px, py, pxr, vc; std::vector<int> lvl; std::vector<uint8_t> obs, pdesc;
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView) continue;                      // :55-59
        if (pMP->isBad()) continue;
        src.push_back((int)iMP);
        px.push_back(pMP->mTrackProjX); py.push_back(pMP->mTrackProjY); pxr.push_back(pMP->mTrackProjXR);
        lvl.push_back(pMP->mnTrackScaleLevel); vc.push_back(pMP->mTrackViewCos); obs.push_back(pMP->Observations() > 0);
        const cv::Mat d = pMP->GetDescriptor();
        pdesc.insert(pdesc.end(), d.ptr(0), d.ptr(0) + 32);
    }
    const std::vector<olf_keypoint> ck = pack_kps(F.mvKeysUn);
    const std::vector<uint8_t> cd = pack_rows(F.mDescriptors), occ = occupied_of(F);
    olf_sbp_map_args a; memset(&a, 0, sizeof(a));
    a.cur_kps = ck.data(); a.cur_desc = cd.data(); a.cur_u_right = F.mvuRight.data(); a.n_cur = F.N; a.cur_occupied = occ.data();
    a.cam = camera_of(F); a.scale_factors = F.mvScaleFactors.data(); a.nlevels = F.mnScaleLevels;
    a.n_points = (int)src.size(); a.proj_x = px.data(); a.proj_y = py.data(); a.proj_xr = pxr.data(); a.pred_level = lvl.data();
    a.view_cos = vc.data(); a.point_observed = obs.data(); a.point_desc = pdesc.data(); a.th = th; a.nn_ratio = mfNNratio;
    std::vector<int> assigned(src.size(), -1); int n = 0;
    if (olf_search_by_projection_map(&a, assigned.data(), &n, OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[SearchByProjection] ") + olf_last_error());
    for (size_t i = 0; i < src.size(); ++i) if (assigned[i] >= 0) F.mvpMapPoints[assigned[i]] = vpMapPoints[src[i]];      // :123 in point order
    return n;
}

// src/ORBmatcher.cc:1474-1618 (and :1330-1472, the same without match12)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, std::map<int, int>& match12) {
    match12.clear();
    const int nl = LastFrame.N;
    std::vector<uint8_t> has(nl, 0), obs(nl, 0), pdesc((size_t)nl * 32, 0); std::vector<float> pos((size_t)nl * 3, 0.f);
    for (int i = 0; i < nl; ++i) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;          // :1503-1507
        has[i] = 1; obs[i] = pMP->Observations() > 0;
        const cv::Mat w = pMP->GetWorldPos(), d = pMP->GetDescriptor();
        for (int r = 0; r < 3; ++r) pos[3 * i + r] = w.at<float>(r, 0);
        memcpy(&pdesc[(size_t)32 * i], d.ptr(0), 32);
    }
    const std::vector<olf_keypoint> ck = pack_kps(CurrentFrame.mvKeysUn), lk = pack_kps(LastFrame.mvKeysUn);
    const std::vector<uint8_t> cd = pack_rows(CurrentFrame.mDescriptors), occ = occupied_of(CurrentFrame);
    olf_sbp_last_args a; memset(&a, 0, sizeof(a));
    a.cur_kps = ck.data(); a.cur_desc = cd.data(); a.cur_u_right = CurrentFrame.mvuRight.data(); a.n_cur = CurrentFrame.N;
    a.cam = camera_of(CurrentFrame); a.scale_factors = CurrentFrame.mvScaleFactors.data(); a.nlevels = CurrentFrame.mnScaleLevels;
    pose_of(CurrentFrame.mTcw, a.Rcw, a.tcw); pose_of(LastFrame.mTcw, a.Rlw, a.tlw);
    a.last_kps = lk.data(); a.n_last = nl; a.last_has_point = has.data(); a.last_point_observed = obs.data();
    a.last_world_pos = pos.data(); a.last_point_desc = pdesc.data();
    a.th = th; a.mono = bMono; a.check_orientation = mbCheckOrientation; a.cur_occupied = occ.data();
    std::vector<int> assigned(nl, -1), cur_point(CurrentFrame.N, -1); int n = 0;
    if (olf_search_by_projection_last(&a, assigned.data(), cur_point.data(), &n, OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[SearchByProjection] ") + olf_last_error());
    // the reference writes CurrentFrame.mvpMapPoints[bestIdx2] = pMP for every accepted point (:1575) and resets the keypoints of
    // the pruned rotation bins to NULL (:1610); match12 holds (keypoint, last index) of the surviving ones (:1577, :1612)
    // (std::map::insert keeps the FIRST point that took a keypoint, mvpMapPoints the last one)
    for (int i = 0; i < nl; ++i) if (assigned[i] >= 0) { CurrentFrame.mvpMapPoints[assigned[i]] = static_cast<MapPoint*>(NULL); match12.insert(std::pair<int, int>(assigned[i], i)); }
    for (int j = 0; j < CurrentFrame.N; ++j) {
        if (cur_point[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[cur_point[j]];
        else match12.erase(j);
    }
    return n;
}
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    std::map<int, int> match12;
    return SearchByProjection(CurrentFrame, LastFrame, th, bMono, match12);
}

// src/ORBmatcher.cc:161-290
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    auto csr = [](const DBoW2::FeatureVector& fv, std::vector<int>& node, std::vector<int>& begin, std::vector<int>& index) {
        begin.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node.push_back((int)it->first);
            for (size_t k = 0; k < it->second.size(); ++k) index.push_back((int)it->second[k]);
            begin.push_back((int)index.size());
        }
    };
    std::vector<int> kn, kb, ki, fn, fb, fi;
    csr(pKF->mFeatVec, kn, kb, ki); csr(F.mFeatVec, fn, fb, fi);
    const int nkf = (int)vpMapPointsKF.size();
    std::vector<uint8_t> has(nkf, 0);
    for (int i = 0; i < nkf; ++i) if (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) has[i] = 1;      // :196-201
    const std::vector<olf_keypoint> kk = pack_kps(pKF->mvKeysUn), fk = pack_kps(F.mvKeys);               // :249 reads pKF->mvKeysUn and F.mvKeys
    const std::vector<uint8_t> kd = pack_rows(pKF->mDescriptors), fd = pack_rows(F.mDescriptors);
    olf_bow_match_args a; memset(&a, 0, sizeof(a));
    a.kf_desc = kd.data(); a.kf_kps_un = kk.data(); a.n_kf = nkf; a.kf_has_point = has.data();
    a.kf_fv_node = kn.data(); a.kf_fv_begin = kb.data(); a.kf_fv_index = ki.data(); a.kf_n_nodes = (int)kn.size();
    a.f_desc = fd.data(); a.f_kps = fk.data(); a.n_f = F.N;
    a.f_fv_node = fn.data(); a.f_fv_begin = fb.data(); a.f_fv_index = fi.data(); a.f_n_nodes = (int)fn.size();
    a.nn_ratio = mfNNratio; a.check_orientation = mbCheckOrientation;
    std::vector<int> match_f(F.N, -1); int n = 0;
    if (olf_search_by_bow(&a, match_f.data(), &n, OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[SearchByBoW] ") + olf_last_error());
    for (int j = 0; j < F.N; ++j) if (match_f[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[match_f[j]];
    return n;
}
}  // namespace ORB_SLAM2
