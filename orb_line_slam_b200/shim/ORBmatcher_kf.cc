// The LocalMapping / LoopClosing / relocalisation / initialisation overloads of ORB_SLAM2::ORBmatcher on the B200 (replaces
// src/ORBmatcher.cc:292-1328, 1620-1747).  Every overload keeps its O(#points) host prologue -- cv::Mat projection, depth / viewing-angle tests,
// MapPoint::PredictScale: the very calls the reference makes, so OpenCV's arithmetic is the reference's by construction -- and hands
// the data-parallel part (grid window x octave window x Hamming distance, first minimum, blocking) to the sm_100a library:
// olf_window_search, olf_search_for_triangulation, olf_search_by_bow_kf (include/olf_abi.h).  The epilogues (AddObservation /
// Replace / vpMatched / rotation histogram) run on the host in the reference's order.
#include "olf_ref_classes.h"
#include "../../include/olf_abi.h"
#include <climits>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#ifdef OLF_IN_REFERENCE_TREE
#ifndef OLF_MATCHER_DEVICE
#define OLF_MATCHER_DEVICE 0
#endif
#else
#define OLF_MATCHER_DEVICE ORBmatcher::device
#endif

namespace ORB_SLAM2 {
namespace {

// what differs between the five projection overloads
enum : unsigned {
    NEED_POSITIVE_DEPTH = 1u,     // `if (p3Dc.at<float>(2) < 0) continue` (all but the relocalisation overload)
    INVZ_IN_DOUBLE = 2u,          // `1.0 / z` (double division, then float) instead of `1 / z` (float division)
    VIEW_ANGLE_TEST = 4u,         // PO.dot(Pn) < 0.5 * dist3D
    DIST_IN_CAMERA = 8u,          // dist3D = |p3Dc| (SearchBySim3) instead of |p3Dw - Ow|
    FRAME_BOUNDS = 16u,           // inclusive Frame bounds and u = fx * xc * invzc + cx (relocalisation) instead of KeyFrame::IsInImage
};

struct Queries {
    std::vector<int> src; std::vector<float> u, v, radius, ur; std::vector<int> lo, hi; std::vector<uint8_t> desc;
    void add(int i, float u_, float v_, float r, float ur_, int lo_, int hi_, const cv::Mat& d) {
        src.push_back(i); u.push_back(u_); v.push_back(v_); radius.push_back(r); ur.push_back(ur_); lo.push_back(lo_); hi.push_back(hi_);
        desc.insert(desc.end(), d.ptr(0), d.ptr(0) + 32);
    }
};

struct Target {                   // the searched key frame / frame in the library's layout
    std::vector<olf_keypoint> kps; std::vector<uint8_t> desc; olf_camera cam; int n;
    Target(const std::vector<cv::KeyPoint>& k, const cv::Mat& d, float fx, float fy, float cx, float cy, float bf, float minx, float maxx, float miny, float maxy)
        : kps(k.size()), desc((size_t)d.rows * 32), cam{fx, fy, cx, cy, bf, minx, maxx, miny, maxy}, n((int)k.size()) {
        for (size_t i = 0; i < k.size(); ++i) kps[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
        for (int i = 0; i < d.rows; ++i) memcpy(desc.data() + (size_t)i * 32, d.ptr(i), 32);
    }
};
Target target_of(KeyFrame* pKF) {
    return Target(pKF->mvKeysUn, pKF->mDescriptors, pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf, (float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY);
}

// The per-point prologue (src/ORBmatcher.cc:854-892, 1011-1051, 318-342, 1160-1191, 1645-1677).  p3Dc: the point in the camera frame of
// `host`; PO: p3Dw - Ow (unused with DIST_IN_CAMERA).  Returns false where the reference `continue`s.
template <class Host>
bool project(MapPoint* pMP, const cv::Mat& p3Dc, const cv::Mat& PO, Host* host, float fx, float fy, float cx, float cy, float bf,
             const std::vector<float>& scale, unsigned flags, float th, int level_above, int src, Queries& Q) {
    const float z = p3Dc.at<float>(2);
    if ((flags & NEED_POSITIVE_DEPTH) && z < 0.0f) return false;
    const float invz = (flags & INVZ_IN_DOUBLE) ? (float)(1.0 / z) : 1 / z;
    float u, v;
    if (flags & FRAME_BOUNDS) { u = fx * p3Dc.at<float>(0) * invz + cx; v = fy * p3Dc.at<float>(1) * invz + cy; }
    else { const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz; u = fx * x + cx; v = fy * y + cy; }
    if (flags & FRAME_BOUNDS) { if (u < host->mnMinX || u > host->mnMaxX) return false; if (v < host->mnMinY || v > host->mnMaxY) return false; }
    else if (!(u >= host->mnMinX && u < host->mnMaxX && v >= host->mnMinY && v < host->mnMaxY)) return false;          // KeyFrame::IsInImage
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const float dist3D = (flags & DIST_IN_CAMERA) ? cv::norm(p3Dc) : cv::norm(PO);
    if (dist3D < minDistance || dist3D > maxDistance) return false;
    if (flags & VIEW_ANGLE_TEST) { cv::Mat Pn = pMP->GetNormal(); if (PO.dot(Pn) < 0.5 * dist3D) return false; }
    const int nPredictedLevel = pMP->PredictScale(dist3D, host);
    Q.add(src, u, v, th * scale[nPredictedLevel], u - bf * invz, nPredictedLevel - 1, nPredictedLevel + level_above, pMP->GetDescriptor());
    return true;
}

void window_search(const Target& T, const Queries& Q, int max_dist, const uint8_t* blocked, bool sequential, const float* u_right, const float* inv_sigma2,
                   int nlevels, std::vector<int>& best_idx, std::vector<int>& best_dist, const char* who) {
    olf_window_search_args a; memset(&a, 0, sizeof(a));
    a.kps = T.kps.data(); a.desc = T.desc.data(); a.n = T.n; a.cam = T.cam;
    a.u_right = u_right; a.inv_level_sigma2 = inv_sigma2; a.nlevels = nlevels; a.chi2_check = inv_sigma2 != nullptr;
    a.n_queries = (int)Q.src.size(); a.u = Q.u.data(); a.v = Q.v.data(); a.radius = Q.radius.data(); a.ur = Q.ur.data();
    a.min_level = Q.lo.data(); a.max_level = Q.hi.data(); a.qdesc = Q.desc.data();
    a.blocked = blocked; a.sequential_blocking = sequential; a.max_dist = max_dist;
    best_idx.assign(Q.src.size() + 1, -1); best_dist.assign(Q.src.size() + 1, 256);
    if (olf_window_search(&a, best_idx.data(), best_dist.data(), OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[") + who + "] " + olf_last_error());
}

void decompose_sim3(const cv::Mat& Scw, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow) {      // :301-305, :988-992
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw / scw;
    tcw = Scw.rowRange(0, 3).col(3) / scw;
    Ow = -Rcw.t() * tcw;
}

void csr_of(const DBoW2::FeatureVector& fv, std::vector<int>& node, std::vector<int>& begin, std::vector<int>& index) {
    begin.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        node.push_back((int)it->first);
        for (size_t k = 0; k < it->second.size(); ++k) index.push_back((int)it->second[k]);
        begin.push_back((int)index.size());
    }
}
}  // namespace

// src/ORBmatcher.cc:827-977
int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    const Target T = target_of(pKF);
    Queries Q;
    const int nMPs = (int)vpMapPoints.size();
    for (int i = 0; i < nMPs; i++) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        project(pMP, Rcw * p3Dw + tcw, p3Dw - Ow, pKF, pKF->fx, pKF->fy, pKF->cx, pKF->cy, pKF->mbf, pKF->mvScaleFactors, NEED_POSITIVE_DEPTH | VIEW_ANGLE_TEST, th, 0, i, Q);
    }
    std::vector<int> best, dist;
    window_search(T, Q, TH_LOW, nullptr, false, pKF->mvuRight.data(), pKF->mvInvLevelSigma2.data(), (int)pKF->mvInvLevelSigma2.size(), best, dist, "Fuse");
    // the reference's epilogue in point order (:953-973).  A point that an earlier Replace turned bad, or that entered the key frame meanwhile,
    // is skipped exactly where the reference's loop head would skip it.
    int nFused = 0;
    for (size_t q = 0; q < Q.src.size(); ++q) {
        MapPoint* pMP = vpMapPoints[Q.src[q]];
        if (best[q] < 0 || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[q]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, best[q]);
            pKF->AddMapPoint(pMP, best[q]);
        }
        nFused++;
    }
    return nFused;
}

// src/ORBmatcher.cc:979-1102
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    cv::Mat Rcw, tcw, Ow;
    decompose_sim3(Scw, Rcw, tcw, Ow);
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    const Target T = target_of(pKF);
    Queries Q;
    const int nPoints = (int)vpPoints.size();
    for (int iMP = 0; iMP < nPoints; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        project(pMP, Rcw * p3Dw + tcw, p3Dw - Ow, pKF, pKF->fx, pKF->fy, pKF->cx, pKF->cy, 0.f, pKF->mvScaleFactors, NEED_POSITIVE_DEPTH | INVZ_IN_DOUBLE | VIEW_ANGLE_TEST, th, 0, iMP, Q);
    }
    std::vector<int> best, dist;
    window_search(T, Q, TH_LOW, nullptr, false, nullptr, nullptr, 0, best, dist, "Fuse");
    int nFused = 0;
    for (size_t q = 0; q < Q.src.size(); ++q) {
        if (best[q] < 0) continue;
        MapPoint* pMP = vpPoints[Q.src[q]];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[q]);
        if (pMPinKF) { if (!pMPinKF->isBad()) vpReplacePoint[Q.src[q]] = pMPinKF; }
        else { pMP->AddObservation(pKF, best[q]); pKF->AddMapPoint(pMP, best[q]); }
        nFused++;
    }
    return nFused;
}

// src/ORBmatcher.cc:292-405
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
    cv::Mat Rcw, tcw, Ow;
    decompose_sim3(Scw, Rcw, tcw, Ow);
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    const Target T = target_of(pKF);
    Queries Q;
    for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        project(pMP, Rcw * p3Dw + tcw, p3Dw - Ow, pKF, pKF->fx, pKF->fy, pKF->cx, pKF->cy, 0.f, pKF->mvScaleFactors, NEED_POSITIVE_DEPTH | VIEW_ANGLE_TEST, (float)th, 0, iMP, Q);
    }
    std::vector<uint8_t> blocked(T.n + 1, 0);
    for (int j = 0; j < T.n && j < (int)vpMatched.size(); ++j) blocked[j] = vpMatched[j] != NULL;
    std::vector<int> best, dist;
    window_search(T, Q, TH_LOW, blocked.data(), true, nullptr, nullptr, 0, best, dist, "SearchByProjection");
    int nmatches = 0;
    for (size_t q = 0; q < Q.src.size(); ++q) if (best[q] >= 0) { vpMatched[best[q]] = vpPoints[Q.src[q]]; nmatches++; }
    return nmatches;
}

// src/ORBmatcher.cc:1104-1328
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th) {
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {
        MapPoint* pMP = vpMatches12[i];
        if (pMP) {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
        }
    }
    // one direction: the map points of `from` go through (Rw, tw) into their own camera and through (sR, t) into the camera of `to`
    auto search = [&](const std::vector<MapPoint*>& pts, const std::vector<bool>& already, const cv::Mat& Rw, const cv::Mat& tw, const cv::Mat& sR, const cv::Mat& t,
                      KeyFrame* to, std::vector<int>& match) {
        const Target T = target_of(to);
        Queries Q;
        for (int i = 0; i < (int)pts.size(); ++i) {
            MapPoint* pMP = pts[i];
            if (!pMP || already[i] || pMP->isBad()) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc_own = Rw * p3Dw + tw;
            project(pMP, sR * p3Dc_own + t, cv::Mat(), to, pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy, 0.f, to->mvScaleFactors, NEED_POSITIVE_DEPTH | INVZ_IN_DOUBLE | DIST_IN_CAMERA, th, 0, i, Q);
        }
        std::vector<int> best, dist;
        window_search(T, Q, TH_HIGH, nullptr, false, nullptr, nullptr, 0, best, dist, "SearchBySim3");
        match.assign(pts.size(), -1);
        for (size_t q = 0; q < Q.src.size(); ++q) match[Q.src[q]] = best[q];
    };
    std::vector<int> vnMatch1, vnMatch2;
    search(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1);
    search(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; i1++) {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
    }
    return nFound;
}

// src/ORBmatcher.cc:1620-1747
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const Target T(CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, CurrentFrame.mbf,
                   CurrentFrame.mnMinX, CurrentFrame.mnMaxX, CurrentFrame.mnMinY, CurrentFrame.mnMaxY);
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    Queries Q;
    for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        project(pMP, Rcw * x3Dw + tcw, x3Dw - Ow, &CurrentFrame, CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, 0.f, CurrentFrame.mvScaleFactors,
                INVZ_IN_DOUBLE | FRAME_BOUNDS, th, 1, (int)i, Q);
    }
    std::vector<uint8_t> blocked(T.n + 1, 0);
    for (int j = 0; j < T.n; ++j) blocked[j] = CurrentFrame.mvpMapPoints[j] != NULL;
    std::vector<int> best, dist;
    window_search(T, Q, ORBdist, blocked.data(), true, nullptr, nullptr, 0, best, dist, "SearchByProjection");
    int nmatches = 0;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    for (size_t q = 0; q < Q.src.size(); ++q) {
        if (best[q] < 0) continue;
        CurrentFrame.mvpMapPoints[best[q]] = vpMPs[Q.src[q]];
        nmatches++;
        if (mbCheckOrientation) {
            float rot = pKF->mvKeysUn[Q.src[q]].angle - CurrentFrame.mvKeysUn[best[q]].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(best[q]);
        }
    }
    if (mbCheckOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL); nmatches--; }
    }
    return nmatches;
}

// src/ORBmatcher.cc:659-825
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo) {
    // epipole in the second image (:666-672)
    cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    olf_triangulation_args a; memset(&a, 0, sizeof(a));
    a.ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    a.ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    const Target T1 = target_of(pKF1), T2 = target_of(pKF2);
    std::vector<uint8_t> skip1(T1.n + 1, 0), skip2(T2.n + 1, 0);
    for (int i = 0; i < T1.n; ++i) skip1[i] = pKF1->GetMapPoint(i) != NULL;
    for (int i = 0; i < T2.n; ++i) skip2[i] = pKF2->GetMapPoint(i) != NULL;
    std::vector<int> n1, b1, i1, n2, b2, i2;
    csr_of(pKF1->mFeatVec, n1, b1, i1); csr_of(pKF2->mFeatVec, n2, b2, i2);
    a.kps1 = T1.kps.data(); a.desc1 = T1.desc.data(); a.n1 = T1.n; a.skip1 = skip1.data(); a.u_right1 = pKF1->mvuRight.data();
    a.fv1_node = n1.data(); a.fv1_begin = b1.data(); a.fv1_index = i1.data(); a.fv1_n_nodes = (int)n1.size();
    a.kps2 = T2.kps.data(); a.desc2 = T2.desc.data(); a.n2 = T2.n; a.skip2 = skip2.data(); a.u_right2 = pKF2->mvuRight.data();
    a.fv2_node = n2.data(); a.fv2_begin = b2.data(); a.fv2_index = i2.data(); a.fv2_n_nodes = (int)n2.size();
    a.scale_factors2 = pKF2->mvScaleFactors.data(); a.level_sigma2_2 = pKF2->mvLevelSigma2.data(); a.nlevels = (int)pKF2->mvScaleFactors.size();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a.F12[3 * r + c] = F12.at<float>(r, c);
    a.only_stereo = bOnlyStereo; a.check_orientation = mbCheckOrientation;
    std::vector<int> vMatches12(T1.n + 1, -1); int nmatches = 0;
    if (olf_search_for_triangulation(&a, vMatches12.data(), &nmatches, OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[SearchForTriangulation] ") + olf_last_error());
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < T1.n; i++) if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)vMatches12[i]));
    return nmatches;
}

// src/ORBmatcher.cc:407-522 (monocular map initialisation)
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize) {
    const Target T1(F1.mvKeysUn, F1.mDescriptors, F1.fx, F1.fy, F1.cx, F1.cy, F1.mbf, F1.mnMinX, F1.mnMaxX, F1.mnMinY, F1.mnMaxY);
    const Target T2(F2.mvKeysUn, F2.mDescriptors, F2.fx, F2.fy, F2.cx, F2.cy, F2.mbf, F2.mnMinX, F2.mnMaxX, F2.mnMinY, F2.mnMaxY);
    std::vector<float> prev(2 * vbPrevMatched.size() + 2);
    for (size_t i = 0; i < vbPrevMatched.size(); ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
    std::vector<int> m12(T1.n + 1, -1); int nmatches = 0;
    if (olf_search_for_initialization(T1.kps.data(), T1.desc.data(), T1.n, T2.kps.data(), T2.desc.data(), T2.n, &T2.cam, prev.data(), windowSize, mfNNratio,
                                      mbCheckOrientation, m12.data(), &nmatches, OLF_MATCHER_DEVICE) != OLF_OK)
        throw std::runtime_error(std::string("[SearchForInitialization] ") + olf_last_error());
    for (int i = 0; i < T1.n; ++i) { vnMatches12[i] = m12[i]; if (m12[i] >= 0) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]); }
    return nmatches;
}

// src/ORBmatcher.cc:524-657
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const Target T1 = target_of(pKF1), T2 = target_of(pKF2);
    std::vector<uint8_t> has1(vpMapPoints1.size() + 1, 0), has2(vpMapPoints2.size() + 1, 0);
    for (size_t i = 0; i < vpMapPoints1.size(); ++i) has1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();
    for (size_t i = 0; i < vpMapPoints2.size(); ++i) has2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
    std::vector<int> n1, b1, i1, n2, b2, i2;
    csr_of(pKF1->mFeatVec, n1, b1, i1); csr_of(pKF2->mFeatVec, n2, b2, i2);
    olf_bow_match_args a; memset(&a, 0, sizeof(a));
    a.kf_desc = T1.desc.data(); a.kf_kps_un = T1.kps.data(); a.n_kf = (int)vpMapPoints1.size(); a.kf_has_point = has1.data();
    a.kf_fv_node = n1.data(); a.kf_fv_begin = b1.data(); a.kf_fv_index = i1.data(); a.kf_n_nodes = (int)n1.size();
    a.f_desc = T2.desc.data(); a.f_kps = T2.kps.data(); a.n_f = (int)vpMapPoints2.size();
    a.f_fv_node = n2.data(); a.f_fv_begin = b2.data(); a.f_fv_index = i2.data(); a.f_n_nodes = (int)n2.size();
    a.nn_ratio = mfNNratio; a.check_orientation = mbCheckOrientation;
    std::vector<int> m12(vpMapPoints1.size() + 1, -1); int nmatches = 0;
    if (olf_search_by_bow_kf(&a, has2.data(), m12.data(), &nmatches, OLF_MATCHER_DEVICE) != OLF_OK) throw std::runtime_error(std::string("[SearchByBoW] ") + olf_last_error());
    for (size_t i = 0; i < vpMapPoints1.size(); ++i) if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];
    return nmatches;
}
}  // namespace ORB_SLAM2
