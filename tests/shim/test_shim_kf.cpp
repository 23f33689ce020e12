// Drives the LocalMapping / LoopClosing / relocalisation overloads of the shim's ORB_SLAM2::ORBmatcher (orb_line_slam_b200/shim/ORBmatcher_kf.cc,
// which calls libolf.so on the GPU) through the same file protocol and the same commands as oracle/_ref/refcli (the reference's own function
// text), so that one set of Python cases (tests/kf_cases.py) checks both against prologue + oracle.
// usage: test_shim_kf <command> <in.bin> <out.bin>; files = int32 count, then per array: int32 dtype (0 u8, 1 i32, 2 f32, 3 f64), int32 ndim,
// int64 dims[ndim], raw data.
#include "../../orb_line_slam_b200/shim/olf_ref_classes.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace ORB_SLAM2;
typedef unsigned char uchar;

struct Arr { int dtype = 0; std::vector<long long> dims; std::vector<char> data;
    size_t count() const { size_t c = 1; for (auto d : dims) c *= (size_t)d; return c; }
    template <typename T> T* as() { return (T*)data.data(); } };
static const int ESZ[4] = {1, 4, 4, 8};
static std::vector<Arr> read_arrays(const char* path) {
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
    int n = 0; if (fread(&n, 4, 1, f) != 1) exit(2);
    std::vector<Arr> v(n);
    for (auto& a : v) {
        int nd = 0; if (fread(&a.dtype, 4, 1, f) != 1 || fread(&nd, 4, 1, f) != 1) exit(2);
        a.dims.resize(nd); if (nd && fread(a.dims.data(), 8, nd, f) != (size_t)nd) exit(2);
        a.data.resize(a.count() * ESZ[a.dtype]); if (!a.data.empty() && fread(a.data.data(), 1, a.data.size(), f) != a.data.size()) exit(2);
    }
    fclose(f); return v;
}
static void write_arrays(const char* path, std::vector<Arr>& v) {
    FILE* f = fopen(path, "wb"); int n = (int)v.size(); fwrite(&n, 4, 1, f);
    for (auto& a : v) { int nd = (int)a.dims.size(); fwrite(&a.dtype, 4, 1, f); fwrite(&nd, 4, 1, f); fwrite(a.dims.data(), 8, nd, f); if (!a.data.empty()) fwrite(a.data.data(), 1, a.data.size(), f); }
    fclose(f);
}
static Arr ints(const std::vector<int>& v) { Arr a; a.dtype = 1; a.dims = {(long long)v.size()}; a.data.resize(v.size() * 4); if (!v.empty()) memcpy(a.data.data(), v.data(), v.size() * 4); return a; }
static cv::Mat mat_f32(const float* p, int r, int c) { cv::Mat m(r, c, CV_32F); for (int i = 0; i < r; ++i) memcpy(m.ptr(i), p + (size_t)i * c, (size_t)c * 4); return m; }

struct Scales { std::vector<float> s, sig, inv_sig; float log_sf; };
static Scales scales(int nlevels, float sf) {            // src/ORBextractor.cc:417-433, src/Frame.cc:149
    Scales S; S.s.assign(nlevels, 1.f); S.sig.assign(nlevels, 1.f); S.inv_sig.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) { S.s[i] = S.s[i - 1] * sf; S.sig[i] = S.s[i] * S.s[i]; }
    for (int i = 0; i < nlevels; ++i) S.inv_sig[i] = 1.0f / S.sig[i];
    S.log_sf = std::log(sf);
    return S;
}
static std::vector<cv::KeyPoint> keypoints_from(Arr& a) {
    std::vector<cv::KeyPoint> v(a.dims[0]);
    for (size_t i = 0; i < v.size(); ++i) { const float* f = a.as<float>() + i * 6; v[i].pt.x = f[0]; v[i].pt.y = f[1]; v[i].size = f[2]; v[i].angle = f[3]; v[i].response = f[4]; v[i].octave = ((const int*)f)[5]; }
    return v;
}
static cv::Mat desc_from(const uchar* p, int n) { cv::Mat m(n, 32, CV_8UC1); for (int r = 0; r < n; ++r) memcpy(m.ptr(r), p + (size_t)r * 32, 32); return m; }
// cam = fx fy cx cy bf w h; pose = Rcw[9] tcw[3] Ow[3] or null
static void fill_keyframe(KeyFrame& K, const Scales& S, Arr& kps, const uchar* desc, Arr* uright, Arr& cam, const float* pose) {
    K.mvKeysUn = keypoints_from(kps); K.N = (int)K.mvKeysUn.size(); K.mDescriptors = desc_from(desc, K.N);
    K.mvuRight.assign(K.N, -1.f); if (uright && uright->count()) K.mvuRight.assign(uright->as<float>(), uright->as<float>() + uright->count());
    const float* c = cam.as<float>();
    K.fx = c[0]; K.fy = c[1]; K.cx = c[2]; K.cy = c[3]; K.mbf = c[4]; K.mnMinX = 0; K.mnMinY = 0; K.mnMaxX = (int)c[5]; K.mnMaxY = (int)c[6];
    K.mvScaleFactors = S.s; K.mvLevelSigma2 = S.sig; K.mvInvLevelSigma2 = S.inv_sig; K.mnScaleLevels = (int)S.s.size(); K.mfLogScaleFactor = S.log_sf;
    K.mvpMapPoints.assign(K.N, static_cast<MapPoint*>(NULL));
    if (pose) { K.Rcw = mat_f32(pose, 3, 3); K.tcw = mat_f32(pose + 9, 3, 1); K.Ow = mat_f32(pose + 12, 3, 1); }
}
static void fill_points(std::vector<MapPoint>& pts, Arr& pos, Arr* normal, Arr& maxd, Arr& mind, Arr& desc, Arr* obs) {
    const int n = (int)maxd.count();
    pts = std::vector<MapPoint>(n);
    for (int i = 0; i < n; ++i) {
        MapPoint& m = pts[i];
        m.pos = mat_f32(pos.as<float>() + 3 * i, 3, 1);
        if (normal && normal->count()) m.normal = mat_f32(normal->as<float>() + 3 * i, 3, 1);
        m.mfMaxDistance = maxd.as<float>()[i]; m.mfMinDistance = mind.as<float>()[i];
        m.desc = desc_from(desc.as<uchar>() + (size_t)32 * i, 1);
        m.nobs = obs && obs->count() ? obs->as<int>()[i] : 1;
    }
}
static void fill_featvec(DBoW2::FeatureVector& fv, Arr& node, Arr& begin, Arr& index) {
    for (size_t a = 0; a < node.count(); ++a) {
        std::vector<unsigned int>& v = fv[(DBoW2::NodeId)node.as<int>()[a]];
        for (int p = begin.as<int>()[a]; p < begin.as<int>()[a + 1]; ++p) v.push_back((unsigned)index.as<int>()[p]);
    }
}
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: test_shim_kf <command> <in.bin> <out.bin>\n"); return 2; }
    const std::string cmd = argv[1];
    std::vector<Arr> in = read_arrays(argv[2]), out;
    try {
        if (cmd == "fuse" || cmd == "fuse_sim3" || cmd == "sbp_kf") {
            const Scales S = scales(in[13].as<int>()[1], in[14].as<float>()[0]);
            KeyFrame KF; fill_keyframe(KF, S, in[0], in[1].as<uchar>(), &in[2], in[3], cmd == "fuse" ? in[4].as<float>() : nullptr);
            std::vector<MapPoint> pts; fill_points(pts, in[5], &in[6], in[7], in[8], in[9], &in[10]);
            const int np = (int)pts.size();
            std::vector<MapPoint*> vp(np); for (int i = 0; i < np; ++i) vp[i] = &pts[i];
            const float th = in[12].as<float>()[0];
            ORBmatcher matcher(0.8, true);
            std::vector<int> res(np, -1); int n = 0;
            if (cmd == "fuse") {
                n = matcher.Fuse(&KF, vp, th);
                for (int i = 0; i < np; ++i) res[i] = pts[i].added_idx;
            } else if (cmd == "fuse_sim3") {
                std::vector<MapPoint*> repl(np, static_cast<MapPoint*>(NULL));
                n = matcher.Fuse(&KF, mat_f32(in[4].as<float>(), 4, 4), vp, th, repl);
                for (int i = 0; i < np; ++i) res[i] = repl[i] ? repl[i]->added_idx : pts[i].added_idx;
            } else {
                MapPoint marker;
                std::vector<MapPoint*> matched(KF.N, static_cast<MapPoint*>(NULL));
                for (int j = 0; j < KF.N && in[11].count(); ++j) if (in[11].as<uchar>()[j]) matched[j] = &marker;
                n = matcher.SearchByProjection(&KF, mat_f32(in[4].as<float>(), 4, 4), vp, matched, (int)th);
                for (int j = 0; j < KF.N; ++j) if (matched[j] && matched[j] != &marker) res[(int)(matched[j] - pts.data())] = j;
            }
            out.push_back(ints(res)); out.push_back(ints({n}));
        } else if (cmd == "sim3") {
            const Scales S = scales(in[19].as<int>()[1], in[20].as<float>()[0]);
            KeyFrame K1, K2; fill_keyframe(K1, S, in[0], in[1].as<uchar>(), nullptr, in[16], in[2].as<float>()); fill_keyframe(K2, S, in[8], in[9].as<uchar>(), nullptr, in[16], in[10].as<float>());
            std::vector<MapPoint> p1, p2; fill_points(p1, in[4], nullptr, in[5], in[6], in[7], nullptr); fill_points(p2, in[12], nullptr, in[13], in[14], in[15], nullptr);
            for (int i = 0; i < K1.N; ++i) if (in[3].as<uchar>()[i]) K1.mvpMapPoints[i] = &p1[i];
            for (int i = 0; i < K2.N; ++i) if (in[11].as<uchar>()[i]) K2.mvpMapPoints[i] = &p2[i];
            const float* sm = in[17].as<float>();
            std::vector<MapPoint*> m12(K1.N, static_cast<MapPoint*>(NULL));
            ORBmatcher matcher(0.75, true);
            const float s12 = sm[0];
            const int n = matcher.SearchBySim3(&K1, &K2, m12, s12, mat_f32(sm + 1, 3, 3), mat_f32(sm + 10, 3, 1), in[18].as<float>()[0]);
            std::vector<int> res(K1.N, -1);
            for (int i = 0; i < K1.N; ++i) if (m12[i]) res[i] = (int)(m12[i] - p2.data());
            out.push_back(ints(res)); out.push_back(ints({n}));
        } else if (cmd == "reloc") {
            const Scales S = scales(in[12].as<int>()[1], in[13].as<float>()[0]);
            Frame Cur;
            Cur.mvKeysUn = keypoints_from(in[0]); Cur.mvKeys = Cur.mvKeysUn; Cur.N = (int)Cur.mvKeysUn.size(); Cur.mDescriptors = desc_from(in[1].as<uchar>(), Cur.N);
            const float* c = in[2].as<float>();
            Cur.fx = c[0]; Cur.fy = c[1]; Cur.cx = c[2]; Cur.cy = c[3]; Cur.mbf = c[4]; Cur.mnMinX = 0; Cur.mnMaxX = c[5]; Cur.mnMinY = 0; Cur.mnMaxY = c[6];
            Cur.mvScaleFactors = S.s; Cur.mnScaleLevels = (int)S.s.size(); Cur.mfLogScaleFactor = S.log_sf;
            const float* ps = in[3].as<float>();
            Cur.mTcw = cv::Mat(4, 4, CV_32F); memset(Cur.mTcw.data, 0, 64); Cur.mTcw.at<float>(3, 3) = 1.f;
            for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) Cur.mTcw.at<float>(r, cc) = ps[3 * r + cc]; Cur.mTcw.at<float>(r, 3) = ps[9 + r]; }
            MapPoint marker;
            Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL));
            for (int j = 0; j < Cur.N; ++j) if (in[4].as<uchar>()[j]) Cur.mvpMapPoints[j] = &marker;
            KeyFrame KF; fill_keyframe(KF, S, in[5], in[10].as<uchar>(), nullptr, in[2], nullptr);
            std::vector<MapPoint> pts; fill_points(pts, in[7], nullptr, in[8], in[9], in[10], nullptr);
            for (int i = 0; i < KF.N; ++i) if (in[6].as<uchar>()[i]) KF.mvpMapPoints[i] = &pts[i];
            ORBmatcher matcher(0.9, false);
            std::set<MapPoint*> found;
            const int n = matcher.SearchByProjection(Cur, &KF, found, in[11].as<float>()[0], (int)in[11].as<float>()[1]);
            std::vector<int> res(Cur.N, -1);
            for (int j = 0; j < Cur.N; ++j) if (Cur.mvpMapPoints[j] && Cur.mvpMapPoints[j] != &marker) res[j] = (int)(Cur.mvpMapPoints[j] - pts.data());
            out.push_back(ints(res)); out.push_back(ints({n}));
        } else if (cmd == "init") {
            Frame F1, F2;
            auto fill = [&](Frame& F, Arr& kps, Arr& desc) {
                F.mvKeysUn = keypoints_from(kps); F.mvKeys = F.mvKeysUn; F.N = (int)F.mvKeysUn.size(); F.mDescriptors = desc_from(desc.as<uchar>(), F.N);
                const float* c = in[4].as<float>();
                F.fx = c[0]; F.fy = c[1]; F.cx = c[2]; F.cy = c[3]; F.mbf = c[4]; F.mnMinX = 0; F.mnMaxX = c[5]; F.mnMinY = 0; F.mnMaxY = c[6];
            };
            fill(F1, in[0], in[1]); fill(F2, in[2], in[3]);
            std::vector<cv::Point2f> prev(F1.N);
            for (int i = 0; i < F1.N; ++i) prev[i] = cv::Point2f(in[5].as<float>()[2 * i], in[5].as<float>()[2 * i + 1]);
            const float* par = in[6].as<float>();
            ORBmatcher matcher(par[1], par[2] != 0);
            std::vector<int> m12;
            const int n = matcher.SearchForInitialization(F1, F2, prev, m12, (int)par[0]);
            std::vector<float> pv(2 * (size_t)F1.N + 2);
            for (int i = 0; i < F1.N; ++i) { pv[2 * i] = prev[i].x; pv[2 * i + 1] = prev[i].y; }
            out.push_back(ints(m12)); out.push_back(ints({n}));
            Arr a; a.dtype = 2; a.dims = {(long long)F1.N, 2}; a.data.resize((size_t)F1.N * 8); if (F1.N) memcpy(a.data.data(), pv.data(), a.data.size()); out.push_back(a);
        } else if (cmd == "triangulation" || cmd == "bow_kf") {
            const Scales S = scales(in[17].as<int>()[1], in[18].as<float>()[0]);
            const float* geo = in[15].as<float>(); const float* fl = in[16].as<float>();
            float pose1[15] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, geo[0], geo[1], geo[2]}, pose2[15];
            memcpy(pose2, geo + 3, 12 * sizeof(float)); pose2[12] = pose2[13] = pose2[14] = 0;
            KeyFrame K1, K2; fill_keyframe(K1, S, in[0], in[1].as<uchar>(), &in[3], in[14], pose1); fill_keyframe(K2, S, in[7], in[8].as<uchar>(), &in[10], in[14], pose2);
            fill_featvec(K1.mFeatVec, in[4], in[5], in[6]); fill_featvec(K2.mFeatVec, in[11], in[12], in[13]);
            MapPoint marker;
            std::vector<int> res(K1.N, -1); int n = 0;
            ORBmatcher matcher(fl[2], fl[1] != 0);
            if (cmd == "triangulation") {
                for (int i = 0; i < K1.N; ++i) if (in[2].as<uchar>()[i]) K1.mvpMapPoints[i] = &marker;
                for (int i = 0; i < K2.N; ++i) if (in[9].as<uchar>()[i]) K2.mvpMapPoints[i] = &marker;
                std::vector<std::pair<size_t, size_t>> pairs;
                n = matcher.SearchForTriangulation(&K1, &K2, mat_f32(geo + 15, 3, 3), pairs, fl[0] != 0);
                for (auto& pr : pairs) res[pr.first] = (int)pr.second;
            } else {
                std::vector<MapPoint> m2(K2.N);
                for (int i = 0; i < K1.N; ++i) if (!in[2].as<uchar>()[i]) K1.mvpMapPoints[i] = &marker;
                for (int i = 0; i < K2.N; ++i) if (!in[9].as<uchar>()[i]) K2.mvpMapPoints[i] = &m2[i];
                std::vector<MapPoint*> m12;
                n = matcher.SearchByBoW(&K1, &K2, m12);
                for (int i = 0; i < K1.N; ++i) if (m12[i]) res[i] = (int)(m12[i] - m2.data());
            }
            out.push_back(ints(res)); out.push_back(ints({n}));
        } else { fprintf(stderr, "test_shim_kf: unknown command %s\n", cmd.c_str()); return 2; }
    } catch (const std::exception& e) { fprintf(stderr, "EXCEPTION %s\n", e.what()); return 1; }
    write_arrays(argv[3], out);
    return 0;
}
