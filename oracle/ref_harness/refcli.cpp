// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// refcli: the REFERENCE'S OWN hot-path code behind a tiny file protocol, so that tests/test_oracle_vs_ref.py can pin the
// CPU oracle (oracle/*.cpp) to the reference itself.  What is compiled here, unmodified, from /root/reference:
//   whole files : src/ORBextractor.cc, src/LineExtractor.cc, src/LineMatcher.cpp, src/gridStructure.cpp, src/LineIterator.cpp,
//                 src/Config.cpp, Thirdparty/line_descriptor/src/LSDDetector_custom.cpp
//   sliced (oracle/ref_harness/slice.py, verbatim function text): BinaryDescriptor ctor / computeSobel / computeImpl /
//                 computeLBD / binaryConversion (ref_lbd.cpp); ORBmatcher::SearchByProjection (map points; last frame), RadiusByViewingCos,
//                 ComputeThreeMaxima, DescriptorDistance; Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid /
//                 ComputeStereoMatches / ComputeStereoMatches_Lines (+ helpers); MapPoint::ComputeDistinctiveDescriptors
// against the stand-in types of cvstub.hpp / frame_stub.hpp.  Un-vendored OpenCV arithmetic comes from oracle/cvprim.hpp and
// oracle/line.cpp (pinned to cv2 4.13 by the golden vectors).
//
// Determinism: DistributeOctTree sorts (size, ExtractorNode*) pairs, i.e. breaks ties by allocation
// address.  This executable replaces the global operator new with a bump allocator (addresses grow in allocation order,
// nothing is reused), which turns that into "creation order" -- the canonical choice of the oracle and the product.
//
// usage: refcli <command> <in.bin> <out.bin>; files = int32 count, then per array: int32 dtype (0 u8,1 i32,2 f32,3 f64),
// int32 ndim, int64 dims[ndim], raw data.
#include "frame_stub.hpp"
#include "ORBmatcher.h"
#include "LineMatcher.h"
#include "LineExtractor.h"
#include "LineIterator.h"
#include "gridStructure.h"
#include "Config.h"
#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"
#include <new>

// ---- bump allocator ---------------------------------------------------------------------------------------------------
static char* g_arena = nullptr; static size_t g_off = 0; static const size_t ARENA = (size_t)6 << 30;
void* operator new(size_t n) {
    if (!g_arena) { g_arena = (char*)malloc(ARENA); if (!g_arena) abort(); }
    n = (n + 15) & ~(size_t)15;
    if (g_off + n > ARENA) { fprintf(stderr, "refcli: arena exhausted\n"); abort(); }
    void* p = g_arena + g_off; g_off += n; return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

// ---- sliced reference functions -----------------------------------------------------------------------------------------
namespace ORB_SLAM2 {
const int ORBmatcher::TH_HIGH = 100;       // src/ORBmatcher.cc:39-41
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
#include "orbmatcher.inc"
#include "frame.inc"
#include "mappoint.inc"
#include "keyframe.inc"
}

// ---- array file protocol ------------------------------------------------------------------------------------------------
struct Arr { int dtype = 0; std::vector<long long> dims; std::vector<char> data;
    size_t count() const { size_t c = 1; for (auto d : dims) c *= (size_t)d; return c; }
    template <typename T> T* as() { return (T*)data.data(); }
    template <typename T> T at(size_t i) { return ((T*)data.data())[i]; } };
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
template <typename T> static Arr make(int dtype, std::vector<long long> dims, const T* src) {
    Arr a; a.dtype = dtype; a.dims = dims; a.data.resize(a.count() * ESZ[dtype]); if (!a.data.empty()) memcpy(a.data.data(), src, a.data.size()); return a;
}
static cv::Mat mat_u8(Arr& a) { cv::Mat m((int)a.dims[0], (int)a.dims[1], CV_8UC1); for (int r = 0; r < m.rows; ++r) memcpy(m.ptr(r), a.as<uchar>() + (size_t)r * m.cols, m.cols); return m; }
static cv::Mat mat_f32(const float* p, int r, int c) { cv::Mat m(r, c, CV_32F); for (int i = 0; i < r; ++i) memcpy(m.ptr(i), p + (size_t)i * c, (size_t)c * 4); return m; }
using cv::line_descriptor::KeyLine;
static std::vector<KeyLine> keylines_from(Arr& a) {           // rows of 17 floats/ints in olf_keyline order
    std::vector<KeyLine> v(a.dims[0]);
    for (size_t i = 0; i < v.size(); ++i) {
        const float* f = a.as<float>() + i * 17; const int* q = (const int*)f; KeyLine& k = v[i];
        k.angle = f[0]; k.class_id = q[1]; k.octave = q[2]; k.pt = cv::Point2f(f[3], f[4]); k.response = f[5]; k.size = f[6];
        k.startPointX = f[7]; k.startPointY = f[8]; k.endPointX = f[9]; k.endPointY = f[10];
        k.sPointInOctaveX = f[11]; k.sPointInOctaveY = f[12]; k.ePointInOctaveX = f[13]; k.ePointInOctaveY = f[14]; k.lineLength = f[15]; k.numOfPixels = q[16];
    }
    return v;
}
static Arr keylines_to(const std::vector<KeyLine>& v) {
    std::vector<float> o(v.size() * 17);
    for (size_t i = 0; i < v.size(); ++i) {
        float* f = o.data() + i * 17; int* q = (int*)f; const KeyLine& k = v[i];
        f[0] = k.angle; q[1] = k.class_id; q[2] = k.octave; f[3] = k.pt.x; f[4] = k.pt.y; f[5] = k.response; f[6] = k.size;
        f[7] = k.startPointX; f[8] = k.startPointY; f[9] = k.endPointX; f[10] = k.endPointY;
        f[11] = k.sPointInOctaveX; f[12] = k.sPointInOctaveY; f[13] = k.ePointInOctaveX; f[14] = k.ePointInOctaveY; f[15] = k.lineLength; q[16] = k.numOfPixels;
    }
    return make<float>(2, {(long long)v.size(), 17}, o.data());
}
static Arr keypoints_to(const std::vector<cv::KeyPoint>& v) {       // olf_keypoint rows: x, y, size, angle, response, octave(int)
    std::vector<float> o(v.size() * 6);
    for (size_t i = 0; i < v.size(); ++i) { float* f = o.data() + i * 6; f[0] = v[i].pt.x; f[1] = v[i].pt.y; f[2] = v[i].size; f[3] = v[i].angle; f[4] = v[i].response; ((int*)f)[5] = v[i].octave; }
    return make<float>(2, {(long long)v.size(), 6}, o.data());
}
static std::vector<cv::KeyPoint> keypoints_from(Arr& a) {
    std::vector<cv::KeyPoint> v(a.dims[0]);
    for (size_t i = 0; i < v.size(); ++i) { const float* f = a.as<float>() + i * 6; v[i] = cv::KeyPoint(f[0], f[1], f[2], f[3], f[4], ((const int*)f)[5]); }
    return v;
}
static Arr mat_to(const cv::Mat& m) {
    std::vector<uchar> o((size_t)m.rows * m.cols * m.elemSize());
    for (int r = 0; r < m.rows; ++r) memcpy(o.data() + (size_t)r * m.cols * m.elemSize(), m.ptr(r), (size_t)m.cols * m.elemSize());
    return make<uchar>(m.depth() == CV_32F ? 2 : 0, {m.rows, m.cols}, o.data());
}
static void fill_frame_points(ORB_SLAM2::Frame& F, ORB_SLAM2::ORBextractor& ex, Arr& kps, Arr& desc, Arr& cam /* fx fy cx cy bf w h */) {
    F.mvKeys = keypoints_from(kps); F.mvKeysUn = F.mvKeys; F.N = (int)F.mvKeys.size();
    F.mDescriptors = cv::Mat((int)desc.dims[0], 32, CV_8UC1); for (int r = 0; r < F.mDescriptors.rows; ++r) memcpy(F.mDescriptors.ptr(r), desc.as<uchar>() + (size_t)r * 32, 32);
    const float* c = cam.as<float>();
    F.fx = c[0]; F.fy = c[1]; F.cx = c[2]; F.cy = c[3]; F.invfx = 1.0f / F.fx; F.invfy = 1.0f / F.fy; F.mbf = c[4]; F.mb = F.mbf / F.fx;     // src/Frame.cc:182-196
    F.mnMinX = 0.0f; F.mnMaxX = c[5]; F.mnMinY = 0.0f; F.mnMaxY = c[6];                                                                     // ComputeImageBounds, rectified
    F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(F.mnMaxX - F.mnMinX);
    F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(F.mnMaxY - F.mnMinY);
    F.mvScaleFactors = ex.GetScaleFactors(); F.mvInvScaleFactors = ex.GetInverseScaleFactors(); F.mnScaleLevels = ex.GetLevels();
    F.mvpMapPoints.assign(F.N, static_cast<ORB_SLAM2::MapPoint*>(NULL)); F.mvbOutlier.assign(F.N, false);
    F.AssignFeaturesToGrid();
}

// ---- key frames and map points for Fuse / SearchBySim3 / SearchByProjection(KF|reloc) / SearchForTriangulation / SearchByBoW(KF,KF)
// cam = fx fy cx cy bf w h; pose = Rcw[9] tcw[3] Ow[3] (may be null: identity); the grid is the Frame's (KeyFrame copies F.mGrid, src/KeyFrame.cc:49-54)
static void fill_keyframe(ORB_SLAM2::KeyFrame& K, ORB_SLAM2::ORBextractor& ex, Arr& kps, Arr& desc, Arr* uright, Arr& cam, const float* pose) {
    using namespace ORB_SLAM2;
    Frame F; fill_frame_points(F, ex, kps, desc, cam);
    K.mvKeysUn = F.mvKeysUn; K.N = F.N; K.mDescriptors = F.mDescriptors.clone();
    K.mvuRight.assign(K.N, -1.f); if (uright && uright->count()) K.mvuRight.assign(uright->as<float>(), uright->as<float>() + uright->count());
    K.fx = F.fx; K.fy = F.fy; K.cx = F.cx; K.cy = F.cy; K.mbf = F.mbf;
    K.mnMinX = (int)F.mnMinX; K.mnMinY = (int)F.mnMinY; K.mnMaxX = (int)F.mnMaxX; K.mnMaxY = (int)F.mnMaxY;
    K.mfGridElementWidthInv = F.mfGridElementWidthInv; K.mfGridElementHeightInv = F.mfGridElementHeightInv;
    K.mGrid.resize(K.mnGridCols);
    for (int i = 0; i < K.mnGridCols; ++i) { K.mGrid[i].resize(K.mnGridRows); for (int j = 0; j < K.mnGridRows; ++j) K.mGrid[i][j] = F.mGrid[i][j]; }
    K.mvScaleFactors = ex.GetScaleFactors(); K.mvLevelSigma2 = ex.GetScaleSigmaSquares(); K.mvInvLevelSigma2 = ex.GetInverseScaleSigmaSquares();
    K.mnScaleLevels = ex.GetLevels(); K.mfLogScaleFactor = log(ex.GetScaleFactor());            // src/Frame.cc:149: log(float) under `using namespace std`
    K.mvpMapPoints.assign(K.N, static_cast<MapPoint*>(NULL));
    if (pose) { K.Rcw_ = mat_f32(pose, 3, 3); K.tcw_ = mat_f32(pose + 9, 3, 1); K.Ow_ = mat_f32(pose + 12, 3, 1); }
}
// map points: pos f32[n,3], normal f32[n,3] (or empty), maxd f32[n], mind f32[n], desc u8[n,32], obs i32[n] (or empty)
static void fill_points(std::vector<ORB_SLAM2::MapPoint>& pts, Arr& pos, Arr* normal, Arr& maxd, Arr& mind, Arr& desc, Arr* obs) {
    const int n = (int)maxd.count();
    pts = std::vector<ORB_SLAM2::MapPoint>(n);
    for (int i = 0; i < n; ++i) {
        ORB_SLAM2::MapPoint& m = pts[i];
        m.pos_ = mat_f32(pos.as<float>() + 3 * i, 3, 1);
        if (normal && normal->count()) m.normal_ = mat_f32(normal->as<float>() + 3 * i, 3, 1);
        m.mfMaxDistance = maxd.as<float>()[i]; m.mfMinDistance = mind.as<float>()[i];
        m.desc_ = cv::Mat(1, 32, CV_8UC1); memcpy(m.desc_.ptr(), desc.as<uchar>() + (size_t)32 * i, 32);
        m.nobs_ = obs && obs->count() ? obs->as<int>()[i] : 1;
    }
}
static void fill_featvec(DBoW2::FeatureVector& fv, Arr& node, Arr& begin, Arr& index) {
    for (size_t a = 0; a < node.count(); ++a) {
        std::vector<unsigned int>& v = fv[(DBoW2::NodeId)node.as<int>()[a]];
        for (int p = begin.as<int>()[a]; p < begin.as<int>()[a + 1]; ++p) v.push_back((unsigned)index.as<int>()[p]);
    }
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: refcli <command> <in.bin> <out.bin>\n"); return 2; }
    const std::string cmd = argv[1];
    std::vector<Arr> in = read_arrays(argv[2]), out;
    using namespace ORB_SLAM2;
    if (cmd == "orb") {                       // in: img, params i32[5] (nfeatures, nlevels, iniTh, minTh, _), scale f32[1] -> kps, desc, level sizes
        cv::Mat img = mat_u8(in[0]); const int* p = in[1].as<int>();
        ORBextractor ex(p[0], in[2].as<float>()[0], p[1], p[2], p[3]);
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        ex(img, cv::Mat(), kps, desc);
        out.push_back(keypoints_to(kps)); out.push_back(mat_to(desc));
    } else if (cmd == "line_extract") {       // in: img, iparams i32[3] (nfeatures, refine, n_bins), dparams f64[7] (min_len, scale, sigma, quant, ang, eps, dens)
        cv::Mat img = mat_u8(in[0]); const int* ip = in[1].as<int>(); const double* dp = in[2].as<double>();
        Lineextractor le(ip[0], dp[0], ip[1], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], ip[2]);
        std::vector<KeyLine> kls; cv::Mat desc;
        le(img, cv::Mat(), kls, desc);
        out.push_back(keylines_to(kls)); out.push_back(mat_to(desc));
    } else if (cmd == "lbd") {                // in: img, keylines -> binary descriptors, float descriptors (BinaryDescriptor::compute)
        cv::Mat img = mat_u8(in[0]); std::vector<KeyLine> kls = keylines_from(in[1]);
        cv::Ptr<cv::line_descriptor::BinaryDescriptor> lbd = cv::line_descriptor::BinaryDescriptor::createBinaryDescriptor();
        cv::Mat d, df; lbd->compute(img, kls, d, false); lbd->compute(img, kls, df, true);
        out.push_back(mat_to(d)); out.push_back(mat_to(df));
    } else if (cmd == "line_iterator") {      // in: f64[n,4] -> per line: coords list (getLineCoords, src/LineIterator.cpp + gridStructure.cpp)
        std::vector<int> flat, cnt;
        for (long long i = 0; i < in[0].dims[0]; ++i) {
            const double* l = in[0].as<double>() + i * 4; std::list<std::pair<int, int>> lc;
            getLineCoords(l[0], l[1], l[2], l[3], lc);
            cnt.push_back((int)lc.size()); for (auto& p : lc) { flat.push_back(p.first); flat.push_back(p.second); }
        }
        out.push_back(make<int>(1, {(long long)cnt.size()}, cnt.data())); out.push_back(make<int>(1, {(long long)flat.size() / 2, 2}, flat.data()));
    } else if (cmd == "match_nnr" || cmd == "match") {   // in: desc1, desc2, f32[1] nnr -> matches12, count
        cv::Mat d1 = mat_u8(in[0]), d2 = mat_u8(in[1]); std::vector<int> m12;
        const int n = cmd == "match" ? match(d1, d2, in[2].as<float>()[0], m12) : matchNNR(d1, d2, in[2].as<float>()[0], m12);
        out.push_back(make<int>(1, {(long long)m12.size()}, m12.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "descriptor_distance") {           // in: desc a [n,32], desc b [n,32] -> ORBmatcher::DescriptorDistance, LineMatcher distance
        cv::Mat a = mat_u8(in[0]), b = mat_u8(in[1]); std::vector<int> d1(a.rows), d2(a.rows);
        for (int i = 0; i < a.rows; ++i) { d1[i] = ORBmatcher::DescriptorDistance(a.row(i), b.row(i)); d2[i] = distance(a.row(i), b.row(i)); }
        out.push_back(make<int>(1, {(long long)a.rows}, d1.data())); out.push_back(make<int>(1, {(long long)a.rows}, d2.data()));
    } else if (cmd == "stereo") {
        // in: imgL, imgR, orb params i32[5], scale f32[1], cam f32[7], line iparams i32[3], line dparams f64[7], has_lines i32[1]
        // -> Frame::Frame(stereo+lines) (src/Frame.cc:136-221): kpsL, descL, kpsR, descR, uRight, depth, klsL, ldescL, klsR, ldescR, disp[n,2], le[n,3]
        cv::Mat imL = mat_u8(in[0]), imR = mat_u8(in[1]); const int* p = in[2].as<int>();
        ORBextractor exL(p[0], in[3].as<float>()[0], p[1], p[2], p[3]), exR(p[0], in[3].as<float>()[0], p[1], p[2], p[3]);
        Frame F; F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
        exL(imL, cv::Mat(), F.mvKeys, F.mDescriptors); exR(imR, cv::Mat(), F.mvKeysRight, F.mDescriptorsRight);
        Arr kl = keypoints_to(F.mvKeys), dl = mat_to(F.mDescriptors);
        fill_frame_points(F, exL, kl, dl, in[4]);
        F.ComputeStereoMatches();
        out.push_back(kl); out.push_back(dl); out.push_back(keypoints_to(F.mvKeysRight)); out.push_back(mat_to(F.mDescriptorsRight));
        out.push_back(make<float>(2, {(long long)F.mvuRight.size()}, F.mvuRight.data())); out.push_back(make<float>(2, {(long long)F.mvDepth.size()}, F.mvDepth.data()));
        if (in[7].as<int>()[0]) {
            const int* ip = in[5].as<int>(); const double* dp = in[6].as<double>();
            Lineextractor leL(ip[0], dp[0], ip[1], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], ip[2]), leR(ip[0], dp[0], ip[1], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], ip[2]);
            leL(imL, cv::Mat(), F.mvKeys_Line, F.mDescriptors_Line); leR(imR, cv::Mat(), F.mvKeysRight_Line, F.mDescriptorsRight_Line);
            F.inv_width = FRAME_GRID_COLS / static_cast<double>(imL.cols); F.inv_height = FRAME_GRID_ROWS / static_cast<double>(imR.rows);   // src/Frame.cc:148-149
            F.ComputeStereoMatches_Lines();
            out.push_back(keylines_to(F.mvKeys_Line)); out.push_back(mat_to(F.mDescriptors_Line)); out.push_back(keylines_to(F.mvKeysRight_Line)); out.push_back(mat_to(F.mDescriptorsRight_Line));
            std::vector<float> disp; std::vector<double> le;
            for (auto& d : F.mvDisparity_l) { disp.push_back(d.first); disp.push_back(d.second); }
            for (auto& v : F.mvle_l) { le.push_back(v(0)); le.push_back(v(1)); le.push_back(v(2)); }
            out.push_back(make<float>(2, {(long long)disp.size() / 2, 2}, disp.data())); out.push_back(make<double>(3, {(long long)le.size() / 3, 3}, le.data()));
        }
    } else if (cmd == "sbp_last") {
        // in: cur kps, cur desc, cur uRight, last kps, last has_point u8, last observed u8, last world f32[n,3], last point desc, cam f32[7],
        //     poses f32[24] (Rcw 9, tcw 3, Rlw 9, tlw 3), scale factors f32[L], th_mono_ori f32[3], orb i32[5]/scale for the extractor tables
        // -> ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono, match12) (src/ORBmatcher.cc:1474-1618): cur_point i32[n_cur], nmatches
        Arr &ck = in[0], &cd = in[1], &cu = in[2], &lk = in[3], &has = in[4], &obs = in[5], &wp = in[6], &ld = in[7], &cam = in[8], &pose = in[9];
        const int* p = in[12].as<int>();
        ORBextractor ex(p[0], in[13].as<float>()[0], p[1], p[2], p[3]);
        Frame Cur, Last;
        fill_frame_points(Cur, ex, ck, cd, cam); Cur.mvuRight.assign(cu.as<float>(), cu.as<float>() + cu.count());
        Arr lde = make<uchar>(0, {(long long)lk.dims[0], 32}, ld.as<uchar>());
        fill_frame_points(Last, ex, lk, lde, cam);
        auto Tcw = [](const float* R, const float* t) { cv::Mat T = cv::Mat::eye(4, 4, CV_32F); for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.at<float>(r, c) = R[3 * r + c]; T.at<float>(r, 3) = t[r]; } return T; };
        const float* ps = pose.as<float>();
        Cur.mTcw = Tcw(ps, ps + 9); Last.mTcw = Tcw(ps + 12, ps + 21);
        std::vector<MapPoint> pts(Last.N);
        for (int i = 0; i < Last.N; ++i) {
            if (!has.as<uchar>()[i]) continue;
            pts[i].nobs_ = obs.as<uchar>()[i] ? 1 : 0; pts[i].pos_ = mat_f32(wp.as<float>() + 3 * i, 3, 1);
            pts[i].desc_ = cv::Mat(1, 32, CV_8UC1); memcpy(pts[i].desc_.ptr(), ld.as<uchar>() + (size_t)32 * i, 32);
            Last.mvpMapPoints[i] = &pts[i];
        }
        const float* tmo = in[11].as<float>();
        ORBmatcher matcher(0.9, tmo[2] != 0);
        std::map<int, int> match12;
        // th_mono_ori[3] != 0 selects the overload without match12 (src/ORBmatcher.cc:1330-1472)
        const int n = in[11].count() > 3 && tmo[3] != 0 ? matcher.SearchByProjection(Cur, Last, tmo[0], tmo[1] != 0)
                                                        : matcher.SearchByProjection(Cur, Last, tmo[0], tmo[1] != 0, match12);
        std::vector<int> cur_point(Cur.N, -1);
        for (int j = 0; j < Cur.N; ++j) if (Cur.mvpMapPoints[j]) cur_point[j] = (int)(Cur.mvpMapPoints[j] - pts.data());
        out.push_back(make<int>(1, {(long long)Cur.N}, cur_point.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "sbp_map") {
        // in: cur kps, cur desc, cur uRight, occupied u8 (or empty), cam, proj f32[n,3] (x,y,xr), level i32[n], viewcos f32[n], observed u8[n], point desc, th_ratio f32[2], orb i32[5], scale
        // -> ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:47-131): assigned_cur i32[n_points], nmatches
        Arr &ck = in[0], &cd = in[1], &cu = in[2], &occ = in[3], &cam = in[4], &proj = in[5], &lvl = in[6], &vc = in[7], &obs = in[8], &pd = in[9];
        const int* p = in[11].as<int>();
        ORBextractor ex(p[0], in[12].as<float>()[0], p[1], p[2], p[3]);
        Frame F; fill_frame_points(F, ex, ck, cd, cam); F.mvuRight.assign(cu.as<float>(), cu.as<float>() + cu.count());
        MapPoint occupied_marker; occupied_marker.nobs_ = 1;
        if (occ.count()) for (int j = 0; j < F.N; ++j) if (occ.as<uchar>()[j]) F.mvpMapPoints[j] = &occupied_marker;
        const int np = (int)lvl.count();
        std::vector<MapPoint> pts(np); std::vector<MapPoint*> vp(np);
        for (int i = 0; i < np; ++i) {
            MapPoint& m = pts[i]; m.mbTrackInView = true; m.mTrackProjX = proj.as<float>()[3 * i]; m.mTrackProjY = proj.as<float>()[3 * i + 1]; m.mTrackProjXR = proj.as<float>()[3 * i + 2];
            m.mnTrackScaleLevel = lvl.as<int>()[i]; m.mTrackViewCos = vc.as<float>()[i]; m.nobs_ = obs.as<uchar>()[i] ? 1 : 0;
            m.desc_ = cv::Mat(1, 32, CV_8UC1); memcpy(m.desc_.ptr(), pd.as<uchar>() + (size_t)32 * i, 32);
            vp[i] = &m;
        }
        const float* tr = in[10].as<float>();
        ORBmatcher matcher(tr[1], true);
        const int n = matcher.SearchByProjection(F, vp, tr[0]);
        std::vector<int> assigned(np, -1);
        for (int j = 0; j < F.N; ++j) if (F.mvpMapPoints[j] && F.mvpMapPoints[j] != &occupied_marker) assigned[(int)(F.mvpMapPoints[j] - pts.data())] = j;
        out.push_back(make<int>(1, {(long long)np}, assigned.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "distinctive") {
        // in: desc u8[total,32], group_begin i32[n+1] -> MapPoint::ComputeDistinctiveDescriptors per group: index of the chosen observation
        Arr &d = in[0], &gb = in[1];
        const int ng = (int)gb.count() - 1;
        std::vector<int> best(ng, -1);
        for (int g = 0; g < ng; ++g) {
            const int b = gb.as<int>()[g], N = gb.as<int>()[g + 1] - b;
            std::vector<KeyFrame> kfs(N);                         // ascending addresses: the observation map iterates in this order
            MapPoint mp;
            for (int i = 0; i < N; ++i) { kfs[i].mDescriptors = cv::Mat(1, 32, CV_8UC1); memcpy(kfs[i].mDescriptors.ptr(), d.as<uchar>() + (size_t)(b + i) * 32, 32); mp.mObservations[&kfs[i]] = 0; }
            mp.ComputeDistinctiveDescriptors();
            for (int i = 0; i < N && !mp.mDescriptor.empty(); ++i) if (!memcmp(mp.mDescriptor.ptr(), kfs[i].mDescriptors.ptr(), 32)) { best[g] = i; break; }
        }
        out.push_back(make<int>(1, {(long long)ng}, best.data()));
    } else if (cmd == "fuse" || cmd == "fuse_sim3" || cmd == "sbp_kf") {
        // in: kf kps, desc, uRight, cam f32[7], pose (fuse: Rcw tcw Ow f32[15]; others: Scw f32[16]), pos, normal, maxd, mind, pdesc, obs i32[n],
        //     blocked u8[n_kf] (sbp_kf: vpMatched entries on entry; else empty), th f32[1], orb i32[5], scale f32[1]
        // -> per point: the key-frame feature it met (-1: none), return value
        Arr &ck = in[0], &cd = in[1], &cu = in[2], &cam = in[3], &pose = in[4];
        const int* p = in[13].as<int>();
        ORBextractor ex(p[0], in[14].as<float>()[0], p[1], p[2], p[3]);
        KeyFrame KF; fill_keyframe(KF, ex, ck, cd, &cu, cam, cmd == "fuse" ? pose.as<float>() : nullptr);
        std::vector<MapPoint> pts; fill_points(pts, in[5], &in[6], in[7], in[8], in[9], &in[10]);
        const int np = (int)pts.size();
        std::vector<MapPoint*> vp(np); for (int i = 0; i < np; ++i) vp[i] = &pts[i];
        const float th = in[12].as<float>()[0];
        ORBmatcher matcher(0.8, true);
        std::vector<int> res(np, -1); int n = 0;
        if (cmd == "fuse") {
            n = matcher.Fuse(&KF, vp, th);
            for (int i = 0; i < np; ++i) res[i] = pts[i].fused_;
        } else if (cmd == "fuse_sim3") {
            std::vector<MapPoint*> repl(np, static_cast<MapPoint*>(NULL));
            n = matcher.Fuse(&KF, mat_f32(pose.as<float>(), 4, 4), vp, th, repl);
            for (int i = 0; i < np; ++i) res[i] = repl[i] ? repl[i]->fused_ : pts[i].fused_;
        } else {
            MapPoint marker;
            std::vector<MapPoint*> matched(KF.N, static_cast<MapPoint*>(NULL));
            for (int j = 0; j < KF.N && in[11].count(); ++j) if (in[11].as<uchar>()[j]) matched[j] = &marker;
            n = matcher.SearchByProjection(&KF, mat_f32(pose.as<float>(), 4, 4), vp, matched, (int)th);
            for (int j = 0; j < KF.N; ++j) if (matched[j] && matched[j] != &marker) res[(int)(matched[j] - pts.data())] = j;
        }
        out.push_back(make<int>(1, {(long long)np}, res.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "sim3") {
        // in: kps1, desc1, pose1 f32[15], has1 u8[N1], pos1, maxd1, mind1, pdesc1, kps2, desc2, pose2 f32[15], has2 u8[N2], pos2, maxd2, mind2, pdesc2,
        //     cam f32[7], sim f32[13] (s12, R12[9], t12[3]), th f32[1], orb i32[5], scale f32[1] -> matches12 i32[N1] (index into KF2), nFound
        const int* p = in[19].as<int>();
        ORBextractor ex(p[0], in[20].as<float>()[0], p[1], p[2], p[3]);
        KeyFrame K1, K2; fill_keyframe(K1, ex, in[0], in[1], nullptr, in[16], in[2].as<float>()); fill_keyframe(K2, ex, in[8], in[9], nullptr, in[16], in[10].as<float>());
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
        out.push_back(make<int>(1, {(long long)K1.N}, res.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "reloc") {
        // in: cur kps, desc, cam f32[7], Tcw (R, t) f32[12], occupied u8[n_cur], kf kps, has u8[n_kf], pos, maxd, mind, pdesc, th_dist f32[2], orb, scale
        // -> SearchByProjection(CurrentFrame, pKF, sAlreadyFound = {}, th, ORBdist) with mbCheckOrientation = false: cur_point i32[n_cur], n
        const int* p = in[12].as<int>();
        ORBextractor ex(p[0], in[13].as<float>()[0], p[1], p[2], p[3]);
        Frame Cur; fill_frame_points(Cur, ex, in[0], in[1], in[2]);
        Cur.mfScaleFactor = ex.GetScaleFactor(); Cur.mfLogScaleFactor = log(Cur.mfScaleFactor);
        const float* ps = in[3].as<float>();
        Cur.mTcw = cv::Mat::eye(4, 4, CV_32F); for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Cur.mTcw.at<float>(r, c) = ps[3 * r + c]; Cur.mTcw.at<float>(r, 3) = ps[9 + r]; }
        MapPoint marker;
        for (int j = 0; j < Cur.N; ++j) if (in[4].as<uchar>()[j]) Cur.mvpMapPoints[j] = &marker;
        Arr kdesc = make<uchar>(0, {(long long)in[5].dims[0], 32}, in[10].as<uchar>());
        KeyFrame KF; fill_keyframe(KF, ex, in[5], kdesc, nullptr, in[2], nullptr);
        std::vector<MapPoint> pts; fill_points(pts, in[7], nullptr, in[8], in[9], in[10], nullptr);
        for (int i = 0; i < KF.N; ++i) if (in[6].as<uchar>()[i]) KF.mvpMapPoints[i] = &pts[i];
        ORBmatcher matcher(0.9, false);
        std::set<MapPoint*> found;
        const int n = matcher.SearchByProjection(Cur, &KF, found, in[11].as<float>()[0], (int)in[11].as<float>()[1]);
        std::vector<int> res(Cur.N, -1);
        for (int j = 0; j < Cur.N; ++j) if (Cur.mvpMapPoints[j] && Cur.mvpMapPoints[j] != &marker) res[j] = (int)(Cur.mvpMapPoints[j] - pts.data());
        out.push_back(make<int>(1, {(long long)Cur.N}, res.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "init") {
        // in: kps1, desc1, kps2, desc2, cam f32[7], prev f32[n1,2], par f32[3] (window, nn_ratio, check_orientation), orb i32[5], scale
        // -> ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522): matches12 i32[n1], n, prev f32[n1,2]
        const int* p = in[7].as<int>();
        ORBextractor ex(p[0], in[8].as<float>()[0], p[1], p[2], p[3]);
        Frame F1, F2; fill_frame_points(F1, ex, in[0], in[1], in[4]); fill_frame_points(F2, ex, in[2], in[3], in[4]);
        std::vector<cv::Point2f> prev(F1.N);
        for (int i = 0; i < F1.N; ++i) prev[i] = cv::Point2f(in[5].as<float>()[2 * i], in[5].as<float>()[2 * i + 1]);
        const float* par = in[6].as<float>();
        ORBmatcher matcher(par[1], par[2] != 0);
        std::vector<int> m12;
        const int n = matcher.SearchForInitialization(F1, F2, prev, m12, (int)par[0]);
        std::vector<float> pv(2 * (size_t)F1.N);
        for (int i = 0; i < F1.N; ++i) { pv[2 * i] = prev[i].x; pv[2 * i + 1] = prev[i].y; }
        out.push_back(make<int>(1, {(long long)F1.N}, m12.data())); out.push_back(make<int>(1, {1}, &n)); out.push_back(make<float>(2, {(long long)F1.N, 2}, pv.data()));
    } else if (cmd == "triangulation" || cmd == "bow_kf") {
        // in: kps1, desc1, skip1 u8, uRight1, node1, begin1, index1, kps2, desc2, skip2, uRight2, node2, begin2, index2, cam f32[7],
        //     geo f32[24] (Cw[3], R2w[9], t2w[3], F12[9]), flags f32[3] (only_stereo, check_orientation, nn_ratio), orb, scale
        // -> matches12 i32[n1], return value.  skip = "has a map point" for triangulation, = "has NO good map point" for bow_kf
        const int* p = in[17].as<int>();
        ORBextractor ex(p[0], in[18].as<float>()[0], p[1], p[2], p[3]);
        const float* geo = in[15].as<float>(); const float* fl = in[16].as<float>();
        float pose1[15] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, geo[0], geo[1], geo[2]}, pose2[15];
        memcpy(pose2, geo + 3, 12 * sizeof(float)); pose2[12] = pose2[13] = pose2[14] = 0;
        KeyFrame K1, K2; fill_keyframe(K1, ex, in[0], in[1], &in[3], in[14], pose1); fill_keyframe(K2, ex, in[7], in[8], &in[10], in[14], pose2);
        fill_featvec(K1.mFeatVec, in[4], in[5], in[6]); fill_featvec(K2.mFeatVec, in[11], in[12], in[13]);
        MapPoint marker;                                   // a good map point
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
        out.push_back(make<int>(1, {(long long)K1.N}, res.data())); out.push_back(make<int>(1, {1}, &n));
    } else if (cmd == "bow_transform") {
        // in: vocabulary in the ORBvoc.txt text format u8[len], desc u8[n,32], levelsup i32[1]
        // -> DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::loadFromTextFile + transform(features, BowVector, FeatureVector, levelsup)
        //    (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1131-1194, 1330-1407; the typedef of include/ORBVocabulary.h:31, the call of src/Frame.cc:548-556):
        //    bow word i32[nw], bow value f64[nw], fv node i32[nn], fv begin i32[nn+1], fv index i32[m]
        const std::string path = std::string(argv[3]) + ".voc";
        { FILE* f = fopen(path.c_str(), "wb"); fwrite(in[0].as<uchar>(), 1, in[0].count(), f); fclose(f); }
        DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> voc;
        if (!voc.loadFromTextFile(path)) { fprintf(stderr, "refcli: loadFromTextFile failed\n"); return 3; }
        remove(path.c_str());
        std::vector<cv::Mat> feats(in[1].dims[0]);
        for (size_t j = 0; j < feats.size(); ++j) { feats[j] = cv::Mat(1, 32, CV_8UC1); memcpy(feats[j].ptr(), in[1].as<uchar>() + 32 * j, 32); }
        DBoW2::BowVector bv; DBoW2::FeatureVector fv;
        voc.transform(feats, bv, fv, in[2].as<int>()[0]);
        std::vector<int> bw, fn, fb(1, 0), fi; std::vector<double> bval;
        for (auto& e : bv) { bw.push_back((int)e.first); bval.push_back(e.second); }
        for (auto& e : fv) { fn.push_back((int)e.first); for (unsigned x : e.second) fi.push_back((int)x); fb.push_back((int)fi.size()); }
        out.push_back(make<int>(1, {(long long)bw.size()}, bw.data())); out.push_back(make<double>(3, {(long long)bval.size()}, bval.data()));
        out.push_back(make<int>(1, {(long long)fn.size()}, fn.data())); out.push_back(make<int>(1, {(long long)fb.size()}, fb.data()));
        out.push_back(make<int>(1, {(long long)fi.size()}, fi.data()));
    } else if (cmd == "bow_kff") {
        // in: kf kps, desc, has u8, node, begin, index, frame kps, desc, node, begin, index, flags f32[2] (nn_ratio, check_orientation), cam f32[7], orb, scale
        // -> SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (src/ORBmatcher.cc:161-290): key-frame index per frame feature i32[nF], return value
        const int* p = in[13].as<int>();
        ORBextractor ex(p[0], in[14].as<float>()[0], p[1], p[2], p[3]);
        const float* fl = in[11].as<float>();
        KeyFrame K; fill_keyframe(K, ex, in[0], in[1], nullptr, in[12], nullptr);
        Frame F; fill_frame_points(F, ex, in[6], in[7], in[12]);
        fill_featvec(K.mFeatVec, in[3], in[4], in[5]); fill_featvec(F.mFeatVec, in[8], in[9], in[10]);
        std::vector<MapPoint> pts(K.N);
        for (int i = 0; i < K.N; ++i) if (in[2].as<uchar>()[i]) K.mvpMapPoints[i] = &pts[i];
        ORBmatcher matcher(fl[0], fl[1] != 0);
        std::vector<MapPoint*> mf;
        int n = matcher.SearchByBoW(&K, F, mf);
        std::vector<int> res(F.N, -1);
        for (int j = 0; j < F.N; ++j) if (mf[j]) res[j] = (int)(mf[j] - pts.data());
        out.push_back(make<int>(1, {(long long)F.N}, res.data())); out.push_back(make<int>(1, {1}, &n));
    } else { fprintf(stderr, "refcli: unknown command %s\n", cmd.c_str()); return 2; }
    write_arrays(argv[3], out);
    return 0;
}
