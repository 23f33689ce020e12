// Exercises the reference-shaped C++ class surfaces (orb_line_slam_b200/shim) end to end on raw 8-bit image files:
//   test_shim <w> <h> <dir>    dir holds l0.raw r0.raw l1.raw r1.raw (two stereo frames) and the tracking inputs written by
//                              tests/test_gpu_shim.py (pose.f32, has.u8, world.f32, map.f32, mapdesc.u8)
// Drives ORBextractor / Lineextractor / Frame::ComputeStereoMatches(_Lines) / matchGrid / ORBmatcher::SearchByProjection (both
// kinds) / match(MapLine*, Frame&) the way Frame.cc and Tracking.cc do and prints counts + FNV hashes that the pytest compares
// with the C-ABI path (which the other GPU tests compare with the oracle).
#include "../../orb_line_slam_b200/shim/ORBextractor.h"
#include "../../orb_line_slam_b200/shim/LineExtractor.h"
#include "../../orb_line_slam_b200/shim/LineMatcher.h"
#include "../../orb_line_slam_b200/shim/olf_ref_classes.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
using namespace ORB_SLAM2;
static unsigned long long fnv(const void* p, size_t n, unsigned long long h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
template <typename T> static std::vector<T> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T));
    if (n && fread(v.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    return v;
}
static cv::Mat load(const std::string& path, int w, int h) {
    std::vector<unsigned char> d = slurp<unsigned char>(path);
    cv::Mat m(h, w, CV_8UC1);
    memcpy(m.data, d.data(), (size_t)w * h);
    return m;
}
// Frame::Frame(stereo+lines) as src/Frame.cc:136-221 does it (serially here)
static void make_frame(Frame& F, ORBextractor& eL, ORBextractor& eR, Lineextractor& lL, Lineextractor& lR, const cv::Mat& L, const cv::Mat& R, const float* pose) {
    cv::Mat mask;
    F.mpORBextractorLeft = &eL; F.mpORBextractorRight = &eR;
    eL(L, mask, F.mvKeys, F.mDescriptors); eR(R, mask, F.mvKeysRight, F.mDescriptorsRight);
    lL(L, mask, F.mvKeys_Line, F.mDescriptors_Line); lR(R, mask, F.mvKeysRight_Line, F.mDescriptorsRight_Line);
    F.N = (int)F.mvKeys.size(); F.N_l = (int)F.mvKeys_Line.size();
    F.mvKeysUn = F.mvKeys;                                        // rectified input: UndistortKeyPoints is the identity
    F.fx = 435.2046959714599f; F.fy = 435.2046959714599f; F.cx = 367.4517211914062f * 640 / 752; F.cy = 252.2008514404297f; F.mbf = 47.90639384423901f;
    F.mb = F.mbf / F.fx;
    F.mnMinX = 0; F.mnMaxX = (float)L.cols; F.mnMinY = 0; F.mnMaxY = (float)L.rows;
    F.mnScaleLevels = eL.GetLevels(); F.mvScaleFactors = eL.GetScaleFactors(); F.mvInvScaleFactors = eL.GetInverseScaleFactors();
    F.inv_width = FRAME_GRID_COLS / static_cast<double>(L.cols); F.inv_height = FRAME_GRID_ROWS / static_cast<double>(R.rows);
    F.mvpMapPoints.assign(F.N, static_cast<MapPoint*>(NULL)); F.mvbOutlier.assign(F.N, false);
    F.ComputeStereoMatches();
    F.ComputeStereoMatches_Lines();
    F.mTcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) F.mTcw.at<float>(r, c) = r == c ? 1.f : 0.f;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) F.mTcw.at<float>(r, c) = pose[3 * r + c]; F.mTcw.at<float>(r, 3) = pose[9 + r]; }
}
int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: test_shim w h dir\n"); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]); const std::string dir = argv[3];
    try {
        ORBextractor eL(1000, 1.2f, 8, 20, 7), eR(1000, 1.2f, 8, 20, 7);
        Lineextractor lL(200, 0.025, 0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024), lR(200, 0.025, 0, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024);
        const std::vector<float> pose = slurp<float>(dir + "/pose.f32");                 // cur: Rcw 9, tcw 3; last: Rlw 9, tlw 3
        Frame Last, Cur;
        make_frame(Last, eL, eR, lL, lR, load(dir + "/l0.raw", w, h), load(dir + "/r0.raw", w, h), pose.data() + 12);
        // ---- frame 0: extraction + stereo association
        unsigned long long hd = 1469598103934665603ull, hl = hd;
        for (int i = 0; i < Last.mDescriptors.rows; ++i) hd = fnv(Last.mDescriptors.ptr(i), 32, hd);
        for (int i = 0; i < Last.mDescriptors_Line.rows; ++i) hl = fnv(Last.mDescriptors_Line.ptr(i), 32, hl);
        std::vector<float> disp;
        for (auto& d : Last.mvDisparity_l) { disp.push_back(d.first); disp.push_back(d.second); }
        eL.SyncImagePyramid();
        printf("nL %zu nR %zu mL %zu mR %zu desc %llu ldesc %llu uright %llu disp %llu pyr7 %dx%d levels %d\n", Last.mvKeys.size(), Last.mvKeysRight.size(),
               Last.mvKeys_Line.size(), Last.mvKeysRight_Line.size(), hd, hl, fnv(Last.mvuRight.data(), Last.mvuRight.size() * 4), fnv(disp.data(), disp.size() * 4),
               eL.mvImagePyramid[7].cols, eL.mvImagePyramid[7].rows, eL.GetLevels());
        // ---- matchGrid(lines) through GridStructure, the way Frame::ComputeStereoMatches_Lines calls it (src/Frame.cc:896-927)
        {
            std::vector<line_2d> coords;
            for (const auto& kl : Last.mvKeys_Line)
                coords.push_back(std::make_pair(std::make_pair(kl.startPointX * Last.inv_width, kl.startPointY * Last.inv_height),
                                                std::make_pair(kl.endPointX * Last.inv_width, kl.endPointY * Last.inv_height)));
            std::list<std::pair<int, int>> line_coords;
            GridStructure grid(FRAME_GRID_ROWS, FRAME_GRID_COLS);
            std::vector<std::pair<double, double>> directions(Last.mvKeysRight_Line.size());
            for (unsigned int idx = 0; idx < Last.mvKeysRight_Line.size(); ++idx) {
                const auto& kl = Last.mvKeysRight_Line[idx];
                std::pair<double, double>& v = directions[idx];
                v = std::make_pair((kl.endPointX - kl.startPointX) * Last.inv_width, (kl.endPointY - kl.startPointY) * Last.inv_height);
                normalize(v);
                getLineCoords(kl.startPointX * Last.inv_width, kl.startPointY * Last.inv_height, kl.endPointX * Last.inv_width, kl.endPointY * Last.inv_height, line_coords);
                for (const std::pair<int, int>& p : line_coords) grid.at(p.first, p.second).push_back(idx);
            }
            GridWindow win; win.width = std::make_pair(OlfConfig::line_match.matching_s_ws, 0); win.height = std::make_pair(0, 0);
            std::vector<int> m12;
            const int nm = matchGrid(coords, Last.mDescriptors_Line, grid, Last.mDescriptorsRight_Line, directions, win, m12);
            printf("matchgrid %d %llu\n", nm, fnv(m12.data(), m12.size() * 4));
        }
        // ---- frame 1 + the per-frame matchers of Tracking (src/Tracking.cc:1296-1308)
        make_frame(Cur, eL, eR, lL, lR, load(dir + "/l1.raw", w, h), load(dir + "/r1.raw", w, h), pose.data());
        const std::vector<unsigned char> has = slurp<unsigned char>(dir + "/has.u8");
        const std::vector<float> world = slurp<float>(dir + "/world.f32");
        std::vector<MapPoint> pts(Last.N);
        for (int i = 0; i < Last.N; ++i) {
            if (!has[i]) continue;
            pts[i].pos = cv::Mat(3, 1, CV_32F); for (int r = 0; r < 3; ++r) pts[i].pos.at<float>(r, 0) = world[3 * i + r];
            pts[i].desc = Last.mDescriptors.row(i).clone();
            Last.mvpMapPoints[i] = &pts[i];
        }
        ORBmatcher matcher(0.9, true);
        std::map<int, int> match12;
        const int nsbp = matcher.SearchByProjection(Cur, Last, 7.0f, false, match12);
        std::vector<int> cur_point(Cur.N, -1);
        for (int j = 0; j < Cur.N; ++j) if (Cur.mvpMapPoints[j]) cur_point[j] = (int)(Cur.mvpMapPoints[j] - pts.data());
        printf("sbplast %d %llu %zu\n", nsbp, fnv(cur_point.data(), cur_point.size() * 4), match12.size());
        // SearchByProjection(F, vpMapPoints, th): map points carry what Frame::isInFrustum leaves in them
        const std::vector<float> mp = slurp<float>(dir + "/map.f32");                     // rows: projx, projy, projxr, level, viewcos
        const std::vector<unsigned char> mdesc = slurp<unsigned char>(dir + "/mapdesc.u8");
        const int nmap = (int)mp.size() / 5;
        std::vector<MapPoint> mpts(nmap); std::vector<MapPoint*> vp(nmap);
        for (int i = 0; i < nmap; ++i) {
            MapPoint& m = mpts[i]; m.mbTrackInView = true; m.mTrackProjX = mp[5 * i]; m.mTrackProjY = mp[5 * i + 1]; m.mTrackProjXR = mp[5 * i + 2];
            m.mnTrackScaleLevel = (int)mp[5 * i + 3]; m.mTrackViewCos = mp[5 * i + 4];
            m.desc = cv::Mat(1, 32, CV_8UC1); memcpy(m.desc.ptr(0), mdesc.data() + (size_t)32 * i, 32);
            vp[i] = &m;
        }
        Cur.mvpMapPoints.assign(Cur.N, static_cast<MapPoint*>(NULL));
        ORBmatcher matcher2(0.8, true);
        const int nmapm = matcher2.SearchByProjection(Cur, vp, 1.0f);
        std::vector<int> assigned(nmap, -1);
        for (int j = 0; j < Cur.N; ++j) if (Cur.mvpMapPoints[j]) assigned[(int)(Cur.mvpMapPoints[j] - mpts.data())] = j;
        printf("sbpmap %d %llu\n", nmapm, fnv(assigned.data(), assigned.size() * 4));
        // lines: match(desc, desc) and match(vector<MapLine*>, Frame&)
        std::vector<int> m12;
        const int nm = match(Last.mDescriptors_Line, Cur.mDescriptors_Line, 0.9f, m12);
        printf("match %d %llu\n", nm, fnv(m12.data(), m12.size() * 4));
        std::vector<MapLine> mls(Last.N_l); std::vector<MapLine*> vml(Last.N_l);
        for (int i = 0; i < Last.N_l; ++i) { mls[i].desc = Last.mDescriptors_Line.row(i).clone(); vml[i] = &mls[i]; }
        const int nml = match(vml, Cur, 0.9f, m12);
        printf("matchmaplines %d %llu\n", nml, fnv(m12.data(), m12.size() * 4));
        printf("distance %d\n", ORBmatcher::DescriptorDistance(Last.mDescriptors.row(0), Last.mDescriptors.row(1)) - distance(Last.mDescriptors.row(0), Last.mDescriptors.row(1)));
    } catch (const std::exception& e) { printf("EXCEPTION %s\n", e.what()); return 1; }
    return 0;
}
