#include "LineMatcher.h"
#include "ORBextractor.h"
#include "olf_ref_classes.h"
#include <climits>
#include <stdexcept>
#include <string>
#include <algorithm>

namespace ORB_SLAM2 {
olf_line_match_params OlfConfig::line_match = {1, 0.9, 0.75, 10, 1.0, 0.1, 0.75, 0.7};
int OlfConfig::device = 0;
double OlfConfig::min_ratio_12_p = 0.75;         // Config::minRatio12P() (src/Config.cpp:56)

static std::vector<uint8_t> pack(const cv::Mat& d) {
    std::vector<uint8_t> v((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; ++i) memcpy(v.data() + (size_t)i * 32, d.ptr(i), 32);
    return v;
}
int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
    matches_12.assign(desc1.rows, -1);
    int n = 0;
    const std::vector<uint8_t> a = pack(desc1), b = pack(desc2);
    if (olf_match_nnr(a.data(), desc1.rows, b.data(), desc2.rows, nnr, matches_12.data(), &n, OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[matchNNR] ") + olf_last_error());
    return n;
}
int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
    matches_12.assign(desc1.rows, -1);
    int n = 0;
    const std::vector<uint8_t> a = pack(desc1), b = pack(desc2);
    if (olf_match_lines(a.data(), desc1.rows, b.data(), desc2.rows, nnr, OlfConfig::line_match.best_lr_matches, matches_12.data(), &n, OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[match] ") + olf_last_error());
    return n;
}
int match(const std::vector<MapLine*>& mvpLocalMapLines, Frame& CurrentFrame, float nnr, std::vector<int>& matches_12) {
    std::vector<uint8_t> a((size_t)mvpLocalMapLines.size() * 32);
    for (size_t i = 0; i < mvpLocalMapLines.size(); ++i) { const cv::Mat d = mvpLocalMapLines[i]->GetDescriptor(); memcpy(a.data() + i * 32, d.ptr(0), 32); }
    const std::vector<uint8_t> b = pack(CurrentFrame.mDescriptors_Line);
    matches_12.assign(mvpLocalMapLines.size(), -1);
    int n = 0;
    if (olf_match_nnr(a.data(), (int)mvpLocalMapLines.size(), b.data(), CurrentFrame.mDescriptors_Line.rows, nnr, matches_12.data(), &n, OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[matchNNR] ") + olf_last_error());
    return n;
}
int matchGrid(const std::vector<line_2d>& lines1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2,
              const std::vector<std::pair<double, double>>& directions2, const GridWindow& w, std::vector<int>& matches_12) {
    if ((int)lines1.size() != desc1.rows) throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");     // :224-225
    matches_12.resize(desc1.rows, -1);
    const int n1 = desc1.rows, n2 = desc2.rows;
    std::vector<int> l1((size_t)n1 * 4), begin(1, 0), items;
    for (int i = 0; i < n1; ++i) { l1[4 * i] = lines1[i].first.first; l1[4 * i + 1] = lines1[i].first.second; l1[4 * i + 2] = lines1[i].second.first; l1[4 * i + 3] = lines1[i].second.second; }
    GridStructure& g = const_cast<GridStructure&>(grid);
    for (int x = 0; x < grid.cols; ++x)
        for (int y = 0; y < grid.rows; ++y) { const std::list<int>& c = g.at(x, y); items.insert(items.end(), c.begin(), c.end()); begin.push_back((int)items.size()); }
    std::vector<double> dir((size_t)n2 * 2);
    for (int i = 0; i < n2 && i < (int)directions2.size(); ++i) { dir[2 * i] = directions2[i].first; dir[2 * i + 1] = directions2[i].second; }
    const olf_grid_csr csr = {grid.rows, grid.cols, begin.data(), items.data()};
    const int win[4] = {w.width.first, w.width.second, w.height.first, w.height.second};
    const std::vector<uint8_t> a = pack(desc1), b = pack(desc2);
    int n = 0;
    if (olf_match_grid_lines(l1.data(), a.data(), n1, &csr, b.data(), n2, dir.data(), win, &OlfConfig::line_match, matches_12.data(), &n, OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[matchGrid] ") + olf_last_error());
    return n;
}
int matchGrid(const std::vector<point_2d>& points1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2, const GridWindow& w, std::vector<int>& matches_12) {
    if ((int)points1.size() != desc1.rows) throw std::runtime_error("[matchGrid] Each point needs a corresponding descriptor!");
    // host loop (the reference never calls this overload on the per-frame path); candidates in ascending index
    const bool lr = OlfConfig::line_match.best_lr_matches != 0;
    matches_12.resize(desc1.rows, -1);
    std::vector<int> m21(lr ? desc2.rows : 0, -1), dist(lr ? desc2.rows : 0, INT_MAX);
    int matches = 0;
    for (int i1 = 0; i1 < desc1.rows; ++i1) {
        std::unordered_set<int> cs; grid.get(points1[i1].first, points1[i1].second, w, cs);
        std::vector<int> cand(cs.begin(), cs.end()); std::sort(cand.begin(), cand.end());
        int bd = INT_MAX, bd2 = INT_MAX, bi = -1;
        for (int i2 : cand) {
            if (i2 < 0 || i2 >= desc2.rows) continue;
            const int d = distance(desc1.row(i1), desc2.row(i2));
            if (lr) { if (d < dist[i2]) { dist[i2] = d; m21[i2] = i1; } else continue; }
            if (d < bd) { bd2 = bd; bd = d; bi = i2; } else if (d < bd2) bd2 = d;
        }
        if (bi >= 0 && bd < bd2 * OlfConfig::min_ratio_12_p) { matches_12[i1] = bi; matches++; }
    }
    if (lr) for (int i1 = 0; i1 < desc1.rows; ++i1) { int& i2 = matches_12[i1]; if (i2 >= 0 && m21[i2] != i1) { i2 = -1; matches--; } }
    return matches;
}
int distance(const cv::Mat& a, const cv::Mat& b) {               // src/LineMatcher.cpp:134-150 (host: single pair)
    const uint32_t* pa = a.ptr<uint32_t>(); const uint32_t* pb = b.ptr<uint32_t>();
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}
static std::vector<olf_keypoint> pack_kps(const std::vector<cv::KeyPoint>& k) {
    std::vector<olf_keypoint> v(k.size());
    for (size_t i = 0; i < k.size(); ++i) v[i] = {k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
    return v;
}
void ComputeStereoMatches(ORBextractor* left, ORBextractor* right, const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc,
                          const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight, float bf, float fx,
                          std::vector<float>& uRight, std::vector<float>& depth) {
    const int N = (int)keys.size();
    uRight.assign(N, -1.0f); depth.assign(N, -1.0f);
    if (N == 0) return;
    const std::vector<olf_keypoint> kl = pack_kps(keys), kr = pack_kps(keysRight);
    const std::vector<uint8_t> dl = pack(desc), dr = pack(descRight);
    if (olf_stereo_points(left->handle(), right->handle(), kl.data(), dl.data(), N, kr.data(), dr.data(), (int)kr.size(), bf, fx, uRight.data(), depth.data()) != OLF_OK)
        throw std::runtime_error(std::string("[ComputeStereoMatches] ") + olf_last_error());
}
void ComputeStereoMatches_Lines(const std::vector<cv::line_descriptor::KeyLine>& keys, const cv::Mat& desc,
                                const std::vector<cv::line_descriptor::KeyLine>& keysRight, const cv::Mat& descRight,
                                int img_width, int img_height, std::vector<std::pair<float, float>>& disparity, std::vector<double>& le) {
    const int n1 = (int)keys.size(), n2 = (int)keysRight.size();
    disparity.assign(n1, std::make_pair(-1.f, -1.f)); le.assign((size_t)n1 * 3, 0.0);
    if (n1 == 0 || n2 == 0) return;
    std::vector<int> m(n1); std::vector<float> d((size_t)n1 * 2);
    const std::vector<uint8_t> dl = pack(desc), dr = pack(descRight);
    if (olf_stereo_lines((const olf_keyline*)keys.data(), dl.data(), n1, (const olf_keyline*)keysRight.data(), dr.data(), n2, img_width, img_height,
                         &OlfConfig::line_match, m.data(), d.data(), le.data(), OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[matchGrid] ") + olf_last_error());
    for (int i = 0; i < n1; ++i) disparity[i] = std::make_pair(d[2 * i], d[2 * i + 1]);
}
}  // namespace ORB_SLAM2
