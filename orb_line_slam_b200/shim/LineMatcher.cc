#include "LineMatcher.h"
#include "ORBextractor.h"
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {
olf_line_match_params OlfConfig::line_match = {1, 0.9, 0.75, 10, 1.0, 0.1, 0.75, 0.7};
int OlfConfig::device = 0;

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
