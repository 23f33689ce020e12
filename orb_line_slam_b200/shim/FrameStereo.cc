// Frame::ComputeStereoMatches (src/Frame.cc:702-876) and Frame::ComputeStereoMatches_Lines (:878-1000) on the B200.
// Replaces those two member function bodies of src/Frame.cc: the stereo point matcher reads the
// pyramids where the extractors left them -- on the device -- instead of ORBextractor::mvImagePyramid, and the line matcher
// runs grid construction, matchGrid and the geometric filters as one call.
#include "olf_ref_classes.h"
#include "ORBextractor.h"
#include "LineMatcher.h"
#include <cstring>
#include <stdexcept>
#include <string>

namespace ORB_SLAM2 {
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
}  // namespace

void Frame::ComputeStereoMatches() {
    mvuRight = std::vector<float>(N, -1.0f);                     // :704-705
    mvDepth = std::vector<float>(N, -1.0f);
    if (N == 0) return;
    const std::vector<olf_keypoint> kl = pack_kps(mvKeys), kr = pack_kps(mvKeysRight);
    const std::vector<uint8_t> dl = pack_rows(mDescriptors), dr = pack_rows(mDescriptorsRight);
    if (olf_stereo_points(mpORBextractorLeft->handle(), mpORBextractorRight->handle(), kl.data(), dl.data(), N, kr.data(), dr.data(), (int)kr.size(),
                          mbf, fx, mvuRight.data(), mvDepth.data()) != OLF_OK)
        throw std::runtime_error(std::string("[ComputeStereoMatches] ") + olf_last_error());
}

void Frame::ComputeStereoMatches_Lines(bool) {
    const int n1 = (int)mvKeys_Line.size(), n2 = (int)mvKeysRight_Line.size();
    mvDisparity_l.clear(); mvle_l.clear();
    mvDisparity_l.resize(n1, std::pair<float, float>(-1, -1));  // doNotDropMonoLines (:880-888)
    mvle_l.resize(n1);
    for (int i = 0; i < n1; ++i) olf_set_le(*this, i, 0, 0, 0);
    if (n1 == 0 || n2 == 0) return;                              // :896-897
    std::vector<int> m(n1); std::vector<float> d((size_t)n1 * 2); std::vector<double> le((size_t)n1 * 3);
    const std::vector<uint8_t> dl = pack_rows(mDescriptors_Line), dr = pack_rows(mDescriptorsRight_Line);
    // inv_width = FRAME_GRID_COLS / imLeft.cols, inv_height = FRAME_GRID_ROWS / imRight.rows (:148-149): the image size follows
    const int w = (int)std::lrint(FRAME_GRID_COLS / inv_width), h = (int)std::lrint(FRAME_GRID_ROWS / inv_height);
    if (olf_stereo_lines((const olf_keyline*)mvKeys_Line.data(), dl.data(), n1, (const olf_keyline*)mvKeysRight_Line.data(), dr.data(), n2, w, h,
                         &OlfConfig::line_match, m.data(), d.data(), le.data(), OlfConfig::device) != OLF_OK)
        throw std::runtime_error(std::string("[matchGrid] ") + olf_last_error());
    for (int i = 0; i < n1; ++i) { mvDisparity_l[i] = std::make_pair(d[2 * i], d[2 * i + 1]); olf_set_le(*this, i, le[3 * i], le[3 * i + 1], le[3 * i + 2]); }
}
}  // namespace ORB_SLAM2
