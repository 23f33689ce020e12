// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load anything under oracle/.
//
// CPU restatement of the un-vendored OpenCV primitives used by the reference hot path
// (OpenCV is a find_package dependency of the reference, CMakeLists.txt:39-46, absent from
// /root/reference).  Semantics are those of cv2 4.13.0 as probed in SURVEY.md Appendix A and
// re-pinned by tests/test_oracle_vs_cv2.py.  Call sites in the reference:
//   cv::resize INTER_LINEAR      src/ORBextractor.cc:1122
//   cv::GaussianBlur (8U)        src/ORBextractor.cc:1088, binary_descriptor_custom.cpp:358, inside cv LSD
//   cv::FAST                     src/ORBextractor.cc:811,816
//   cv::fastAtan2                src/ORBextractor.cc:105, inside cv LSD
//   cv::Sobel                    binary_descriptor_custom.cpp:395-396
//   cv::resize INTER_LINEAR_EXACT inside cv::LineSegmentDetector (LSDDetector_custom.cpp:246-262)
//
// All float arithmetic is single IEEE operations in source order (build with -ffp-contract=off).
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <cstring>

namespace orc {

struct Image8 {            // owning, contiguous 8-bit image
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    Image8() {}
    Image8(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
    uint8_t* row(int y) { return d.data() + (size_t)y * w; }
    const uint8_t* row(int y) const { return d.data() + (size_t)y * w; }
    uint8_t at(int x, int y) const { return d[(size_t)y * w + x]; }
};

static inline int cv_round(float v) { return (int)lrintf(v); }     // cvRound: round-half-even
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(float v) { return (int)floorf(v); }
static inline int reflect101(int i, int n) {                        // BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

// ---- cv::fastAtan2 (degrees), SURVEY A.4 -----------------------------------------------------
static inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- cv::resize INTER_LINEAR, 8UC1, SURVEY A.2 ------------------------------------------------
struct LinCoef { int s; short c0, c1; };
static inline void linear_coeffs(int src, int dst, std::vector<LinCoef>& out) {
    double inv_scale = (double)dst / src;
    double scale = 1.0 / inv_scale;
    out.resize(dst);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        out[d].s = s;
        out[d].c0 = (short)cv_round((1.f - f) * 2048.f);
        out[d].c1 = (short)cv_round(f * 2048.f);
    }
}
static inline void resize_linear(const Image8& src, Image8& dst) {   // dst.w/h preset
    std::vector<LinCoef> cx, cy;
    linear_coeffs(src.w, dst.w, cx);
    linear_coeffs(src.h, dst.h, cy);
    std::vector<int> rows[2] = {std::vector<int>(dst.w), std::vector<int>(dst.w)};
    int have[2] = {-1, -1};
    auto hrow = [&](int sy) -> const int* {              // horizontal pass of source row sy, cached (2 rows live)
        for (int k = 0; k < 2; ++k) if (have[k] == sy) return rows[k].data();
        const int k = (have[0] < have[1]) ? 0 : 1;       // evict the older (smaller) row
        const uint8_t* s = src.row(sy);
        int* o = rows[k].data();
        for (int x = 0; x < dst.w; ++x) { const int x0 = cx[x].s, x1 = std::min(x0 + 1, src.w - 1); o[x] = s[x0] * cx[x].c0 + s[x1] * cx[x].c1; }
        have[k] = sy;
        return o;
    };
    for (int y = 0; y < dst.h; ++y) {
        const int y0 = cy[y].s, y1 = std::min(y0 + 1, src.h - 1);
        const int* r0 = hrow(y0);
        const int* r1 = hrow(y1);
        r0 = hrow(y0);                                    // (y1's fill may have evicted nothing of y0: 2 slots, distinct rows)
        const int b0 = cy[y].c0, b1 = cy[y].c1;
        uint8_t* o = dst.row(y);
        for (int x = 0; x < dst.w; ++x)
            o[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---- cv::resize INTER_LINEAR_EXACT, 8UC1, SURVEY A.5 ------------------------------------------
struct ExCoef { int s; int c0, c1; };
static inline void exact_coeffs(int src, int dst, double scale /*dst/src factor*/, std::vector<ExCoef>& out) {
    out.resize(dst);
    for (int d = 0; d < dst; ++d) {
        double s = (d + 0.5) / scale - 0.5;
        if (s < 0) { out[d] = {0, 256, 0}; continue; }
        if (s >= src - 1) { out[d] = {src - 1, 256, 0}; continue; }
        int o = (int)std::floor(s);
        int c1 = (int)lrint((s - o) * 256.0);
        out[d] = {o, 256 - c1, c1};
    }
}
static inline void resize_linear_exact(const Image8& src, double fx, double fy, Image8& dst) {
    dst = Image8((int)lrint(src.w * fx), (int)lrint(src.h * fy));
    std::vector<ExCoef> cx, cy;
    exact_coeffs(src.w, dst.w, fx, cx);
    exact_coeffs(src.h, dst.h, fy, cy);
    std::vector<uint32_t> r0(dst.w), r1(dst.w);
    for (int y = 0; y < dst.h; ++y) {
        int y0 = cy[y].s, y1 = std::min(y0 + 1, src.h - 1);
        const uint8_t* s0 = src.row(y0);
        const uint8_t* s1 = src.row(y1);
        for (int x = 0; x < dst.w; ++x) {
            int x0 = cx[x].s, x1 = std::min(x0 + 1, src.w - 1);
            r0[x] = (uint32_t)(s0[x0] * cx[x].c0 + s0[x1] * cx[x].c1);   // 8.8
            r1[x] = (uint32_t)(s1[x0] * cx[x].c0 + s1[x1] * cx[x].c1);
        }
        uint8_t* o = dst.row(y);
        for (int x = 0; x < dst.w; ++x) {
            uint32_t acc = r0[x] * (uint32_t)cy[y].c0 + r1[x] * (uint32_t)cy[y].c1;   // 16.16
            o[x] = (uint8_t)((acc + (1u << 15)) >> 16);
        }
    }
}

// ---- cv::GaussianBlur on CV_8U (fixed point 8.8 kernels), SURVEY A.3 --------------------------
static inline std::vector<int> gauss_kernel_q8(int n, double sigma) {
    std::vector<double> k(n);
    double sum = 0, c = (n - 1) * 0.5;
    for (int i = 0; i < n; ++i) { double x = i - c; k[i] = std::exp(-(x * x) / (2.0 * sigma * sigma)); sum += k[i]; }
    for (int i = 0; i < n; ++i) k[i] /= sum;
    std::vector<int> q(n, 0);
    double carry = 0; int acc = 0;
    for (int i = 0; i < n / 2; ++i) {
        double adj = k[i] * 256.0 + carry;
        int v = (int)lrint(adj);
        carry = adj - v;
        q[i] = q[n - 1 - i] = v;
        acc += 2 * v;
    }
    q[n / 2] = 256 - acc;
    return q;
}
static inline void gaussian_blur_q8(const Image8& src, const std::vector<int>& q, Image8& dst) {
    const int n = (int)q.size(), r = n / 2, w = src.w, h = src.h;
    std::vector<uint16_t> tmp((size_t)w * h);
    std::vector<uint8_t> prow((size_t)w + 2 * r);
    std::vector<uint32_t> acc(w);
    for (int y = 0; y < h; ++y) {                       // horizontal pass, 8.8 fixed point
        const uint8_t* s = src.row(y);
        for (int x = -r; x < w + r; ++x) prow[x + r] = s[reflect101(x, w)];
        std::fill(acc.begin(), acc.end(), 0u);
        for (int k = 0; k < n; ++k) {
            const uint32_t qk = (uint32_t)q[k];
            const uint8_t* pk = prow.data() + k;
            for (int x = 0; x < w; ++x) acc[x] += qk * pk[x];
        }
        uint16_t* t = tmp.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) t[x] = (uint16_t)acc[x];
    }
    dst = Image8(w, h);
    for (int y = 0; y < h; ++y) {                       // vertical pass, 16.16 fixed point
        std::fill(acc.begin(), acc.end(), 0u);
        for (int k = 0; k < n; ++k) {
            const uint16_t* t = tmp.data() + (size_t)reflect101(y + k - r, h) * w;
            const uint32_t qk = (uint32_t)q[k];
            for (int x = 0; x < w; ++x) acc[x] += qk * t[x];
        }
        uint8_t* o = dst.row(y);
        for (int x = 0; x < w; ++x) o[x] = (uint8_t)((acc[x] + (1u << 15)) >> 16);
    }
}

// ---- cv::Sobel 8U->16S ksize 3, SURVEY A.8 ----------------------------------------------------
static inline void sobel3(const Image8& src, std::vector<int16_t>& dx, std::vector<int16_t>& dy) {
    const int w = src.w, h = src.h;
    dx.assign((size_t)w * h, 0); dy.assign((size_t)w * h, 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* a = src.row(reflect101(y - 1, h));
        const uint8_t* b = src.row(y);
        const uint8_t* c = src.row(reflect101(y + 1, h));
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            dx[(size_t)y * w + x] = (int16_t)((a[xp] - a[xm]) + 2 * (b[xp] - b[xm]) + (c[xp] - c[xm]));
            dy[(size_t)y * w + x] = (int16_t)((c[xm] + 2 * c[x] + c[xp]) - (a[xm] + 2 * a[x] + a[xp]));
        }
    }
}

// ---- cv::FAST TYPE_9_16, SURVEY A.1 -----------------------------------------------------------
static const int FAST_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int FAST_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// threshold-free corner score: max over the 16 cyclic 9-arcs of min(d) / min(-d), minus 1.
// corner(th) <=> score >= th.  Returned clamped to [0,254].
static inline int fast_score_px(const uint8_t* p, int stride) {
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[FAST_DY[k] * stride + FAST_DX[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -255, B = -255;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        A = std::max(A, mn);
        B = std::max(B, -mx);
    }
    int s = std::max(A, B) - 1;
    return s < 0 ? 0 : s;
}
// full-image score map; border of 3 px = 0.  Scores below `th` are never consulted by FAST(th) and are stored as 0,
// which allows the usual early rejection (any 9-arc contains one pixel of every opposite pair).
static inline void fast_score_map(const uint8_t* img, int w, int h, int stride, std::vector<uint8_t>& score, int th = 1) {
    score.assign((size_t)w * h, 0);
    std::vector<uint8_t> cand(w, 0);
    const unsigned band = (unsigned)(2 * th);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* r0 = img + (size_t)y * stride;
        const uint8_t *u3 = r0 - 3 * stride, *d3 = r0 + 3 * stride, *u2 = r0 - 2 * stride, *d2 = r0 + 2 * stride;
        // pass 1 (auto-vectorised): a 9-arc contains one pixel of every opposite pair, so a pair inside the band rejects
        for (int x = 3; x < w - 3; ++x) {
            const int v = r0[x] - th;
            const bool p0 = (unsigned)(d3[x] - v) <= band, p8 = (unsigned)(u3[x] - v) <= band;
            const bool p4 = (unsigned)(r0[x + 3] - v) <= band, p12 = (unsigned)(r0[x - 3] - v) <= band;
            const bool p2 = (unsigned)(d2[x + 2] - v) <= band, p10 = (unsigned)(u2[x - 2] - v) <= band;
            const bool p6 = (unsigned)(u2[x + 2] - v) <= band, p14 = (unsigned)(d2[x - 2] - v) <= band;
            cand[x] = !((p0 & p8) | (p4 & p12) | (p2 & p10) | (p6 & p14));
        }
        uint8_t* srow = score.data() + (size_t)y * w;
        for (int x = 3; x < w - 3; ++x) {
            if (!cand[x]) continue;
            const uint8_t* p = r0 + x;
            const int v = p[0], lo = v - th, hi = v + th;
            unsigned mpos = 0, mneg = 0;
            for (int k = 0; k < 16; ++k) { const int q = p[FAST_DY[k] * stride + FAST_DX[k]]; mpos |= (unsigned)(q < lo) << k; mneg |= (unsigned)(q > hi) << k; }
            auto run9 = [](unsigned m) { unsigned u = m | (m << 16); unsigned r = u & (u >> 1); r &= r >> 2; r &= r >> 4; r &= u >> 8; return r != 0; };
            if (!run9(mpos) && !run9(mneg)) continue;
            const int s = fast_score_px(p, stride);
            srow[x] = (uint8_t)(s >= th ? s : 0);
        }
    }
}
struct FastKp { int x, y, score; };
// cv::FAST(window, th, nonmax=true) on a w x h window (row-major output, window-local coords)
static inline void fast_detect(const uint8_t* img, int w, int h, int stride, int th, bool nms, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<uint8_t> sc;
    fast_score_map(img, w, h, stride, sc, th);
    auto S = [&](int x, int y) -> int {
        if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
        int s = sc[(size_t)y * w + x];
        return s >= th ? s : 0;
    };
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (s < th) continue;
            if (nms) {
                bool ok = s > S(x - 1, y - 1) && s > S(x, y - 1) && s > S(x + 1, y - 1) && s > S(x - 1, y) &&
                          s > S(x + 1, y) && s > S(x - 1, y + 1) && s > S(x, y + 1) && s > S(x + 1, y + 1);
                if (!ok) continue;
            }
            out.push_back({x, y, s});
        }
}

// ---- Hamming distance on 32-byte descriptors (src/ORBmatcher.cc:1795-1811, src/LineMatcher.cpp:134-150)
static inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

}  // namespace orc
