// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Not part of the product path.
//
// CPU restatement of ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc), with the OpenCV
// primitives from cvprim.hpp.  Determinism choices: quadtree tie-break by node
// creation sequence (C.1), no FMA contraction (C.4), cosf/sinf from glibc (C.5).
#include "oracle.h"
#include "cvprim.hpp"
#include "../include/olf_brief_pattern.h"
#include <list>
#include <utility>

using namespace orc;

struct orc_orb {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactor;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> feats_per_level;
    int umax[16];
    std::vector<Image8> pyr;            // mvImagePyramid (unpadded ROI contents)
    std::vector<int> last_cand;         // {level,x,y,score}*
};

// ORBextractor::ORBextractor  (src/ORBextractor.cc:412-472)
extern "C" orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    if (nlevels < 1 || nlevels > OLF_MAX_LEVELS || nfeatures < 0) return nullptr;
    orc_orb* h = new orc_orb();
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->iniTh = iniTh; h->minTh = minTh; h->scaleFactor = scaleFactor;
    h->scale.resize(nlevels); h->sigma2.resize(nlevels); h->inv_scale.resize(nlevels); h->inv_sigma2.resize(nlevels);
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { h->scale[i] = h->scale[i - 1] * scaleFactor; h->sigma2[i] = h->scale[i] * h->scale[i]; }
    for (int i = 0; i < nlevels; i++) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    h->feats_per_level.resize(nlevels);
    float factor = 1.0f / scaleFactor;
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { h->feats_per_level[l] = cv_round(nDesired); sum += h->feats_per_level[l]; nDesired *= factor; }
    h->feats_per_level[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax (:456-471)
    const int HP = 15;
    int v, v0, vmax = cv_floor(HP * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HP * sqrtf(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) h->umax[v] = cv_round(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) { while (h->umax[v0] == h->umax[v0 + 1]) ++v0; h->umax[v] = v0; ++v0; }
    h->pyr.resize(nlevels);
    return h;
}
extern "C" void orc_orb_destroy(orc_orb* h) { delete h; }

// ORBextractor::ComputePyramid (src/ORBextractor.cc:1109-1134).  The 19-px reflect border the reference adds
// is never read on the hot path, so levels are stored unpadded.
static void compute_pyramid(orc_orb* h, const uint8_t* img, int w, int hgt, int stride) {
    for (int l = 0; l < h->nlevels; ++l) {
        float s = h->inv_scale[l];
        int lw = cv_round((float)w * s), lh = cv_round((float)hgt * s);
        if (l == 0) {
            h->pyr[0] = Image8(w, hgt);
            for (int y = 0; y < hgt; ++y) memcpy(h->pyr[0].row(y), img + (size_t)y * stride, w);
        } else {
            h->pyr[l] = Image8(lw, lh);
            resize_linear(h->pyr[l - 1], h->pyr[l]);
        }
    }
}

struct Cand { float x, y; int score; float angle; };   // x,y relative to (minBorderX,minBorderY)

// per-cell FAST with threshold fallback (src/ORBextractor.cc:773-831)
static void detect_level(const orc_orb* h, int level, std::vector<Cand>& out) {
    const Image8& im = h->pyr[level];
    const float W = 30;
    const int minBorderX = 16, minBorderY = 16;
    const int maxBorderX = im.w - 19 + 3, maxBorderY = im.h - 19 + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    out.clear();
    if (width < W || height < W) return;     // reference would divide by zero; no keypoints on such a level
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    std::vector<FastKp> cell;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            fast_detect(im.row(y0) + x0, cw, ch, im.w, h->iniTh, true, cell);
            if (cell.empty()) fast_detect(im.row(y0) + x0, cw, ch, im.w, h->minTh, true, cell);
            for (const FastKp& k : cell) out.push_back({(float)(k.x + j * wCell), (float)(k.y + i * hCell), k.score, -1.f});
        }
    }
}

// ---- quadtree (src/ORBextractor.cc:483-765) -----------------------------------------------------
struct Node {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<int> keys;        // indices into the candidate array, in insertion order
    bool noMore = false;
    long seq = 0;                 // creation sequence number (canonical tie-break)
    std::list<Node>::iterator lit;
};
static void divide_node(const Node& n, const std::vector<Cand>& c, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)ceilf((float)(n.URx - n.ULx) / 2);
    const int halfY = (int)ceilf((float)(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (int idx : n.keys) {
        const Cand& kp = c[idx];
        if (kp.x < n1.URx) { if (kp.y < n1.BRy) n1.keys.push_back(idx); else n3.keys.push_back(idx); }
        else if (kp.y < n1.BRy) n2.keys.push_back(idx);
        else n4.keys.push_back(idx);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

static void distribute_quadtree(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& result) {
    result.clear();
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) {             // taller-than-wide level: reference divides by zero (hX = inf); treat as one root
        // canonical: single root covering the level
    }
    const int nRoots = std::max(nIni, 1);
    const float hX = (float)(maxX - minX) / nRoots;
    std::list<Node> nodes;
    std::vector<Node*> ini(nRoots);
    long seq = 0;
    for (int i = 0; i < nRoots; i++) {
        Node ni;
        ni.ULx = (int)(hX * (float)i); ni.ULy = 0; ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY; ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < c.size(); i++) {
        int r = (int)(c[i].x / hX);
        if (r >= nRoots) r = nRoots - 1;
        ini[r]->keys.push_back((int)i);
    }
    auto lit = nodes.begin();
    while (lit != nodes.end()) {
        if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
        else if (lit->keys.empty()) lit = nodes.erase(lit);
        else lit++;
    }
    bool finish = false;
    std::vector<std::pair<int, Node*>> sizeAndNode;
    auto push_child = [&](Node& n, int& nToExpand, bool count) {
        if (n.keys.size() > 0) {
            n.seq = seq++;
            nodes.push_front(n);
            if (n.keys.size() > 1) {
                if (count) nToExpand++;
                sizeAndNode.push_back(std::make_pair((int)n.keys.size(), &nodes.front()));
                nodes.front().lit = nodes.begin();
            }
        }
    };
    auto cmp = [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
        if (a.first != b.first) return a.first < b.first;
        return a.second->seq < b.second->seq;
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        lit = nodes.begin();
        int nToExpand = 0;
        sizeAndNode.clear();
        while (lit != nodes.end()) {
            if (lit->noMore) { lit++; continue; }
            Node n1, n2, n3, n4;
            divide_node(*lit, c, n1, n2, n3, n4);
            push_child(n1, nToExpand, true); push_child(n2, nToExpand, true);
            push_child(n3, nToExpand, true); push_child(n4, nToExpand, true);
            lit = nodes.erase(lit);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
        else if (((int)nodes.size() + nToExpand * 3) > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                std::vector<std::pair<int, Node*>> prev = sizeAndNode;
                sizeAndNode.clear();
                std::sort(prev.begin(), prev.end(), cmp);
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    divide_node(*prev[j].second, c, n1, n2, n3, n4);
                    int dummy = 0;
                    push_child(n1, dummy, false); push_child(n2, dummy, false);
                    push_child(n3, dummy, false); push_child(n4, dummy, false);
                    nodes.erase(prev[j].second->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    for (auto it = nodes.begin(); it != nodes.end(); ++it) {
        int best = it->keys[0];
        float maxResp = (float)c[best].score;
        for (size_t k = 1; k < it->keys.size(); k++)
            if ((float)c[it->keys[k]].score > maxResp) { best = it->keys[k]; maxResp = (float)c[best].score; }
        result.push_back(best);
    }
}

// IC_Angle (src/ORBextractor.cc:79-106)
static float ic_angle(const Image8& im, int px, int py, const int* umax) {
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = im.row(py) + px;
    for (int u = -15; u <= 15; ++u) m_10 += u * center[u];
    const int step = im.w;
    for (int v = 1; v <= 15; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2_deg((float)m_01, (float)m_10);
}

// computeOrbDescriptor (src/ORBextractor.cc:110-149)
static void orb_descriptor(const Image8& blurred, int px, int py, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t* center = blurred.row(py) + px;
    const int step = blurred.w;
    const signed char* pat = OLF_BRIEF_PATTERN;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k, pat += 4) {
            float x0 = pat[0], y0 = pat[1], x1 = pat[2], y1 = pat[3];
            int t0 = center[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
            int t1 = center[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ORBextractor::operator() (src/ORBextractor.cc:1045-1107)
extern "C" int orc_orb_extract(orc_orb* h, const uint8_t* img, int w, int hgt, int stride,
                               olf_keypoint* kps, uint8_t* desc, int cap, int* n) {
    if (!h || !n) return OLF_ERR_ARG;
    *n = 0;
    if (!img || w <= 0 || hgt <= 0) return OLF_OK;       // _image.empty() -> silent return (:1048)
    compute_pyramid(h, img, w, hgt, stride);
    static const std::vector<int> q7 = gauss_kernel_q8(7, 2.0);
    h->last_cand.clear();
    int total = 0;
    std::vector<Cand> cand;
    std::vector<int> keep;
    for (int level = 0; level < h->nlevels; ++level) {
        const Image8& im = h->pyr[level];
        detect_level(h, level, cand);
        for (const Cand& c : cand) { h->last_cand.push_back(level); h->last_cand.push_back((int)c.x + 16); h->last_cand.push_back((int)c.y + 16); h->last_cand.push_back(c.score); }
        if (cand.empty()) continue;
        const int minBX = 16, minBY = 16, maxBX = im.w - 16, maxBY = im.h - 16;
        distribute_quadtree(cand, minBX, maxBX, minBY, maxBY, h->feats_per_level[level], keep);
        if (keep.empty()) continue;
        if (total + (int)keep.size() > cap) return OLF_ERR_CAPACITY;
        Image8 blurred;
        gaussian_blur_q8(im, q7, blurred);
        const int scaledPatchSize = (int)(31 * h->scale[level]);
        for (int idx : keep) {
            const Cand& c = cand[idx];
            int px = (int)c.x + minBX, py = (int)c.y + minBY;
            olf_keypoint& kp = kps[total];
            kp.angle = ic_angle(im, px, py, h->umax);
            kp.response = (float)c.score;
            kp.octave = level;
            kp.size = (float)scaledPatchSize;
            orb_descriptor(blurred, px, py, kp.angle, desc + (size_t)total * 32);
            kp.x = (float)px; kp.y = (float)py;
            if (level != 0) { kp.x *= h->scale[level]; kp.y *= h->scale[level]; }
            ++total;
        }
    }
    *n = total;
    return OLF_OK;
}

extern "C" int orc_orb_level_size(const orc_orb* h, int level, int* w, int* hh) {
    if (!h || level < 0 || level >= h->nlevels) return OLF_ERR_ARG;
    *w = h->pyr[level].w; *hh = h->pyr[level].h; return OLF_OK;
}
extern "C" int orc_orb_get_level(orc_orb* h, int level, uint8_t* dst, int dst_stride) {
    if (!h || level < 0 || level >= h->nlevels) return OLF_ERR_ARG;
    const Image8& im = h->pyr[level];
    for (int y = 0; y < im.h; ++y) memcpy(dst + (size_t)y * dst_stride, im.row(y), im.w);
    return OLF_OK;
}
extern "C" int orc_orb_scale_factors(const orc_orb* h, float* s, float* is, float* s2, float* is2) {
    if (!h) return OLF_ERR_ARG;
    for (int i = 0; i < h->nlevels; ++i) { if (s) s[i] = h->scale[i]; if (is) is[i] = h->inv_scale[i]; if (s2) s2[i] = h->sigma2[i]; if (is2) is2[i] = h->inv_sigma2[i]; }
    return OLF_OK;
}
extern "C" int orc_orb_features_per_level(const orc_orb* h, int* out) {
    if (!h) return OLF_ERR_ARG;
    for (int i = 0; i < h->nlevels; ++i) out[i] = h->feats_per_level[i];
    return OLF_OK;
}
extern "C" int orc_orb_last_candidates(orc_orb* h, int* out, int cap, int* n) {
    if (!h) return OLF_ERR_ARG;
    int cnt = (int)h->last_cand.size() / 4;
    *n = cnt;
    if (cnt > cap) return OLF_ERR_CAPACITY;
    memcpy(out, h->last_cand.data(), h->last_cand.size() * sizeof(int));
    return OLF_OK;
}
const Image8* orc_orb_level_image(const orc_orb* h, int level) { return &h->pyr[level]; }
const float* orc_orb_scales(const orc_orb* h) { return h->scale.data(); }
const float* orc_orb_inv_scales(const orc_orb* h) { return h->inv_scale.data(); }

// ---- thin wrappers exposing the cv primitives for the cv2 pinning tests ---------------------------
extern "C" void orc_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    Image8 s(sw, sh), d(dw, dh);
    memcpy(s.d.data(), src, (size_t)sw * sh);
    resize_linear(s, d);
    memcpy(dst, d.d.data(), (size_t)dw * dh);
}
extern "C" void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, double fx, double fy, uint8_t* dst, int* dw, int* dh) {
    Image8 s(sw, sh), d;
    memcpy(s.d.data(), src, (size_t)sw * sh);
    resize_linear_exact(s, fx, fy, d);
    *dw = d.w; *dh = d.h;
    if (dst) memcpy(dst, d.d.data(), (size_t)d.w * d.h);
}
extern "C" void orc_gaussian_blur(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* dst) {
    Image8 s(w, h), d;
    memcpy(s.d.data(), src, (size_t)w * h);
    gaussian_blur_q8(s, gauss_kernel_q8(ksize, sigma), d);
    memcpy(dst, d.d.data(), (size_t)w * h);
}
extern "C" void orc_gauss_kernel_q8(int ksize, double sigma, int* out) {
    std::vector<int> q = gauss_kernel_q8(ksize, sigma);
    for (int i = 0; i < ksize; ++i) out[i] = q[i];
}
extern "C" void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) {
    Image8 s(w, h);
    memcpy(s.d.data(), src, (size_t)w * h);
    std::vector<int16_t> a, b;
    sobel3(s, a, b);
    memcpy(dx, a.data(), a.size() * 2); memcpy(dy, b.data(), b.size() * 2);
}
extern "C" float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
extern "C" int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int th, int nms, int* xys /*cap x 3*/, int cap) {
    std::vector<FastKp> k;
    fast_detect(img, w, h, stride, th, nms != 0, k);
    int n = (int)k.size();
    for (int i = 0; i < n && i < cap; ++i) { xys[3 * i] = k[i].x; xys[3 * i + 1] = k[i].y; xys[3 * i + 2] = k[i].score; }
    return n;
}
extern "C" void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score) {
    std::vector<uint8_t> s;
    fast_score_map(img, w, h, stride, s);
    memcpy(score, s.data(), s.size());
}

extern "C" void orc_cosf_sinf(const float* x, long n, float* c, float* s) { for (long i = 0; i < n; ++i) { c[i] = cosf(x[i]); s[i] = sinf(x[i]); } }
