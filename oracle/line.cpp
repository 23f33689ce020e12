// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Not part of the product path.
//
// CPU restatement of the line half of the reference front end:
//   * cv::LineSegmentDetector (refine = LSD_REFINE_NONE) -- un-vendored OpenCV, semantics of cv2 4.13.0
//     (SURVEY Appendix A.6), configured as in Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:246-253
//   * LSDDetectorC::detectImpl KeyLine construction   LSDDetector_custom.cpp:76-102, 264-308
//   * Lineextractor::operator()                       src/LineExtractor.cc:31-67
//   * BinaryDescriptor::compute / computeLBD          binary_descriptor_custom.cpp:217-259, 350-412, 539-687, 1026-1372
// Canonical choices: stable seed order (bin desc, then row-major), stable sort of keylines by
// response (C.3), no FMA contraction (C.4), libm double cos/sin/atan2 where the reference promotes to double (C.5).
#include "oracle.h"
#include "cvprim.hpp"
#include <cstdio>

using namespace orc;

static const double NOTDEF = -1024.0;
static const double M_3_2_PI = 4.71238898038;
static const double M_2__PI = 6.28318530718;
static const double DEG_TO_RADS = M_PI / 180.0;

struct LsdState {
    int W = 0, H = 0;
    std::vector<double> angles, modgrad;
    std::vector<uint8_t> used;
};

static inline bool is_aligned(const LsdState& s, int x, int y, double theta, double prec) {
    if (x < 0 || y < 0 || x >= s.W || y >= s.H) return false;
    const double a = s.angles[(size_t)y * s.W + x];
    if (a == NOTDEF) return false;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > M_3_2_PI) {
        n_theta -= M_2__PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

struct RegPt { int x, y; double angle, modgrad; };

static void region_grow(LsdState& s, int sx, int sy, std::vector<RegPt>& reg, double& reg_angle, double prec) {
    reg.clear();
    reg_angle = s.angles[(size_t)sy * s.W + sx];
    reg.push_back({sx, sy, reg_angle, s.modgrad[(size_t)sy * s.W + sx]});
    float sumdx = (float)std::cos(reg_angle);
    float sumdy = (float)std::sin(reg_angle);
    s.used[(size_t)sy * s.W + sx] = 1;
    for (size_t i = 0; i < reg.size(); i++) {
        const int px = reg[i].x, py = reg[i].y;
        int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, s.W - 1);
        int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, s.H - 1);
        for (int yy = yy_min; yy <= yy_max; ++yy)
            for (int xx = xx_min; xx <= xx_max; ++xx) {
                uint8_t& u = s.used[(size_t)yy * s.W + xx];
                if (u != 1 && is_aligned(s, xx, yy, reg_angle, prec)) {
                    const double angle = s.angles[(size_t)yy * s.W + xx];
                    u = 1;
                    reg.push_back({xx, yy, angle, s.modgrad[(size_t)yy * s.W + xx]});
                    sumdx += (float)std::cos((double)(float)angle);
                    sumdy += (float)std::sin((double)(float)angle);
                    reg_angle = fast_atan2_deg(sumdy, sumdx) * DEG_TO_RADS;
                }
            }
    }
}

static inline double angle_diff(double a, double b) {
    double diff = a - b;
    while (diff <= -M_PI) diff += M_2__PI;
    while (diff > M_PI) diff -= M_2__PI;
    if (diff < 0) diff = -diff;
    return diff;
}

static void region2rect(const std::vector<RegPt>& reg, double reg_angle, double prec, double out[4]) {
    double x = 0, y = 0, sum = 0;
    for (const RegPt& p : reg) { x += (double)p.x * p.modgrad; y += (double)p.y * p.modgrad; sum += p.modgrad; }
    x /= sum; y /= sum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (const RegPt& p : reg) {
        double dx = (double)p.x - x, dy = (double)p.y - y;
        Ixx += dy * dy * p.modgrad;
        Iyy += dx * dx * p.modgrad;
        Ixy -= dx * dy * p.modgrad;
    }
    double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                                    : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
    theta *= DEG_TO_RADS;
    if (angle_diff(theta, reg_angle) > prec) theta += M_PI;
    double dx = std::cos(theta), dy = std::sin(theta);
    double l_min = 0, l_max = 0;
    for (const RegPt& p : reg) {
        double regdx = (double)p.x - x, regdy = (double)p.y - y;
        double l = regdx * dx + regdy * dy;
        if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
    }
    out[0] = x + l_min * dx; out[1] = y + l_min * dy; out[2] = x + l_max * dx; out[3] = y + l_max * dy;
}

struct LsdStats { int n_defined, n_regions, n_accepted, max_region, min_reg_size, ws, hs; };

// cv::LineSegmentDetectorImpl::detect / lsd() with refine NONE
static void lsd_detect(const Image8& img, const olf_line_params& P, std::vector<float>& segs, LsdStats* st) {
    segs.clear();
    const double scale = P.lsd_scale;
    const double prec = M_PI * P.lsd_ang_th / 180.0;
    const double p = P.lsd_ang_th / 180.0;
    const double rho = P.lsd_quant / std::sin(prec);
    const int n_bins = P.lsd_n_bins;
    Image8 scaled;
    if (scale != 1.0) {
        const double sigma = (scale < 1) ? (P.lsd_sigma_scale / scale) : P.lsd_sigma_scale;
        const double sprec = 3;
        const unsigned int hk = (unsigned int)(std::ceil(sigma * std::sqrt(2 * sprec * std::log(10.0))));
        Image8 g;
        gaussian_blur_q8(img, gauss_kernel_q8(1 + 2 * (int)hk, sigma), g);
        resize_linear_exact(g, scale, scale, scaled);
    } else scaled = img;
    LsdState s;
    const int W = s.W = scaled.w, H = s.H = scaled.h;
    s.angles.assign((size_t)W * H, NOTDEF);
    s.modgrad.assign((size_t)W * H, 0.0);
    s.used.assign((size_t)W * H, 0);
    double max_grad = -1;
    int n_defined = 0;
    for (int y = 0; y < H - 1; ++y) {
        const uint8_t* r0 = scaled.row(y);
        const uint8_t* r1 = scaled.row(y + 1);
        for (int x = 0; x < W - 1; ++x) {
            int DA = r1[x + 1] - r0[x];
            int BC = r0[x + 1] - r1[x];
            int gx = DA + BC, gy = DA - BC;
            double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
            s.modgrad[(size_t)y * W + x] = norm;
            if (norm <= rho) s.angles[(size_t)y * W + x] = NOTDEF;
            else {
                s.angles[(size_t)y * W + x] = fast_atan2_deg((float)gx, (float)-gy) * DEG_TO_RADS;
                if (norm > max_grad) max_grad = norm;
                ++n_defined;
            }
        }
    }
    // stable counting sort of the (W-1)x(H-1) interior by bin, descending
    const double bin_coef = (max_grad > 0) ? (double)(n_bins - 1) / max_grad : 0;
    std::vector<int> cnt(n_bins + 1, 0);
    std::vector<uint16_t> bin((size_t)W * H, 0);
    for (int y = 0; y < H - 1; ++y)
        for (int x = 0; x < W - 1; ++x) {
            int b = (int)(s.modgrad[(size_t)y * W + x] * bin_coef);
            bin[(size_t)y * W + x] = (uint16_t)b;
            cnt[b]++;
        }
    std::vector<int> start(n_bins + 1, 0);
    { int acc = 0; for (int b = n_bins - 1; b >= 0; --b) { start[b] = acc; acc += cnt[b]; } }
    std::vector<int> order((size_t)(W - 1) * (H - 1));
    for (int y = 0; y < H - 1; ++y)
        for (int x = 0; x < W - 1; ++x) order[start[bin[(size_t)y * W + x]]++] = y * W + x;

    const double logNT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const unsigned int min_reg_size = (unsigned int)(-logNT / std::log10(p));
    std::vector<RegPt> reg;
    int n_regions = 0, n_acc = 0, max_region = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const int idx = order[i];
        if (s.used[idx] == 0 && s.angles[idx] != NOTDEF) {
            double reg_angle;
            region_grow(s, idx % W, idx / W, reg, reg_angle, prec);
            ++n_regions;
            max_region = std::max(max_region, (int)reg.size());
            if (reg.size() < min_reg_size) continue;
            double r[4];
            region2rect(reg, reg_angle, prec, r);
            for (int k = 0; k < 4; ++k) { r[k] += 0.5; if (scale != 1.0) r[k] /= scale; }
            segs.push_back((float)r[0]); segs.push_back((float)r[1]); segs.push_back((float)r[2]); segs.push_back((float)r[3]);
            ++n_acc;
        }
    }
    if (st) *st = {n_defined, n_regions, n_acc, max_region, (int)min_reg_size, W, H};
}

// checkLineExtremes (LSDDetector_custom.cpp:76-102)
static void check_extremes(float e[4], int w, int h) {
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= w) e[0] = (float)w - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= w) e[2] = (float)w - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= h) e[1] = (float)h - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= h) e[3] = (float)h - 1.0f;
}

// KeyLine construction (LSDDetector_custom.cpp:264-308), numOctaves = 1, scale = (int)lsd_scale
static void make_keylines(const std::vector<float>& segs, int w, int h, double min_length, std::vector<olf_keyline>& out) {
    out.clear();
    int class_counter = -1;
    const float octaveScale = 1.0f;       // pow((float)scale, 0)
    for (size_t k = 0; k < segs.size() / 4; ++k) {
        float e[4] = {segs[4 * k], segs[4 * k + 1], segs[4 * k + 2], segs[4 * k + 3]};
        check_extremes(e, w, h);
        double length = (float)std::sqrt(std::pow((double)(e[0] - e[2]), 2) + std::pow((double)(e[1] - e[3]), 2));
        if (!(length > min_length)) continue;
        olf_keyline kl;
        kl.startPointX = e[0] * octaveScale; kl.startPointY = e[1] * octaveScale;
        kl.endPointX = e[2] * octaveScale; kl.endPointY = e[3] * octaveScale;
        kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1]; kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
        kl.lineLength = (float)length;
        // cv::LineIterator(img, Point2f, Point2f).count, 8-connected; endpoints already inside the image
        int x1 = cv_round(e[0]), y1 = cv_round(e[1]), x2 = cv_round(e[2]), y2 = cv_round(e[3]);
        kl.numOfPixels = std::max(std::abs(x2 - x1), std::abs(y2 - y1)) + 1;
        // atan2( float, float ) (LSDDetector_custom.cpp:298): precomp_custom.hpp -> bitarray_custom.hpp -> <math.h>, whose C++ wrapper puts
        // the float overloads into the global namespace, so the reference calls atan2f (pinned by oracle/_ref, tests/test_oracle_vs_ref.py)
        kl.angle = atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);
        kl.class_id = ++class_counter;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, h);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2; kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        out.push_back(kl);
    }
}

// ---- LBD (binary_descriptor_custom.cpp) ---------------------------------------------------------
static const int LBD_COMB[32][2] = {{0,1},{0,2},{0,3},{0,4},{0,5},{0,6},{1,2},{1,3},{1,4},{1,5},{1,6},{2,3},{2,4},{2,5},{2,6},{2,7},
    {2,8},{3,4},{3,5},{3,6},{3,7},{3,8},{4,5},{4,6},{4,7},{4,8},{5,6},{5,7},{5,8},{6,7},{6,8},{7,8}};

struct LbdWeights { double L[21], G[63]; };
static LbdWeights lbd_weights() {               // BinaryDescriptor ctor :217-259 (integer divisions preserved)
    LbdWeights w;
    const int wb = 7, nb = 9;
    double u = (wb * 3 - 1) / 2;
    double sigma = (wb * 2 + 1) / 2;
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < wb * 3; i++) { double dis = i - u; w.L[i] = std::exp(dis * dis * invsigma2); }
    u = (nb * wb - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < nb * wb; i++) { double dis = i - u; w.G[i] = std::exp(dis * dis * invsigma2); }
    return w;
}

// computeLBD for one line (:1026-1372) -> 72 floats
static void lbd_one(const olf_keyline& kl, const int16_t* pdx, const int16_t* pdy, int imw, int imh, const LbdWeights& wt, float* desVec) {
    const int NB = 9, WB = 7;
    float dL[2], dO[2];
    const short heightOfLSP = (short)(WB * NB);
    float pgdLBandSum[9] = {0}, ngdLBandSum[9] = {0}, pgdL2BandSum[9] = {0}, ngdL2BandSum[9] = {0};
    float pgdOBandSum[9] = {0}, ngdOBandSum[9] = {0}, pgdO2BandSum[9] = {0}, ngdO2BandSum[9] = {0};
    const short halfHeight = (heightOfLSP - 1) / 2;
    const short realWidth = (short)imw;
    const short imageWidth = realWidth - 1;
    const short imageHeight = (short)(imh - 1);
    const short lengthOfLSP = (short)kl.numOfPixels;
    const short halfWidth = (lengthOfLSP - 1) / 2;
    const float lineMiddlePointX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
    const float lineMiddlePointY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
    // cos( float ) / sin( float ) (binary_descriptor_custom.cpp:1130-1131) resolve to cosf / sinf for the same reason as atan2f above
    dL[0] = cosf(kl.angle);
    dL[1] = sinf(kl.angle);
    dO[0] = -dL[1];
    dO[1] = dL[0];
    float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
    float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
    for (short hID = 0; hID < heightOfLSP; hID++) {
        float sCorX = sCorX0, sCorY = sCorY0;
        float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
        for (short wID = 0; wID < lengthOfLSP; wID++) {
            short tempCor = (short)(int)std::round(sCorX);
            short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
            tempCor = (short)(int)std::round(sCorY);
            short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
            short dx = pdx[yCor * realWidth + xCor];
            short dy = pdy[yCor * realWidth + xCor];
            float gDL = dx * dL[0] + dy * dL[1];
            float gDO = dx * dO[0] + dy * dO[1];
            if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
            if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
            sCorX += dL[0];
            sCorY += dL[1];
        }
        sCorX0 -= dL[1];
        sCorY0 += dL[0];
        float coef = (float)wt.G[hID];
        pgdLRowSum = coef * pgdLRowSum;
        ngdLRowSum = coef * ngdLRowSum;
        float pgdL2RowSum = pgdLRowSum * pgdLRowSum;
        float ngdL2RowSum = ngdLRowSum * ngdLRowSum;
        pgdORowSum = coef * pgdORowSum;
        ngdORowSum = coef * ngdORowSum;
        float pgdO2RowSum = pgdORowSum * pgdORowSum;
        float ngdO2RowSum = ngdORowSum * ngdORowSum;
        auto fold = [&](short bandID, float c) {
            pgdLBandSum[bandID] += c * pgdLRowSum;
            ngdLBandSum[bandID] += c * ngdLRowSum;
            pgdL2BandSum[bandID] += c * c * pgdL2RowSum;
            ngdL2BandSum[bandID] += c * c * ngdL2RowSum;
            pgdOBandSum[bandID] += c * pgdORowSum;
            ngdOBandSum[bandID] += c * ngdORowSum;
            pgdO2BandSum[bandID] += c * c * pgdO2RowSum;
            ngdO2BandSum[bandID] += c * c * ngdO2RowSum;
        };
        short bandID = (short)(hID / WB);
        fold(bandID, (float)wt.L[hID % WB + WB]);
        bandID--;
        if (bandID >= 0) fold(bandID, (float)wt.L[hID % WB + 2 * WB]);
        bandID = bandID + 2;
        if (bandID < NB) fold(bandID, (float)wt.L[hID % WB]);
    }
    const float invN2 = (float)(1.0 / (WB * 2.0));
    const float invN3 = (float)(1.0 / (WB * 3.0));
    for (short bandID = 0; bandID < NB; bandID++) {
        float invN = (bandID == 0 || bandID == NB - 1) ? invN2 : invN3;
        short desID = bandID * 8;
        float temp = pgdLBandSum[bandID] * invN;
        desVec[desID] = temp;
        desVec[desID + 4] = sqrtf(pgdL2BandSum[bandID] * invN - temp * temp);
        temp = ngdLBandSum[bandID] * invN;
        desVec[desID + 1] = temp;
        desVec[desID + 5] = sqrtf(ngdL2BandSum[bandID] * invN - temp * temp);
        temp = pgdOBandSum[bandID] * invN;
        desVec[desID + 2] = temp;
        desVec[desID + 6] = sqrtf(pgdO2BandSum[bandID] * invN - temp * temp);
        temp = ngdOBandSum[bandID] * invN;
        desVec[desID + 3] = temp;
        desVec[desID + 7] = sqrtf(ngdO2BandSum[bandID] * invN - temp * temp);
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < NB; ++b) {
        const float* d = desVec + 8 * b;
        tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
        tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
    }
    tempM = 1 / sqrtf(tempM);
    tempS = 1 / sqrtf(tempS);
    for (int b = 0; b < NB; ++b) {
        float* d = desVec + 8 * b;
        d[0] = d[0] * tempM; d[1] = d[1] * tempM; d[2] = d[2] * tempM; d[3] = d[3] * tempM;
        d[4] = d[4] * tempS; d[5] = d[5] * tempS; d[6] = d[6] * tempS; d[7] = d[7] * tempS;
    }
    for (int i = 0; i < 72; i++) if ((double)desVec[i] > 0.4) desVec[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < 72; i++) temp += desVec[i] * desVec[i];
    temp = 1 / sqrtf(temp);
    for (int i = 0; i < 72; i++) desVec[i] = desVec[i] * temp;
}

static void lbd_binarise(const float* desVec, uint8_t* out) {   // :401-412, :662-666
    for (int c = 0; c < 32; ++c) {
        const float* f1 = desVec + 8 * LBD_COMB[c][0];
        const float* f2 = desVec + 8 * LBD_COMB[c][1];
        uint8_t r = 0;
        for (int i = 0; i < 8; ++i) if (f1[i] > f2[i]) r += (uint8_t)(1 << i);
        out[c] = r;
    }
}

static void lbd_compute(const Image8& img, const olf_keyline* kls, int n, uint8_t* desc, float* fdesc) {
    static const LbdWeights wt = lbd_weights();
    static const std::vector<int> q5 = gauss_kernel_q8(5, 1.0);
    Image8 blurred;
    gaussian_blur_q8(img, q5, blurred);                     // computeGaussianPyramid :358
    std::vector<int16_t> dx, dy;
    sobel3(blurred, dx, dy);                                // computeSobel :395-396
    float d[72];
    for (int i = 0; i < n; ++i) {
        lbd_one(kls[i], dx.data(), dy.data(), img.w, img.h, wt, d);
        if (desc) lbd_binarise(d, desc + (size_t)i * 32);
        if (fdesc) memcpy(fdesc + (size_t)i * 72, d, sizeof(d));
    }
}

// ---- C interface ---------------------------------------------------------------------------------
struct orc_line { olf_line_params p; LsdStats st; };

extern "C" orc_line* orc_line_create(const olf_line_params* p) {
    if (!p || p->lsd_refine != 0 || p->lsd_n_bins < 1 || p->lsd_n_bins > 65535) return nullptr;
    orc_line* h = new orc_line(); h->p = *p; h->st = {}; return h;
}
extern "C" void orc_line_destroy(orc_line* h) { delete h; }

static Image8 wrap(const uint8_t* img, int w, int h, int stride) {
    Image8 im(w, h);
    for (int y = 0; y < h; ++y) memcpy(im.row(y), img + (size_t)y * stride, w);
    return im;
}

extern "C" int orc_lsd_detect(orc_line* h, const uint8_t* img, int w, int hgt, int stride, float* segs, int cap, int* n) {
    if (!h || !img || !n) return OLF_ERR_ARG;
    std::vector<float> s;
    lsd_detect(wrap(img, w, hgt, stride), h->p, s, &h->st);
    *n = (int)s.size() / 4;
    if (*n > cap) return OLF_ERR_CAPACITY;
    memcpy(segs, s.data(), s.size() * sizeof(float));
    return OLF_OK;
}
extern "C" int orc_lsd_stats(const orc_line* h, int* out7) {
    out7[0] = h->st.n_defined; out7[1] = h->st.n_regions; out7[2] = h->st.n_accepted; out7[3] = h->st.max_region;
    out7[4] = h->st.min_reg_size; out7[5] = h->st.ws; out7[6] = h->st.hs;
    return OLF_OK;
}
extern "C" int orc_keylines_from_segments(const float* segs, int nseg, int w, int h, double min_length, olf_keyline* out, int cap, int* n) {
    std::vector<float> s(segs, segs + (size_t)nseg * 4);
    std::vector<olf_keyline> k;
    make_keylines(s, w, h, min_length, k);
    *n = (int)k.size();
    if (*n > cap) return OLF_ERR_CAPACITY;
    memcpy(out, k.data(), k.size() * sizeof(olf_keyline));
    return OLF_OK;
}
extern "C" int orc_lbd_compute(orc_line* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, uint8_t* desc) {
    if (!h || !img) return OLF_ERR_ARG;
    if (n == 0) return OLF_OK;                 // "keypoint list is empty": descriptors untouched (:556-560)
    lbd_compute(wrap(img, w, hgt, stride), kls, n, desc, nullptr);
    return OLF_OK;
}
extern "C" int orc_lbd_compute_float(orc_line* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, float* fdesc) {
    if (!h || !img) return OLF_ERR_ARG;
    lbd_compute(wrap(img, w, hgt, stride), kls, n, nullptr, fdesc);
    return OLF_OK;
}

// Lineextractor::operator() (src/LineExtractor.cc:31-67)
extern "C" int orc_line_extract(orc_line* h, const uint8_t* img, int w, int hgt, int stride, olf_keyline* kls, uint8_t* desc, int cap, int* n) {
    if (!h || !img || !n) return OLF_ERR_ARG;
    *n = 0;
    Image8 im = wrap(img, w, hgt, stride);
    std::vector<float> segs;
    lsd_detect(im, h->p, segs, &h->st);
    std::vector<olf_keyline> k;
    const double min_length = h->p.min_line_length * std::min(w, hgt);
    make_keylines(segs, w, hgt, min_length, k);
    const int nf = h->p.lsd_nfeatures;
    if ((int)k.size() > nf && nf != 0) {
        std::stable_sort(k.begin(), k.end(), [](const olf_keyline& a, const olf_keyline& b) { return a.response > b.response; });
        k.resize(nf);
        for (int i = 0; i < nf; i++) k[i].class_id = i;
    }
    if ((int)k.size() > cap) return OLF_ERR_CAPACITY;
    *n = (int)k.size();
    if (k.empty()) return OLF_OK;
    memcpy(kls, k.data(), k.size() * sizeof(olf_keyline));
    lbd_compute(im, k.data(), (int)k.size(), desc, nullptr);
    return OLF_OK;
}
