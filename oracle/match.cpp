// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Not part of the product path.
//
// CPU restatement of the matching half of the reference front end:
//   cv::BFMatcher(NORM_HAMMING).knnMatch + matchNNR + match        src/LineMatcher.cpp:42-132
//   Frame::ComputeStereoMatches                                      src/Frame.cc:702-876
//   Frame::ComputeStereoMatches_Lines + matchGrid(lines) + GridStructure + LineIterator
//                                        src/Frame.cc:878-1048, src/LineMatcher.cpp:220-299, src/gridStructure.cpp, src/LineIterator.cpp
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea     src/Frame.cc:334-349, 517-582
//   ORBmatcher::SearchByProjection (last frame, local map)          src/ORBmatcher.cc:47-139, 1474-1618, 1749-1790
// Canonical choices: candidates of matchGrid iterated in ascending index (C.2); no FMA (C.4).
#include "oracle.h"
#include "cvprim.hpp"
#include <climits>
#include <limits>
#include <set>
#include <map>

using namespace orc;

// ---- brute force kNN(2), ties -> lowest train index --------------------------------------------------
extern "C" int orc_knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx0, int* dist0, int* idx1, int* dist1) {
    if (n1 < 0 || n2 < 0) return OLF_ERR_ARG;
    for (int i = 0; i < n1; ++i) {
        int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
        for (int j = 0; j < n2; ++j) {
            int d = hamming256(d1 + (size_t)i * 32, d2 + (size_t)j * 32);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
            else if (d < b1) { b1 = d; i1 = j; }
        }
        idx0[i] = i0; dist0[i] = b0; idx1[i] = i1; dist1[i] = b1;
    }
    return OLF_OK;
}

// matchNNR (src/LineMatcher.cpp:42-62).  n2 < 2 is UB in the reference (matches_[idx][1]); canonical: no match.
extern "C" int orc_match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12, int* nmatches) {
    std::vector<int> i0(n1), s0(n1), i1(n1), s1(n1);
    orc_knn2_hamming(d1, n1, d2, n2, i0.data(), s0.data(), i1.data(), s1.data());
    int matches = 0;
    for (int i = 0; i < n1; ++i) {
        m12[i] = -1;
        if (n2 < 2) continue;
        if ((float)s0[i] < (float)s1[i] * nnr) { m12[i] = i0[i]; matches++; }
    }
    *nmatches = matches;
    return OLF_OK;
}

// match(desc1, desc2, nnr, matches_12) (src/LineMatcher.cpp:104-132)
extern "C" int orc_match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int* m12, int* nmatches) {
    int matches;
    orc_match_nnr(d1, n1, d2, n2, nnr, m12, &matches);
    if (best_lr) {
        std::vector<int> m21(n2);
        int tmp;
        orc_match_nnr(d2, n2, d1, n1, nnr, m21.data(), &tmp);
        for (int i1 = 0; i1 < n1; ++i1) {
            int& i2 = m12[i1];
            if (i2 >= 0 && m21[i2] != i1) { i2 = -1; matches--; }
        }
    }
    *nmatches = matches;
    return OLF_OK;
}

// ---- Frame::ComputeStereoMatches (src/Frame.cc:702-876) ----------------------------------------------
static inline uint8_t px_reflect(const Image8& im, int x, int y) {
    return im.at(reflect101(x, im.w), reflect101(y, im.h));      // reference reads its 19-px REFLECT_101 border
}
extern "C" int orc_stereo_points(orc_orb* left, orc_orb* right,
                                 const olf_keypoint* kl, const uint8_t* dl, int N,
                                 const olf_keypoint* kr, const uint8_t* dr, int Nr,
                                 float mbf, float fx, float* uRight, float* depth) {
    if (!left || !right) return OLF_ERR_ARG;
    for (int i = 0; i < N; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (OLF_TH_HIGH + OLF_TH_LOW) / 2;
    const Image8* l0 = orc_orb_level_image(left, 0);
    const int nRows = l0->h;
    const float* scales = orc_orb_scales(left);
    const float* inv_scales = orc_orb_inv_scales(left);
    std::vector<std::vector<size_t>> rows(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * scales[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);
    }
    const float mb = mbf / fx;
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const olf_keypoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const size_t rowi = (size_t)vL;
        if (rowi >= (size_t)nRows) continue;
        const std::vector<size_t>& cand = rows[rowi];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = OLF_TH_HIGH;
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < cand.size(); iC++) {
            const size_t iR = cand[iC];
            const olf_keypoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dl + (size_t)iL * 32, dr + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = inv_scales[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const Image8& IL = *orc_orb_level_image(left, kpL.octave);
            const Image8& IR = *orc_orb_level_image(right, kpL.octave);
            const int cxL = (int)scaleduL, cyL = (int)scaledvL, cxR = (int)scaleduR0;
            int bestDistS = INT_MAX, bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= IR.w) continue;
            const int cL = px_reflect(IL, cxL, cyL);
            for (int incR = -L; incR <= +L; incR++) {
                const int cR = px_reflect(IR, cxR + incR, cyL);
                float dist = 0;      // L1 norm of integer-valued float differences: exact in any order
                int acc = 0;
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        int a = px_reflect(IL, cxL + dx, cyL + dy) - cL;
                        int b = px_reflect(IR, cxR + incR + dx, cyL + dy) - cR;
                        acc += std::abs(a - b);
                    }
                dist = (float)acc;
                if (dist < (float)bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scales[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = (float)(uL - 0.01); }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::make_pair(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return OLF_OK;           // reference indexes an empty vector here (UB); canonical: nothing to prune
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if ((float)vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
    }
    return OLF_OK;
}

// ---- LineIterator (src/LineIterator.cpp:34-77) + getLineCoords (src/gridStructure.cpp:33-41) -----------
static void line_coords(double x1, double y1, double x2, double y2, std::vector<std::pair<int, int>>& out) {
    out.clear();
    const bool steep = std::abs(y2 - y1) > std::abs(x2 - x1);
    if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
    if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
    const double dx = x2 - x1, dy = std::abs(y2 - y1);
    double error = dx / 2.0;
    const int ystep = (y1 < y2) ? 1 : -1;
    int x = (int)x1, y = (int)y1;
    const int maxX = (int)x2;
    while (x <= maxX) {
        out.push_back(steep ? std::make_pair(y, x) : std::make_pair(x, y));
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
        x++;
    }
}

// Frame::ComputeStereoMatches_Lines (src/Frame.cc:878-1000) with matchGrid(lines) (src/LineMatcher.cpp:220-299)
extern "C" int orc_stereo_lines(const olf_keyline* kl, const uint8_t* dl, int n1,
                                const olf_keyline* kr, const uint8_t* dr, int n2,
                                int img_w, int img_h, const olf_line_match_params* P,
                                int* matches12, float* disp, double* le) {
    if (!P) return OLF_ERR_ARG;
    for (int i = 0; i < n1; ++i) { matches12[i] = -1; disp[2 * i] = -1; disp[2 * i + 1] = -1; le[3 * i] = le[3 * i + 1] = le[3 * i + 2] = 0; }
    if (n1 == 0 || n2 == 0) return OLF_OK;
    const double inv_width = OLF_GRID_COLS / (double)img_w;
    const double inv_height = OLF_GRID_ROWS / (double)img_h;
    const int cols = OLF_GRID_COLS, rows = OLF_GRID_ROWS;
    // coords of left lines: double -> int truncation on conversion to line_2d
    std::vector<int> c1((size_t)n1 * 4);
    for (int i = 0; i < n1; ++i) {
        c1[4 * i] = (int)(kl[i].startPointX * inv_width); c1[4 * i + 1] = (int)(kl[i].startPointY * inv_height);
        c1[4 * i + 2] = (int)(kl[i].endPointX * inv_width); c1[4 * i + 3] = (int)(kl[i].endPointY * inv_height);
    }
    std::vector<std::vector<int>> grid((size_t)cols * rows);       // grid[x*rows + y]
    std::vector<std::pair<double, double>> dir2(n2);
    std::vector<std::pair<int, int>> lc;
    for (int idx = 0; idx < n2; ++idx) {
        const olf_keyline& k = kr[idx];
        double vx = (k.endPointX - k.startPointX) * inv_width, vy = (k.endPointY - k.startPointY) * inv_height;
        double mag = std::sqrt(vx * vx + vy * vy);
        dir2[idx] = std::make_pair(vx / mag, vy / mag);
        line_coords(k.startPointX * inv_width, k.startPointY * inv_height, k.endPointX * inv_width, k.endPointY * inv_height, lc);
        for (auto& p : lc) if (p.first >= 0 && p.first < cols && p.second >= 0 && p.second < rows) grid[(size_t)p.first * rows + p.second].push_back(idx);
    }
    const int ww0 = P->matching_s_ws, ww1 = 0, wh0 = 0, wh1 = 0;
    std::vector<int> m21, distances;
    if (P->best_lr_matches) { m21.assign(n2, -1); distances.assign(n2, INT_MAX); }
    auto grid_get = [&](int x, int y, std::set<int>& out) {
        int min_x = std::max(0, x - ww0), max_x = std::min(cols, x + ww1 + 1);
        int min_y = std::max(0, y - wh0), max_y = std::min(rows, y + wh1 + 1);
        for (int x_ = min_x; x_ < max_x; ++x_)
            for (int y_ = min_y; y_ < max_y; ++y_)
                out.insert(grid[(size_t)x_ * rows + y_].begin(), grid[(size_t)x_ * rows + y_].end());
    };
    for (int i1 = 0; i1 < n1; ++i1) {
        int best_d = INT_MAX, best_d2 = INT_MAX, best_idx = -1;
        const int spx = c1[4 * i1], spy = c1[4 * i1 + 1], epx = c1[4 * i1 + 2], epy = c1[4 * i1 + 3];
        double vx = epx - spx, vy = epy - spy;
        double mag = std::sqrt(vx * vx + vy * vy);
        vx /= mag; vy /= mag;
        std::set<int> cand;                       // canonical: ascending index
        grid_get(spx, spy, cand);
        grid_get(epx, epy, cand);
        if (cand.empty()) continue;
        for (int i2 : cand) {
            if (i2 < 0 || i2 >= n2) continue;
            if (std::abs(vx * dir2[i2].first + vy * dir2[i2].second) < P->line_sim_th) continue;
            const int d = hamming256(dl + (size_t)i1 * 32, dr + (size_t)i2 * 32);
            if (P->best_lr_matches) {
                if (d < distances[i2]) { distances[i2] = d; m21[i2] = i1; }
                else continue;
            }
            if (d < best_d) { best_d2 = best_d; best_d = d; best_idx = i2; }
            else if (d < best_d2) best_d2 = d;
        }
        if (best_d < best_d2 * P->min_ratio_12_l) matches12[i1] = best_idx;
    }
    if (P->best_lr_matches)
        for (int i1 = 0; i1 < n1; ++i1) { int& i2 = matches12[i1]; if (i2 >= 0 && m21[i2] != i1) i2 = -1; }

    // geometric post-filter (src/Frame.cc:934-958, 1002-1048)
    for (int i1 = 0; i1 < n1; ++i1) {
        const int i2 = matches12[i1];
        if (i2 < 0) continue;
        double sp_l[3] = {kl[i1].startPointX, kl[i1].startPointY, 1.0};
        double ep_l[3] = {kl[i1].endPointX, kl[i1].endPointY, 1.0};
        double le_l[3] = {sp_l[1] * ep_l[2] - sp_l[2] * ep_l[1], sp_l[2] * ep_l[0] - sp_l[0] * ep_l[2], sp_l[0] * ep_l[1] - sp_l[1] * ep_l[0]};
        double nrm = std::sqrt(le_l[0] * le_l[0] + le_l[1] * le_l[1]);
        le_l[0] = le_l[0] / nrm; le_l[1] = le_l[1] / nrm; le_l[2] = le_l[2] / nrm;
        double sp_r[3] = {kr[i2].startPointX, kr[i2].startPointY, 1.0};
        double ep_r[3] = {kr[i2].endPointX, kr[i2].endPointY, 1.0};
        // lineSegmentOverlapStereo(sp_l(1), ep_l(1), sp_r(1), ep_r(1))
        double overlap = 1.f;
        {
            const double spl_obs = sp_l[1], epl_obs = ep_l[1], spl_proj = sp_r[1], epl_proj = ep_r[1];
            if (std::fabs(epl_obs - spl_obs) > P->line_horiz_th) {
                double sln = std::min(spl_obs, epl_obs), eln = std::max(spl_obs, epl_obs);
                double spn = std::min(spl_proj, epl_proj), epn = std::max(spl_proj, epl_proj);
                double length = eln - spn;
                if ((epn < sln) || (spn > eln)) overlap = 0.f;
                else {
                    if ((epn > eln) && (spn < sln)) overlap = eln - sln;
                    else overlap = std::min(eln, epn) - std::max(sln, spn);
                }
                if (length > 0.01f) overlap = overlap / length; else overlap = 0.f;
                if (overlap > 1.f) overlap = 1.f;
            }
        }
        // Eigen comma-initialisers evaluate all three expressions before assignment (sp_r updated as a whole),
        // and the second statement then reads the UPDATED sp_r (src/Frame.cc:947-948).
        {
            double nx = (sp_r[0] * (sp_l[1] - ep_r[1]) + ep_r[0] * (sp_r[1] - sp_l[1])) / (sp_r[1] - ep_r[1]);
            sp_r[0] = nx; sp_r[1] = sp_l[1]; sp_r[2] = 1.0;
            double mx = (sp_r[0] * (ep_l[1] - ep_r[1]) + ep_r[0] * (sp_r[1] - ep_l[1])) / (sp_r[1] - ep_r[1]);
            ep_r[0] = mx; ep_r[1] = ep_l[1]; ep_r[2] = 1.0;
        }
        double disp_s = sp_l[0] - sp_r[0], disp_e = ep_l[0] - ep_r[0];
        if (std::min(disp_s, disp_e) / std::max(disp_s, disp_e) < P->ls_min_disp_ratio) { disp_s = -1.0; disp_e = -1.0; }
        if (disp_s >= P->min_disp && disp_e >= P->min_disp && std::abs(sp_l[1] - ep_l[1]) > P->line_horiz_th &&
            std::abs(sp_r[1] - ep_r[1]) > P->line_horiz_th && overlap > P->stereo_overlap_th) {
            disp[2 * i1] = (float)disp_s; disp[2 * i1 + 1] = (float)disp_e;
            le[3 * i1] = le_l[0]; le[3 * i1 + 1] = le_l[1]; le[3 * i1 + 2] = le_l[2];
        }
    }
    return OLF_OK;
}

// ---- Frame grid (src/Frame.cc:334-349, 517-582) --------------------------------------------------------
struct FrameGrid {
    std::vector<std::vector<int>> cell;     // [ix*48+iy]
    float minX, minY, invW, invH;
};
static void build_grid(const olf_keypoint* kps, int N, const olf_camera& cam, FrameGrid& g) {
    g.cell.assign((size_t)OLF_GRID_COLS * OLF_GRID_ROWS, {});
    g.minX = cam.min_x; g.minY = cam.min_y;
    g.invW = (float)OLF_GRID_COLS / (cam.max_x - cam.min_x);
    g.invH = (float)OLF_GRID_ROWS / (cam.max_y - cam.min_y);
    for (int i = 0; i < N; i++) {
        int posX = (int)roundf((kps[i].x - g.minX) * g.invW);
        int posY = (int)roundf((kps[i].y - g.minY) * g.invH);
        if (posX < 0 || posX >= OLF_GRID_COLS || posY < 0 || posY >= OLF_GRID_ROWS) continue;
        g.cell[(size_t)posX * OLF_GRID_ROWS + posY].push_back(i);
    }
}
static void features_in_area(const FrameGrid& g, const olf_keypoint* kps, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)floorf((x - g.minX - r) * g.invW));
    if (nMinCellX >= OLF_GRID_COLS) return;
    const int nMaxCellX = std::min((int)OLF_GRID_COLS - 1, (int)ceilf((x - g.minX + r) * g.invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)floorf((y - g.minY - r) * g.invH));
    if (nMinCellY >= OLF_GRID_ROWS) return;
    const int nMaxCellY = std::min((int)OLF_GRID_ROWS - 1, (int)ceilf((y - g.minY + r) * g.invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<int>& vCell = g.cell[(size_t)ix * OLF_GRID_ROWS + iy];
            for (int j : vCell) {
                const olf_keypoint& kp = kps[j];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) out.push_back(j);
            }
        }
}

static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {   // src/ORBmatcher.cc:1749-1790
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static inline void mat3_mul_vec(const float* R, const float* v, float* o) {      // cv::Mat float gemm: row . vec, sequential
    for (int r = 0; r < 3; ++r) o[r] = R[3 * r] * v[0] + R[3 * r + 1] * v[1] + R[3 * r + 2] * v[2];
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono, match12) (src/ORBmatcher.cc:1474-1618)
extern "C" int orc_search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches_out) {
    if (!a) return OLF_ERR_ARG;
    FrameGrid g;
    build_grid(a->cur_kps, a->n_cur, a->cam, g);
    int nmatches = 0;
    std::vector<int> rotHist[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    // twc = -Rcw^T * tcw ; tlc = Rlw*twc + tlw
    float twc[3], tlc[3];
    for (int r = 0; r < 3; ++r) twc[r] = -(a->Rcw[r] * a->tcw[0] + a->Rcw[3 + r] * a->tcw[1] + a->Rcw[6 + r] * a->tcw[2]);
    mat3_mul_vec(a->Rlw, twc, tlc);
    for (int r = 0; r < 3; ++r) tlc[r] = tlc[r] + a->tlw[r];
    const float mb = a->cam.bf / a->cam.fx;
    const bool bForward = tlc[2] > mb && !a->mono;
    const bool bBackward = -tlc[2] > mb && !a->mono;
    for (int j = 0; j < a->n_cur; ++j) cur_point[j] = -1;
    std::vector<uint8_t> blocked(a->n_cur, 0);             // CurrentFrame.mvpMapPoints[i2] && Observations()>0
    if (a->cur_occupied) for (int j = 0; j < a->n_cur; ++j) blocked[j] = a->cur_occupied[j];
    std::vector<int> cand;
    for (int i = 0; i < a->n_last; i++) {
        assigned_cur[i] = -1;
        if (!a->last_has_point[i]) continue;
        float x3Dc[3];
        mat3_mul_vec(a->Rcw, a->last_world_pos + 3 * i, x3Dc);
        for (int r = 0; r < 3; ++r) x3Dc[r] = x3Dc[r] + a->tcw[r];
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / x3Dc[2]);
        if (invzc < 0) continue;
        float u = a->cam.fx * xc * invzc + a->cam.cx;
        float v = a->cam.fy * yc * invzc + a->cam.cy;
        if (u < a->cam.min_x || u > a->cam.max_x) continue;
        if (v < a->cam.min_y || v > a->cam.max_y) continue;
        int nLastOctave = a->last_kps[i].octave;
        float radius = a->th * a->scale_factors[nLastOctave];
        if (bForward) features_in_area(g, a->cur_kps, u, v, radius, nLastOctave, -1, cand);
        else if (bBackward) features_in_area(g, a->cur_kps, u, v, radius, 0, nLastOctave, cand);
        else features_in_area(g, a->cur_kps, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (blocked[i2]) continue;
            if (a->cur_u_right[i2] > 0) {
                const float ur = u - a->cam.bf * invzc;
                const float er = fabsf(ur - a->cur_u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = hamming256(a->last_point_desc + (size_t)i * 32, a->cur_desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= OLF_TH_HIGH) {
            cur_point[bestIdx2] = i;
            if (a->last_point_observed[i]) blocked[bestIdx2] = 1;
            assigned_cur[i] = bestIdx2;
            nmatches++;
            if (a->check_orientation) {
                float rot = a->last_kps[i].angle - a->cur_kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == OLF_HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (a->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, OLF_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < OLF_HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) { cur_point[rotHist[i][j]] = -1; nmatches--; }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}

// ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:47-131)
extern "C" int orc_search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur, int* nmatches_out) {
    if (!a) return OLF_ERR_ARG;
    FrameGrid g;
    build_grid(a->cur_kps, a->n_cur, a->cam, g);
    int nmatches = 0;
    const bool bFactor = a->th != 1.0;
    std::vector<uint8_t> blocked(a->n_cur, 0);
    for (int j = 0; j < a->n_cur; ++j) blocked[j] = a->cur_occupied ? a->cur_occupied[j] : 0;
    std::vector<int> cand;
    for (int iMP = 0; iMP < a->n_points; iMP++) {
        assigned_cur[iMP] = -1;
        const int nPredictedLevel = a->pred_level[iMP];
        float r = (a->view_cos[iMP] > 0.998) ? 2.5f : 4.0f;
        if (bFactor) r *= a->th;
        features_in_area(g, a->cur_kps, a->proj_x[iMP], a->proj_y[iMP], r * a->scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (blocked[idx]) continue;
            if (a->cur_u_right[idx] > 0) {
                const float er = fabsf(a->proj_xr[iMP] - a->cur_u_right[idx]);
                if (er > r * a->scale_factors[nPredictedLevel]) continue;
            }
            const int dist = hamming256(a->point_desc + (size_t)iMP * 32, a->cur_desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = a->cur_kps[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = a->cur_kps[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= OLF_TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > a->nn_ratio * bestDist2) continue;
            assigned_cur[iMP] = bestIdx;
            if (a->point_observed[iMP]) blocked[bestIdx] = 1;
            nmatches++;
        }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:254-322; src/MapLine.cc:257-322 is the same on LBD descriptors):
// full distance matrix, per row the sorted distances' element [0.5*(N-1)], the first row with the least such median wins.
extern "C" int orc_distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best) {
    for (int g = 0; g < n_groups; ++g) {
        const int b = group_begin[g], N = group_begin[g + 1] - b;
        if (N <= 0) { best[g] = -1; continue; }
        std::vector<std::vector<float>> D(N, std::vector<float>(N, 0.f));
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++) { const int d = hamming256(desc + (size_t)(b + i) * 32, desc + (size_t)(b + j) * 32); D[i][j] = (float)d; D[j][i] = (float)d; }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            std::vector<int> vDists(D[i].begin(), D[i].end());
            std::sort(vDists.begin(), vDists.end());
            const int median = vDists[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best[g] = BestIdx;
    }
    return 0;
}

// ---- the remaining ORBmatcher overloads ----------------------------------------------------------------
// The candidate loop shared by Fuse (src/ORBmatcher.cc:894-974, 1053-1098), SearchByProjection(KF, Scw, ..) (:344-400), SearchBySim3
// (:1193-1226, :1273-1306) and SearchByProjection(Frame, KF, found, th, ORBdist) (:1680-1721): KeyFrame::GetFeaturesInArea
// (src/KeyFrame.cc:747-786) in cell-major order, octave window, optional chi-square gate, first minimum, optional blocking.
extern "C" int orc_window_search(const olf_window_search_args* a, int* best_idx, int* best_dist) {
    if (!a || !best_idx || !best_dist) return OLF_ERR_ARG;
    FrameGrid g;
    build_grid(a->kps, a->n, a->cam, g);
    std::vector<uint8_t> blocked(std::max(a->n, 1), 0);
    if (a->blocked) for (int j = 0; j < a->n; ++j) blocked[j] = a->blocked[j];
    std::vector<int> cand;
    for (int i = 0; i < a->n_queries; ++i) {
        best_idx[i] = -1; best_dist[i] = 256;
        const float u = a->u[i], v = a->v[i];
        features_in_area(g, a->kps, u, v, a->radius[i], -1, -1, cand);            // the KeyFrame overload has no level arguments (:747)
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (blocked[idx]) continue;
            const olf_keypoint& kp = a->kps[idx];
            if (kp.octave < a->min_level[i] || kp.octave > a->max_level[i]) continue;
            if (a->chi2_check) {                                                 // :916-940
                const float ex = u - kp.x, ey = v - kp.y;
                if (a->u_right[idx] >= 0) {
                    const float er = a->ur[i] - a->u_right[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * a->inv_level_sigma2[kp.octave] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * a->inv_level_sigma2[kp.octave] > 5.99) continue;
                }
            }
            const int dist = hamming256(a->qdesc + (size_t)i * 32, a->desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestIdx >= 0 && bestDist <= a->max_dist) {
            best_idx[i] = bestIdx; best_dist[i] = bestDist;
            if (a->sequential_blocking) blocked[bestIdx] = 1;
        }
    }
    return OLF_OK;
}

// ORBmatcher::CheckDistEpipolarLine (src/ORBmatcher.cc:142-159)
static bool check_dist_epipolar_line(const olf_keypoint& kp1, const olf_keypoint& kp2, const float* F12, const float* level_sigma2) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * level_sigma2[kp2.octave];
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:659-825).  vbMatched2 is never set in the reference (:679 is its only
// write), so a feature of KF2 can serve several features of KF1.
extern "C" int orc_search_for_triangulation(const olf_triangulation_args* a, int* matches12, int* nmatches_out) {
    if (!a || !matches12 || !nmatches_out) return OLF_ERR_ARG;
    int nmatches = 0;
    for (int i = 0; i < a->n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    int i1n = 0, i2n = 0;
    while (i1n < a->fv1_n_nodes && i2n < a->fv2_n_nodes) {
        if (a->fv1_node[i1n] == a->fv2_node[i2n]) {
            for (int p = a->fv1_begin[i1n]; p < a->fv1_begin[i1n + 1]; ++p) {
                const int idx1 = a->fv1_index[p];
                if (a->skip1[idx1]) continue;
                const bool bStereo1 = a->u_right1[idx1] >= 0;
                if (a->only_stereo && !bStereo1) continue;
                const olf_keypoint& kp1 = a->kps1[idx1];
                int bestDist = OLF_TH_LOW, bestIdx2 = -1;
                for (int q = a->fv2_begin[i2n]; q < a->fv2_begin[i2n + 1]; ++q) {
                    const int idx2 = a->fv2_index[q];
                    if (a->skip2[idx2]) continue;
                    const bool bStereo2 = a->u_right2[idx2] >= 0;
                    if (a->only_stereo && !bStereo2) continue;
                    const int dist = hamming256(a->desc1 + (size_t)idx1 * 32, a->desc2 + (size_t)idx2 * 32);
                    if (dist > OLF_TH_LOW || dist > bestDist) continue;
                    const olf_keypoint& kp2 = a->kps2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = a->ex - kp2.x, distey = a->ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * a->scale_factors2[kp2.octave]) continue;
                    }
                    if (check_dist_epipolar_line(kp1, kp2, a->F12, a->level_sigma2_2)) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    nmatches++;
                    if (a->check_orientation) {
                        float rot = kp1.angle - a->kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == OLF_HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
            ++i1n; ++i2n;
        } else if (a->fv1_node[i1n] < a->fv2_node[i2n]) {
            while (i1n < a->fv1_n_nodes && a->fv1_node[i1n] < a->fv2_node[i2n]) ++i1n;      // lower_bound (:785)
        } else {
            while (i2n < a->fv2_n_nodes && a->fv2_node[i2n] < a->fv1_node[i1n]) ++i2n;
        }
    }
    if (a->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, OLF_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
        }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522)
extern "C" int orc_search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2,
                                             const olf_camera* cam, float* prev_matched, int window_size, float nn_ratio, int check_orientation,
                                             int* vnMatches12, int* nmatches_out) {
    if (!cam || !vnMatches12 || !nmatches_out) return OLF_ERR_ARG;
    FrameGrid g;
    build_grid(kps2, n2, *cam, g);
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) vnMatches12[i] = -1;
    std::vector<int> rotHist[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    std::vector<int> vMatchedDistance(std::max(n2, 1), INT_MAX), vnMatches21(std::max(n2, 1), -1), cand;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;
        features_in_area(g, kps2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, cand);
        if (cand.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = hamming256(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= OLF_TH_LOW) {
            if (bestDist < (float)bestDist2 * nn_ratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_orientation) {
                    float rot = kps1[i1].angle - kps2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == OLF_HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, OLF_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) { prev_matched[2 * i1] = kps2[vnMatches12[i1]].x; prev_matched[2 * i1 + 1] = kps2[vnMatches12[i1]].y; }
    *nmatches_out = nmatches;
    return OLF_OK;
}
