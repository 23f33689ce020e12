// TEST INFRASTRUCTURE.  Host emulation of the parallel LSD region-growing scheme of
// orb_line_slam_b200/csrc/lsd_core.h: the SAME grow_seed()/region_rect_a() source is compiled for the CPU and the
// rounds are replayed sequentially with a random seed order per round (emulating arbitrary GPU scheduling).
// The result must equal the oracle's sequential LSD; the test also reports waves / rounds / work amplification.
#include "../../orb_line_slam_b200/csrc/lsd_core.h"
#include "../../oracle/cvprim.hpp"
#include "../../include/olf_abi.h"
#include <random>
#include <numeric>

using namespace olf::lsd;

extern "C" int emul_lsd_detect(const uint8_t* img, int w, int h, const olf_line_params* P, unsigned rng_seed, int first_wave,
                               float* segs, int cap, int* nseg, long long* stats /*[6]: waves, rounds, grown_px, final_px, regions, max_rounds_in_wave*/) {
    orc::Image8 im(w, h);
    memcpy(im.d.data(), img, (size_t)w * h);
    const double scale = P->lsd_scale;
    const double prec = M_PI * P->lsd_ang_th / 180.0;
    const double p = P->lsd_ang_th / 180.0;
    const double rho = P->lsd_quant / std::sin(prec);
    const int n_bins = P->lsd_n_bins;
    orc::Image8 scaled;
    if (scale != 1.0) {
        const double sigma = (scale < 1) ? (P->lsd_sigma_scale / scale) : P->lsd_sigma_scale;
        const unsigned hk = (unsigned)(std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0))));
        orc::Image8 g;
        orc::gaussian_blur_q8(im, orc::gauss_kernel_q8(1 + 2 * (int)hk, sigma), g);
        orc::resize_linear_exact(g, scale, scale, scaled);
    } else scaled = im;
    const int W = scaled.w, H = scaled.h, S = W * H;
    std::vector<float> ang(S, -1.f);
    std::vector<short2_t> dabc(S, short2_t{0, 0});
    std::vector<int> n2(S, 0);
    int n2max = 0;
    for (int y = 0; y < H - 1; ++y)
        for (int x = 0; x < W - 1; ++x) {
            const int DA = scaled.at(x + 1, y + 1) - scaled.at(x, y), BC = scaled.at(x + 1, y) - scaled.at(x, y + 1);
            const int gx = DA + BC, gy = DA - BC, q = y * W + x;
            dabc[q] = short2_t{(short)DA, (short)BC};
            n2[q] = gx * gx + gy * gy;
            const double norm = std::sqrt(n2[q] / 4.0);
            if (!(norm <= rho)) { ang[q] = orc::fast_atan2_deg((float)gx, (float)-gy); n2max = std::max(n2max, n2[q]); }
        }
    // trig tables from host libm (the product library does the same at handle creation)
    std::vector<float2_t> tab_seed((size_t)kTabDim * kTabDim), tab_acc((size_t)kTabDim * kTabDim);
    for (int DA = -255; DA <= 255; ++DA)
        for (int BC = -255; BC <= 255; ++BC) {
            const int gx = DA + BC, gy = DA - BC;
            const double a = orc::fast_atan2_deg((float)gx, (float)-gy) * kDegToRads;
            const size_t i = (size_t)(DA + 255) * kTabDim + (BC + 255);
            tab_seed[i] = float2_t{(float)std::cos(a), (float)std::sin(a)};
            tab_acc[i] = float2_t{(float)std::cos((double)(float)a), (float)std::sin((double)(float)a)};
        }
    const double max_grad = std::sqrt(n2max / 4.0);
    const double bin_coef = (max_grad > 0) ? (double)(n_bins - 1) / max_grad : 0;
    // seeds grouped by bin (descending); order inside a bin is arbitrary on the GPU -> shuffle
    std::mt19937 rng(rng_seed);
    std::vector<std::vector<int>> by_bin(n_bins);
    for (int q = 0; q < S; ++q) if (ang[q] >= 0.f) by_bin[(int)(std::sqrt(n2[q] / 4.0) * bin_coef)].push_back(q);
    std::vector<int> seeds; std::vector<u64> prio; std::vector<int> bin_end;   // bin_end: cumulative count after each bin (desc)
    for (int b = n_bins - 1; b >= 0; --b) {
        std::shuffle(by_bin[b].begin(), by_bin[b].end(), rng);
        for (int q : by_bin[b]) { seeds.push_back(q); prio.push_back(make_prio(n_bins - 1 - b, q)); }
        bin_end.push_back((int)seeds.size());
    }
    const int n = (int)seeds.size();
    // waves: whole bins, cumulative targets doubling
    std::vector<int> wave_start{0};
    { long long target = first_wave; for (int e : bin_end) { if (e - wave_start.back() >= target && e < n) { wave_start.push_back(e); target *= 2; } } wave_start.push_back(n); }
    const double logNT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const int min_reg_size = (int)(unsigned)(-logNT / std::log10(p));

    std::vector<u64> claim0(S, kClaimNone), claim1(S, kClaimNone);
    const unsigned pool_chunks = 1u << 22;
    std::vector<unsigned> pool0((size_t)pool_chunks * kChunk), pool1((size_t)pool_chunks * kChunk);
    unsigned ctr0 = 0, ctr1 = 0;
    GrowArgs A;
    A.W = W; A.H = H; A.ang = ang.data(); A.dabc = dabc.data(); A.tab_seed = tab_seed.data(); A.tab_acc = tab_acc.data();
    A.claim[0] = claim0.data(); A.claim[1] = claim1.data(); A.pool[0] = pool0.data(); A.pool[1] = pool1.data();
    A.pool_ctr[0] = &ctr0; A.pool_ctr[1] = &ctr1; A.pool_chunks = pool_chunks; A.prec = prec;
    std::vector<unsigned> head[2] = {std::vector<unsigned>(n, kNull), std::vector<unsigned>(n, kNull)};
    std::vector<int> cnt[2] = {std::vector<int>(n, 0), std::vector<int>(n, 0)};
    std::vector<double> regang(n, 0.0);
    struct Seg { u64 prio; float v[4]; };
    std::vector<Seg> out;
    long long waves = 0, rounds = 0, grown = 0, final_px = 0, regions = 0, max_rw = 0;
    unsigned round = 1;
    std::vector<int> order;
    for (size_t wv = 0; wv + 1 < wave_start.size(); ++wv) {
        const int lo = wave_start[wv], hi = wave_start[wv + 1];
        if (lo == hi) continue;
        ++waves;
        for (int i = lo; i < hi; ++i) { cnt[0][i] = cnt[1][i] = 0; }
        long long rw = 0;
        for (;;) {
            ++rounds; ++rw;
            *A.pool_ctr[round & 1] = 0;
            order.resize(hi - lo); std::iota(order.begin(), order.end(), lo);
            std::shuffle(order.begin(), order.end(), rng);
            bool changed = false;
            for (int i : order) {
                GrowResult r = grow_seed(A, round, seeds[i], prio[i], head[(round - 1) & 1][i], cnt[(round - 1) & 1][i]);
                if (r.overflow) return -3;
                head[round & 1][i] = r.head; cnt[round & 1][i] = r.count; regang[i] = r.reg_angle;
                if (!r.same_as_prev) changed = true;
                grown += r.count;
            }
            if (!changed) break;
            ++round;
        }
        max_rw = std::max(max_rw, rw);
        // finalise: stamp 0 in both claim arrays, emit accepted regions
        std::vector<unsigned> pix;
        for (int i = lo; i < hi; ++i) {
            const int c = cnt[round & 1][i];
            if (c == 0) continue;
            ++regions; final_px += c;
            pix.resize(c);
            ListReader rd; rd.init(A.pool[round & 1], head[round & 1][i]);
            for (int k = 0; k < c; ++k) { pix[k] = rd.next(); claim0[pix[k]] = prio[i]; claim1[pix[k]] = prio[i]; }
            if (c < min_reg_size) continue;
            RectA ra = region_rect_a(pix.data(), c, dabc.data(), W, regang[i], prec);
            const double dx = std::cos(ra.theta), dy = std::sin(ra.theta);
            double l_min = 0, l_max = 0;
            for (int k = 0; k < c; ++k) { const double l = region_proj(pix[k], W, ra.x, ra.y, dx, dy); if (l > l_max) l_max = l; else if (l < l_min) l_min = l; }
            double r[4] = {ra.x + l_min * dx, ra.y + l_min * dy, ra.x + l_max * dx, ra.y + l_max * dy};
            Seg s; s.prio = prio[i];
            for (int k = 0; k < 4; ++k) { r[k] += 0.5; if (scale != 1.0) r[k] /= scale; s.v[k] = (float)r[k]; }
            out.push_back(s);
        }
        ++round;
    }
    std::sort(out.begin(), out.end(), [](const Seg& a, const Seg& b) { return a.prio < b.prio; });
    *nseg = (int)out.size();
    if ((int)out.size() > cap) return -3;
    for (size_t i = 0; i < out.size(); ++i) memcpy(segs + 4 * i, out[i].v, 16);
    stats[0] = waves; stats[1] = rounds; stats[2] = grown; stats[3] = final_px; stats[4] = regions; stats[5] = max_rw;
    return 0;
}
