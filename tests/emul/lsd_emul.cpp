// TEST INFRASTRUCTURE.  Host emulation of the PERSISTENT-CLAIMS variant (orb_line_slam_b200/csrc/lsd_sticky.h) of
// orb_line_slam_b200/csrc/lsd_core.h: the SAME scan / verify_seed() / grow_step() / finalize_seed() / region_rect_a()
// source the kernels run is compiled for the CPU; every pass is replayed in a random order and the growths of a round
// are stepped in a random interleaving (emulating arbitrary GPU scheduling and partial visibility of claims).
// The result must equal the oracle's sequential LSD; the test also reports waves / rounds / work amplification.
#include "../../orb_line_slam_b200/csrc/lsd_sticky.h"
#include "../../oracle/cvprim.hpp"
#include "../../include/olf_abi.h"
#include <random>
#include <numeric>

using namespace olf::lsd;

// event_scan (the bookkeeping of k_lsd_scan / k_lsd_verify in line.cu): per seed a state byte (alive at the last evaluation) and a packed tile box of its
// region; from the second round of a wave a candidate is re-evaluated only where the previous round dirtied the tile of its pixel, the verify pass looks at
// the 4-byte tile box before the seed record and writes a death it discovers back into the state byte.  stats[9] = candidates NOT re-evaluated.
static unsigned emul_tbox_pack(const SeedRec3& r) {
    const int tx0 = r.x0 >> kTileShift, ty0 = r.y0 >> kTileShift, tx1 = r.x1 >> kTileShift, ty1 = r.y1 >> kTileShift;
    if ((tx1 | ty1) > 254) return 0xFFFFFFFEu;
    return (unsigned)tx0 | ((unsigned)ty0 << 8) | ((unsigned)tx1 << 16) | ((unsigned)ty1 << 24);
}
static bool emul_tbox_dirty(const Ctx3& C, unsigned round, unsigned tb) {
    SeedRec3 r; r.x0 = (unsigned short)((tb & 255) << kTileShift); r.y0 = (unsigned short)(((tb >> 8) & 255) << kTileShift);
    r.x1 = (unsigned short)(((tb >> 16) & 255) << kTileShift); r.y1 = (unsigned short)((tb >> 24) << kTileShift);
    return bbox_dirty(C, round, r);
}
static bool emul_pixel_tile_dirty(const Ctx3& C, unsigned round, int pix) {
    const int y = pix / C.W, x = pix - y * C.W, tx = x >> kTileShift, ty = y >> kTileShift;
    return (C.dirty[(round - 1) & 1][ty * C.tile_wpr + (tx >> 5)] >> (tx & 31)) & 1u;
}
// grow pass as k_lsd_grow<true> does it (emul_set_pipelined); g_stale_views counts the entries decided on an earlier view
static int g_pipelined = 0; static long long g_stale_views = 0, g_rechecks = 0;
extern "C" void emul_set_pipelined(int on) { g_pipelined = on; g_stale_views = 0; g_rechecks = 0; }
extern "C" long long emul_rechecks() { return g_rechecks; }
extern "C" long long emul_stale_views() { return g_stale_views; }
extern "C" int emul_lsd_detect2(const uint8_t* img, int w, int h, const olf_line_params* P, unsigned rng_seed, int first_wave, int defer, int exact_align, int event_scan,
                                float* segs, int cap, int* nseg, long long* stats /*[10]: waves, rounds, grown_px, final_px, regions, max_rounds_in_wave, carried, walked, live, skipped*/) {
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

    std::vector<PxRec> px(S);
    for (int q = 0; q < S; ++q) {
        PxRec r; r.claim[0] = ang[q] >= 0.f ? kClaimNone : 0ull; r.claim[1] = r.claim[0]; r.ang = ang[q]; r.cx = 0.f; r.cy = 0.f; r.binrev = 0;
        if (ang[q] >= 0.f) {
            const float2_t c = tab_acc[tab_index(dabc[q])];
            r.cx = c.x; r.cy = c.y;
            r.binrev = (unsigned)(n_bins - 1 - (int)(std::sqrt(n2[q] / 4.0) * bin_coef));
        }
        px[q] = r;
    }
    const unsigned pool_chunks = 1u << 21;
    std::vector<unsigned> pool((size_t)pool_chunks * kChunk);
    unsigned pool_ctr = 0;
    std::vector<SeedRec3> srec(n);
    std::vector<double> regang(n, 0.0);
    const int tiles_x = (W + 31) >> 5, tiles_y = (H + 31) >> 5, wpr = (tiles_x + 31) / 32;
    std::vector<unsigned> dirty0((size_t)tiles_y * wpr, 0), dirty1((size_t)tiles_y * wpr, 0);
    Ctx3 C;
    C.W = W; C.H = H; C.px = px.data(); C.dabc = dabc.data(); C.tab_seed = tab_seed.data();
    C.pool = pool.data(); C.pool_ctr = &pool_ctr; C.pool_chunks = pool_chunks;
    C.srec = srec.data(); C.regang = regang.data(); C.seed_pix = seeds.data(); C.seed_prio = prio.data(); C.prec = prec;
    C.dirty[0] = dirty0.data(); C.dirty[1] = dirty1.data(); C.tile_wpr = wpr;
    {
        const double margin = 0.1 * M_PI / 180.0;
        C.fast_align = (exact_align == 0) && (prec + margin < 80.0 * M_PI / 180.0);
        C.c_hi2 = (float)(std::cos(prec - margin) * std::cos(prec - margin));
        C.c_lo2 = (float)(std::cos(prec + margin) * std::cos(prec + margin));
    }
    std::vector<unsigned> final_pool(S); unsigned final_ctr = 0, nreg = 0;
    std::vector<LsdRegion> regs(S / std::max(min_reg_size, 1) + 16);
    FinalOut F; F.final_pool = final_pool.data(); F.final_ctr = &final_ctr; F.regs = regs.data(); F.nreg = &nreg;
    F.reg_cap = (unsigned)regs.size(); F.min_reg_size = min_reg_size;
    long long waves = 0, rounds = 0, grown = 0, final_px = 0, regions = 0, max_rw = 0, carried = 0, walked = 0, live_px = 0;
    unsigned round = 1;
    std::vector<int> order, wl0, wl2;
    std::vector<std::pair<int, bool>> wl1;
    std::vector<unsigned char> sstate(n, 0); std::vector<unsigned> tbox(n, kNull);
    long long skipped = 0;
    for (size_t wv = 0; wv + 1 < wave_start.size(); ++wv) {
        const int lo = wave_start[wv], hi = wave_start[wv + 1];
        if (lo == hi) continue;
        ++waves;
        pool_ctr = 0;
        C.stamp = (u64)(0xFFFFFFu - (unsigned)(wv + 1)) << 40;
        for (int i = lo; i < hi; ++i) { SeedRec3 z; z.head = kNull; z.cnt = 0; z.bhead = kNull; z.bcnt = 0; z.x0 = z.y0 = z.x1 = z.y1 = 0; z.pad0 = z.pad1 = 0; srec[i] = z; }
        std::fill(dirty0.begin(), dirty0.end(), 0u); std::fill(dirty1.begin(), dirty1.end(), 0u);
        long long rw = 0;
        bool first_round = true, force_next = false;
        for (;;) {
            ++rounds; ++rw;
            bool changed = false;
            const bool force = force_next;                     // the full verification of the wave failed: this round looks at everything (st->force in line.cu)
            force_next = false;
            std::fill(C.dirty[round & 1], C.dirty[round & 1] + (size_t)tiles_y * wpr, 0u);     // events of THIS round
            // pass 1: scan
            if (first_round) { order.resize(hi - lo); std::iota(order.begin(), order.end(), lo); }
            else order = wl0;
            std::shuffle(order.begin(), order.end(), rng);
            if (first_round) wl0.clear();
            wl1.clear(); wl2.clear();
            for (int i : order) {
                if (first_round) {
                    if (s3_final(C, seeds[i])) continue;
                    wl0.push_back(i);
                    tbox[i] = kNull;
                }
                bool alive;
                if (!event_scan) {
                    alive = s3_alive(C, seeds[i], prio[i]);
                    if (!alive && srec[i].cnt > 0) wl1.push_back({i, false});
                } else if (first_round || force || emul_pixel_tile_dirty(C, round, seeds[i])) {
                    alive = s3_alive(C, seeds[i], prio[i]);
                    sstate[i] = alive ? 1 : 0;
                    if (!alive && !first_round && srec[i].cnt > 0) wl1.push_back({i, false});       // died owning a region
                } else { alive = sstate[i] != 0; ++skipped; }
                if (alive && first_round && defer && s3_deferred(C, seeds[i], prio[i])) { changed = true; continue; }    // sits this round out
                if (alive) wl1.push_back({i, true});
            }
            // pass 2: verify
            std::shuffle(wl1.begin(), wl1.end(), rng);
            for (auto& e : wl1) {
                if (e.second && srec[e.first].cnt > 0) { live_px += srec[e.first].cnt; if (bbox_dirty(C, round, srec[e.first])) walked += srec[e.first].cnt; }
                if (event_scan && !force && e.second && tbox[e.first] < 0xFFFFFFFEu && !emul_tbox_dirty(C, round, tbox[e.first])) { ++carried; continue; }     // 4 bytes instead of the record
                const Verify3 v = s3_verify(C, round, e.first, e.second, &changed, force);
                if (v != kV3Carried) tbox[e.first] = kNull;
                if (e.second && v == kV3Dead) sstate[e.first] = 0;
                if (v == kV3Grow) wl2.push_back(e.first);
                else if (v == kV3Carried) ++carried;
            }
            // pass 3: grow, randomly interleaved
            {
                std::vector<GrowSt3> act;
                std::vector<View3> views;                    // pipelined walk: the view of the entry each thread expands next, sampled one turn earlier
                size_t next = 0;
                const size_t lanes = 1 + rng() % 48;
                while (next < wl2.size() || !act.empty()) {
                    while (act.size() < lanes && next < wl2.size()) {
                        GrowSt3 st; s3_begin(C, wl2[next++], st);
                        if (st.overflow) return -3;
                        act.push_back(st);
                        View3 v; v.valid = false; v.p = -1; views.push_back(v);
                    }
                    const size_t k = rng() % act.size();
                    bool more;
                    if (g_pipelined) {
                        // k_lsd_grow<true>: the loads of the entry AFTER the one expanded now leave first (if the queue holds it already), then the
                        // candidates of the current entry are decided on the view sampled a turn ago -- any number of other threads' turns ago
                        View3 nv; s3_peek(C, act[k], 1, nv);
                        unsigned claimed[8]; int nc = 0;
                        more = s3_step_view(C, round, act[k], &views[k], claimed, &nc);
                        for (int j = 0; j < nc; ++j) s3_view_patch(C, nv, claimed[j], act[k].mine);
                        views[k] = nv;
                        if (nv.valid) ++g_stale_views;
                    } else more = s3_step(C, round, act[k]);
                    if (!more) {
                        if (act[k].overflow) return -3;
                        grown += act[k].count;
                        s3_end(C, act[k]);
                        tbox[act[k].i] = emul_tbox_pack(srec[act[k].i]);
                        act[k] = act.back(); act.pop_back();
                        views[k] = views.back(); views.pop_back();
                    }
                }
            }
            first_round = false;
            if (!changed) {
                // before the wave is finalised every live region is verified regardless of dirty tiles
                bool clean = true;
                if (g_pipelined) {
                    // as mode 3 of the kernels: look only, release nothing; a failure sends the wave back to the rounds with everything dirty.
                    // (With views sampled turns ago the narrow race this pass exists for -- two parties each missing the other's claim -- does occur.)
                    for (int i : wl0) {
                        const SeedRec3 r = srec[i];
                        if (r.cnt <= 0) continue;
                        const u64 mine = key_of(C, prio[i]);
                        ListReader rd; rd.init(r.head);
                        for (int j = 0; j < r.cnt && clean; ++j) clean = ld_claim0(&C.px[rd.next(C.pool)]) == mine;
                        ListReader rb; rb.init(r.bhead);
                        for (int j = 0; j < r.bcnt && clean; ++j) clean = ld_claim0(&C.px[rb.next(C.pool)]) < mine;
                        if (!clean) break;
                    }
                    if (clean) break;
                    force_next = true; ++g_rechecks;
                } else {
                    for (int i : wl0) {
                        if (srec[i].cnt <= 0 || !s3_alive(C, seeds[i], prio[i])) { if (srec[i].cnt > 0) clean = false; continue; }
                        bool chg2 = false;
                        if (s3_verify(C, round, i, true, &chg2, true) != kV3Carried) { clean = false; wl2.push_back(i); }
                    }
                    if (clean) break;
                    return -5;                                      // cannot happen in a sequential emulation (steps are atomic)
                }
            }
            ++round;
            if (rw > 4000) return -4;
        }
        max_rw = std::max(max_rw, rw);
        for (int i : wl0) {
            const int c = srec[i].cnt;
            if (c > 0) { ++regions; final_px += c; }
            if (!s3_finalize(C, i, F)) return -3;
        }
        ++round;
    }
    // rectangle fit of the accepted regions
    struct Seg { u64 prio; float v[4]; };
    std::vector<Seg> out;
    for (unsigned r = 0; r < nreg; ++r) {
        const LsdRegion& R = regs[r];
        const unsigned* pix = final_pool.data() + R.off;
        const int c = R.count;
        RectA ra = region_rect_a(pix, c, dabc.data(), W, R.reg_angle, prec);
        const double dx = std::cos(ra.theta), dy = std::sin(ra.theta);
        double l_min = 0, l_max = 0;
        for (int k = 0; k < c; ++k) { const double l = region_proj(pix[k], W, ra.x, ra.y, dx, dy); if (l > l_max) l_max = l; else if (l < l_min) l_min = l; }
        double rr[4] = {ra.x + l_min * dx, ra.y + l_min * dy, ra.x + l_max * dx, ra.y + l_max * dy};
        Seg sg; sg.prio = R.prio;
        for (int k = 0; k < 4; ++k) { rr[k] += 0.5; if (scale != 1.0) rr[k] /= scale; sg.v[k] = (float)rr[k]; }
        out.push_back(sg);
    }
    std::sort(out.begin(), out.end(), [](const Seg& a, const Seg& b) { return a.prio < b.prio; });
    *nseg = (int)out.size();
    if ((int)out.size() > cap) return -3;
    for (size_t i = 0; i < out.size(); ++i) memcpy(segs + 4 * i, out[i].v, 16);
    stats[0] = waves; stats[1] = rounds; stats[2] = grown; stats[3] = final_px; stats[4] = regions; stats[5] = max_rw; stats[6] = carried; stats[7] = walked; stats[8] = live_px; stats[9] = skipped;
    return 0;
}
extern "C" int emul_lsd_detect(const uint8_t* img, int w, int h, const olf_line_params* P, unsigned rng_seed, int first_wave, int defer, int exact_align,
                               float* segs, int cap, int* nseg, long long* stats) {
    return emul_lsd_detect2(img, w, h, P, rng_seed, first_wave, defer, exact_align, 0, segs, cap, nseg, stats);
}
