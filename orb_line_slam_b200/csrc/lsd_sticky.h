// Parallel LSD region growing with PERSISTENT claims (the executable specification the kernels in line.cu follow and the
// host emulation in tests/emul replays).
//
// The reference visits seeds in priority order (gradient bin descending, then row-major) and each region marks its pixels
// USED for every later seed.  That is the least fixed point of
//     region(s) = dead                                 if s is inside region(q) for some q of higher priority
//               = grow(s, used = U_{q<s} region(q))    otherwise
// which is well-founded on the priority order, so ANY iteration of that operator that reaches a fixed point reaches the
// sequential result (induction on priority: the highest-priority wrong seed would have seen only correct claims).  Waves
// are priority prefixes (whole bins): a wave is finalised before lower-priority seeds are considered.  One round = three
// data-parallel passes, one thread per seed: scan (dead or alive?), verify (is my last growth still exact?), grow.
// A claim stays valid for the whole wave, so an unchanged region costs O(1) per round:
//   * one claim word per pixel, [stamp:24 | prio:40] with atomicMin; stamp 0 = final, stamp S_w = 0xFFFFFF - wave = a claim
//     of the current wave, anything else = free.  For a seed with key mine = S_w<<40 | prio:  c < mine  <=> final or held by a
//     higher-priority seed;  c == mine <=> already mine;  c > mine <=> free (a lower-priority claim is taken over).
//   * a seed that re-grows (or dies) RELEASES the claims of its old list (compare-and-swap back to "none");
//   * every event that can invalidate somebody else's growth -- a release, taking a pixel from a lower-priority claim, or
//     finding out (when an accepted pixel is expanded) that it went to a higher-priority seed -- marks the 32x32-pixel tile
//     DIRTY for the next round; a seed is re-verified only if the bounding box of its pixels and refused candidates touches
//     a dirty tile;
//   * verification = every pixel of the list is still mine, every refused candidate is still held (or final);
//   * claims in the grow pass are fire-and-forget atomicMin (no return value on the critical path), so one narrow race is
//     not seen by either party (A loads q free, B claims and expands q, then A's claim lands): before a wave is finalised
//     EVERY live region is verified regardless of dirty tiles; a failure dirties everything and the rounds go on.
// A round without events, confirmed by that full verification, is the fixed point.
#pragma once
#include "lsd_core.h"

namespace olf {
namespace lsd {

constexpr int kTileShift = 5;
struct alignas(32) SeedRec3 { unsigned head; int cnt; unsigned bhead; int bcnt; unsigned short x0, y0, x1, y1; unsigned pad0, pad1; };
struct Ctx3 {
    int W, H;
    PxRec* px;                         // claim[0] is THE claim word (claim[1] unused)
    const short2_t* dabc; const float2_t* tab_seed;
    unsigned* pool; unsigned* pool_ctr; unsigned pool_chunks;
    SeedRec3* srec; double* regang;
    const int* seed_pix; const u64* seed_prio;
    double prec; int fast_align; float c_hi2, c_lo2;
    unsigned* dirty[2]; int tile_wpr;  // dirty bitmaps by round parity, words per tile row
    u64 stamp;                         // S_w << 40 of the current wave
};
#if defined(__CUDA_ARCH__)
// claims are read with a strong (relaxed, gpu-scope) load: program order + coherence make a thread's own earlier
// atomicMin on the same word visible to it, which is how "already mine" is decided without any side table
OLF_HD u64 ld_claim0(const PxRec* r) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&r->claim[0]) : "memory"); return v; }
OLF_HD u64 cas64(u64* p, u64 expect, u64 v) { return atomicCAS(p, expect, v); }
OLF_HD void or32(unsigned* p, unsigned v) { atomicOr(p, v); }
#else
OLF_HD u64 ld_claim0(const PxRec* r) { return r->claim[0]; }
OLF_HD u64 cas64(u64* p, u64 expect, u64 v) { const u64 o = *p; if (o == expect) *p = v; return o; }
OLF_HD void or32(unsigned* p, unsigned v) { *p |= v; }
#endif
OLF_HD void mark_dirty_xy(const Ctx3& C, unsigned round, int x, int y) {
    const int tx = x >> kTileShift, ty = y >> kTileShift;
    or32(&C.dirty[round & 1][ty * C.tile_wpr + (tx >> 5)], 1u << (tx & 31));
}
OLF_HD void mark_dirty(const Ctx3& C, unsigned round, unsigned q) { mark_dirty_xy(C, round, (int)(q % (unsigned)C.W), (int)(q / (unsigned)C.W)); }
// does the bounding box touch a tile that was dirtied in round-1 ?
OLF_HD bool bbox_dirty(const Ctx3& C, unsigned round, const SeedRec3& r) {
    const unsigned* d = C.dirty[(round - 1) & 1];
    const int tx0 = r.x0 >> kTileShift, tx1 = r.x1 >> kTileShift, ty0 = r.y0 >> kTileShift, ty1 = r.y1 >> kTileShift;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int w = tx0 >> 5; w <= tx1 >> 5; ++w) {
            const int lo = tx0 > w * 32 ? tx0 - w * 32 : 0, hi = tx1 < w * 32 + 31 ? tx1 - w * 32 : 31;
            const unsigned mask = (hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            if (d[ty * C.tile_wpr + w] & mask) return true;
        }
    return false;
}
OLF_HD u64 key_of(const Ctx3& C, u64 prio) { return C.stamp | prio; }
OLF_HD bool claim_valid(const Ctx3& C, u64 c) { return (c & ~kPrioMask) == C.stamp; }

// ---- scan ----
OLF_HD bool s3_final(const Ctx3& C, int seed) { return (ld_claim0(&C.px[seed]) >> 40) == 0; }
OLF_HD bool s3_alive(const Ctx3& C, int seed, u64 prio) { return !(ld_claim0(&C.px[seed]) < key_of(C, prio)); }
// Work saver for the FIRST round of a wave (does not change the fixed point): a seed that has a live, higher-priority,
// aligned 8-neighbour will almost surely be absorbed by that neighbour's region, so it sits the round out; from the
// second round on every live seed grows as usual.
OLF_HD bool s3_deferred(const Ctx3& C, int seed, u64 prio) {
    const int W = C.W, H = C.H, py = seed / W, px = seed - py * W;
    float a_s, t0, t1; unsigned b0;
    ld_lo(&C.px[seed], a_s, t0, t1, b0);
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = px + dx, yy = py + dy;
            if ((dx | dy) == 0 || xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
            const int q = yy * W + xx;
            float a_q, cx, cy; unsigned binrev;
            ld_lo(&C.px[q], a_q, cx, cy, binrev);
            if (a_q < 0.f) continue;
            if (make_prio((int)binrev, q) >= prio) continue;
            if ((ld_claim0(&C.px[q]) >> 40) == 0) continue;                    // finalised long ago: cannot absorb us now
            float d = fabsf(a_s - a_q); if (d > 180.f) d = 360.f - d;
            if (d <= 20.f) return true;
        }
    return false;
}
// give back the claims of a list (the seed re-grows or died); every pixel given back dirties its tile
OLF_HD void s3_release(const Ctx3& C, unsigned round, unsigned head, int cnt, u64 mine, unsigned keep) {
    ListReader rd; rd.init(head);
    for (int k = 0; k < cnt; ++k) {
        const unsigned v = rd.next(C.pool);
        if (v == keep) continue;
        if (cas64(&C.px[v].claim[0], mine, kClaimNone) == mine) mark_dirty(C, round, v);
    }
}
// ---- verify ----
enum Verify3 { kV3Dead = 0, kV3Carried = 1, kV3Grow = 2 };
// `alive` = what the scan found.  *changed: the seed's state changed (death, release, first growth pending).
OLF_HD Verify3 s3_verify(const Ctx3& C, unsigned round, int i, bool alive, bool* changed, bool force = false) {
    const int seed = C.seed_pix[i];
    const u64 mine = key_of(C, C.seed_prio[i]);
    SeedRec3 r = C.srec[i];
    if (!alive) {                                                   // died: give everything back
        if (r.cnt > 0) { s3_release(C, round, r.head, r.cnt, mine, kNull); r.cnt = 0; r.bcnt = 0; r.head = kNull; r.bhead = kNull; C.srec[i] = r; *changed = true; }
        return kV3Dead;
    }
    if (r.cnt <= 0) {                                               // (re)born: claim the seed pixel, then grow
        const u64 old = atomic_min64(&C.px[seed].claim[0], mine);
        if (old < mine) return kV3Dead;                             // lost it to a higher-priority seed meanwhile
        if (old != mine && claim_valid(C, old)) mark_dirty(C, round, (unsigned)seed);      // taken from a lower-priority region
        *changed = true;
        return kV3Grow;
    }
    if (!force && !bbox_dirty(C, round, r)) return kV3Carried;      // nothing happened near this region: O(1)
    bool ok = true;
    {
        ListReader rd; rd.init(r.head);
        for (int k = 0; k < r.cnt && ok; ++k) ok = ld_claim0(&C.px[rd.next(C.pool)]) == mine;
    }
    if (ok) {
        ListReader rb; rb.init(r.bhead);
        for (int k = 0; k < r.bcnt && ok; ++k) ok = ld_claim0(&C.px[rb.next(C.pool)]) < mine;
    }
    if (ok) return kV3Carried;
    s3_release(C, round, r.head, r.cnt, mine, (unsigned)seed);      // keeps the seed pixel (still mine: the scan said alive)
    if (ld_claim0(&C.px[seed]) != mine) {                           // ... unless it went meanwhile
        const u64 old = atomic_min64(&C.px[seed].claim[0], mine);
        if (old < mine) { r.cnt = 0; r.bcnt = 0; r.head = kNull; r.bhead = kNull; C.srec[i] = r; *changed = true; return kV3Dead; }
        if (old != mine && claim_valid(C, old)) mark_dirty(C, round, (unsigned)seed);
    }
    r.cnt = 0; r.bcnt = 0; r.head = kNull; r.bhead = kNull; C.srec[i] = r;
    *changed = true;
    return kV3Grow;
}
// ---- grow ----
struct GrowSt3 {
    int i, seed; u64 prio, mine;
    ListWriter wr, bw; ListReader rd;
    int count, done, bcnt;
    int x0, y0, x1, y1;
    float sumdx, sumdy, u2; double reg_angle; bool dirty, overflow;
};
OLF_HD void s3_bbox(GrowSt3& s, int x, int y) { if (x < s.x0) s.x0 = x; if (x > s.x1) s.x1 = x; if (y < s.y0) s.y0 = y; if (y > s.y1) s.y1 = y; }
OLF_HD void s3_begin(const Ctx3& C, int i, GrowSt3& s) {
    s.i = i; s.seed = C.seed_pix[i]; s.prio = C.seed_prio[i]; s.mine = key_of(C, s.prio);
    s.wr.init(); s.bw.init(); s.count = 0; s.done = 0; s.bcnt = 0; s.overflow = false;
    const int sy = s.seed / C.W, sx = s.seed - sy * C.W;
    s.x0 = s.x1 = sx; s.y0 = s.y1 = sy;
    s.wr.push(C.pool, C.pool_ctr, C.pool_chunks, (unsigned)s.seed); if (s.wr.overflow) { s.overflow = true; return; }
    s.count = 1;
    s.rd.init(s.wr.head);
    float a, cx, cy; unsigned b;
    ld_lo(&C.px[s.seed], a, cx, cy, b);
    s.reg_angle = d_mul((double)a, kDegToRads);
    const float2_t t0 = C.tab_seed[tab_index(C.dabc[s.seed])];
    s.sumdx = t0.x; s.sumdy = t0.y;
    s.u2 = f_add(f_mul(s.sumdx, s.sumdx), f_mul(s.sumdy, s.sumdy));
    s.dirty = false;
}
OLF_HD bool s3_aligned(const Ctx3& C, GrowSt3& s, float aq, float cx, float cy) {
    if (C.fast_align && s.u2 > 1e-3f) {
        const float dot = f_add(f_mul(s.sumdx, cx), f_mul(s.sumdy, cy)), d2 = f_mul(dot, dot);
        if (dot > 0.f && d2 >= f_mul(C.c_hi2, s.u2)) return true;
        if (dot <= 0.f || d2 <= f_mul(C.c_lo2, s.u2)) return false;
    }
    if (s.dirty) { s.reg_angle = d_mul((double)fast_atan2_deg(s.sumdy, s.sumdx), kDegToRads); s.dirty = false; }
    double n_theta = d_sub(s.reg_angle, d_mul((double)aq, kDegToRads));
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > k3_2Pi) { n_theta = d_sub(n_theta, k2Pi); if (n_theta < 0) n_theta = -n_theta; }
    return n_theta <= C.prec;
}
// The claim words of a queue entry's 3 x 3 neighbourhood as a thread saw them EARLIER (k_lsd_grow<true> issues the loads of the next entry
// before it decides the candidates of the current one): c[4] is the entry itself.  What the thread itself claimed after the sample was taken
// must be patched in (s3_view_patch) -- a stale view of everybody else's claims is what the passes tolerate anyway.
struct View3 { int p; bool valid; u64 c[9]; };
// sample the view of the entry `ahead` positions behind the next one to be expanded (0 = the next one); false if the queue does not hold it yet
OLF_HD bool s3_peek(const Ctx3& C, const GrowSt3& s, int ahead, View3& v) {
    v.valid = false;
    if (s.done + ahead >= s.count) return false;
    ListReader r = s.rd;
    unsigned p = 0;
    for (int k = 0; k <= ahead; ++k) p = r.next(C.pool);
    v.p = (int)p;
    const int py = v.p / C.W, px = v.p - py * C.W;
    for (int k = 0; k < 9; ++k) {
        const int xx = px + (k % 3) - 1, yy = py + (k / 3) - 1;
        v.c[k] = (xx < 0 || xx >= C.W || yy < 0 || yy >= C.H) ? 0 : ld_claim0(&C.px[yy * C.W + xx]);
    }
    v.valid = true;
    return true;
}
OLF_HD void s3_view_patch(const Ctx3& C, View3& v, unsigned q, u64 mine) {
    if (!v.valid) return;
    const int py = v.p / C.W, px = v.p - py * C.W, qy = (int)q / C.W, qx = (int)q - qy * C.W;
    const int ex = qx - px, ey = qy - py;
    if (ex >= -1 && ex <= 1 && ey >= -1 && ey <= 1) v.c[(ey + 1) * 3 + ex + 1] = mine;
}
// view: the claim words to decide on (nullptr: read them now); claimed / n_claimed: the pixels this step claimed (at most 8), for s3_view_patch
OLF_HD bool s3_step_view(const Ctx3& C, unsigned round, GrowSt3& s, const View3* view, unsigned* claimed, int* n_claimed) {
    const int p = (int)s.rd.next(C.pool);
    ++s.done;
    if (n_claimed) *n_claimed = 0;
    if (view && (!view->valid || view->p != p)) view = nullptr;
    const int py = p / C.W, px = p - py * C.W;
    // the entry itself: accepted a while ago with a fire-and-forget claim -- did it go to a higher-priority seed after all?
    if ((view ? view->c[4] : ld_claim0(&C.px[p])) != s.mine) mark_dirty_xy(C, round, px, py);
    for (int k = 0; k < 9; ++k) {
        if (k == 4) continue;
        const int xx = px + (k % 3) - 1, yy = py + (k / 3) - 1;
        if (xx < 0 || xx >= C.W || yy < 0 || yy >= C.H) continue;
        const int q = yy * C.W + xx;
        float ang, cx, cy; unsigned b;
        ld_lo(&C.px[q], ang, cx, cy, b);
        if (ang < 0.f) continue;                                    // NOTDEF
        const u64 c = view ? view->c[k] : ld_claim0(&C.px[q]);
        if ((c >> 40) == 0 || c == s.mine) continue;                // final / already in this region
        if (!s3_aligned(C, s, ang, cx, cy)) continue;
        if (c < s.mine) {                                           // aligned but held by a higher-priority seed
            s.bw.push(C.pool, C.pool_ctr, C.pool_chunks, (unsigned)q); if (s.bw.overflow) { s.overflow = true; return false; }
            ++s.bcnt; s3_bbox(s, xx, yy);
            continue;
        }
        if (claim_valid(C, c)) mark_dirty_xy(C, round, xx, yy);     // taken from a lower-priority region: it must re-verify
        red_min64(&C.px[q].claim[0], s.mine);
        if (claimed && n_claimed) claimed[(*n_claimed)++] = (unsigned)q;
        s.wr.push(C.pool, C.pool_ctr, C.pool_chunks, (unsigned)q); if (s.wr.overflow) { s.overflow = true; return false; }
        ++s.count; s3_bbox(s, xx, yy);
        s.sumdx = f_add(s.sumdx, cx);
        s.sumdy = f_add(s.sumdy, cy);
        s.u2 = f_add(f_mul(s.sumdx, s.sumdx), f_mul(s.sumdy, s.sumdy));
        s.dirty = true;
    }
    return s.done < s.count;
}
OLF_HD bool s3_step(const Ctx3& C, unsigned round, GrowSt3& s) { return s3_step_view(C, round, s, nullptr, nullptr, nullptr); }
OLF_HD void s3_end(const Ctx3& C, GrowSt3& s) {
    SeedRec3 r; r.pad0 = r.pad1 = 0;
    if (s.overflow) { r.head = kNull; r.cnt = 0; r.bhead = kNull; r.bcnt = 0; r.x0 = r.y0 = r.x1 = r.y1 = 0; C.srec[s.i] = r; return; }
    if (s.dirty) { s.reg_angle = d_mul((double)fast_atan2_deg(s.sumdy, s.sumdx), kDegToRads); s.dirty = false; }
    r.head = s.wr.head; r.cnt = s.count; r.bhead = s.bw.head; r.bcnt = s.bcnt;
    r.x0 = (unsigned short)s.x0; r.y0 = (unsigned short)s.y0; r.x1 = (unsigned short)s.x1; r.y1 = (unsigned short)s.y1;
    C.srec[s.i] = r;
    C.regang[s.i] = s.reg_angle;
}
// ---- finalise a converged wave ----
OLF_HD bool s3_finalize(const Ctx3& C, int i, const FinalOut& F) {
    const SeedRec3 r = C.srec[i];
    if (r.cnt <= 0) return true;
    const u64 prio = C.seed_prio[i];
    const bool accept = r.cnt >= F.min_reg_size;
    unsigned off = 0;
    if (accept) off = atomic_add32(F.final_ctr, (unsigned)r.cnt);
    ListReader rd; rd.init(r.head);
    for (int k = 0; k < r.cnt; ++k) {
        const unsigned v = rd.next(C.pool);
        C.px[v].claim[0] = prio;                                    // stamp 0: final
        if (accept) F.final_pool[off + k] = v;
    }
    if (!accept) return true;
    const unsigned slot = atomic_inc32(F.nreg);
    if (slot >= F.reg_cap) return false;
    LsdRegion R; R.prio = prio; R.off = off; R.count = r.cnt; R.reg_angle = C.regang[i];
    F.regs[slot] = R;
    return true;
}

}  // namespace lsd
}  // namespace olf
