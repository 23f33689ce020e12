// Core of the parallel LSD region growing (replaces the sequential region_grow loop inside
// cv::LineSegmentDetector, reached from Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:262).
//
// The reference visits seeds in priority order (gradient bin descending, then row-major) and each region marks
// its pixels USED for every later seed.  That is the least fixed point of
//     region(s) = dead                                 if s is inside region(q) for some q of higher priority
//               = grow(s, used = U_{q<s} region(q))    otherwise
// which is well-founded on the priority order, so ANY iteration of that operator that reaches a fixed point
// reaches the sequential result.  We iterate it in parallel: in round t every live seed of the current wave
// grows against the claims higher-priority seeds made in round t-1 (complete) and so far in round t (partial,
// same regions once those seeds have stabilised).  A round in which no seed's pixel sequence changed proves the
// fixed point (induction on priority: the highest-priority wrong seed would have seen only correct claims).
// Waves are priority prefixes (whole bins), so each wave is finalised before lower-priority seeds are considered.
//
// Claims: one 64-bit word per pixel and round parity, [stamp:24 | prio:40], updated with atomicMin.
//   stamp = 0xFFFFFF - round  (a later round always wins the min, so stale claims need no clearing)
//   stamp = 0                 pixel belongs to a finalised region (wins forever)
//   prio  = (n_bins-1-bin) << 30 | pixel index   (smaller = visited earlier by the reference)
//
// One round = three data-parallel passes, each ONE THREAD PER SEED (the functions below):
//   scan    every seed of the wave: dead (its pixel is held by a higher-priority claim) or alive -> work list
//   verify  every alive seed: claim the seed pixel; if last round's growth is still exact (every pixel it accepted is
//           still free of higher-priority claims, every aligned candidate it was refused is still held) the region is
//           carried over -- its claims are re-stamped for this round -- otherwise the seed goes to the grow list
//   grow    every seed of the grow list re-grows: breadth-first, one queue entry per step (grow_step)
// A growth depends only on (a) the pixels it accepted and (b) the aligned candidates it was refused by a NON-final
// higher-priority claim; both sets are recorded, which is what makes the O(n) parallel check in `verify` exact.
//
// The same source compiles for the device (line.cu) and for the host (tests/emul: sequential emulation of the
// passes with random schedules, used to validate the algorithm against the oracle without a GPU).
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define OLF_HD __host__ __device__ __forceinline__
#else
#define OLF_HD inline
#endif

namespace olf {
namespace lsd {

typedef unsigned long long u64;

constexpr double kDegToRads = 0x1.1df46a2529d39p-6;      // CV_PI / 180
constexpr double k3_2Pi = 4.71238898038;                  // M_3_2_PI in OpenCV's lsd.cpp
constexpr double k2Pi = 6.28318530718;                    // M_2__PI
constexpr u64 kPrioMask = (1ull << 40) - 1;
constexpr u64 kClaimNone = ~0ull;
constexpr unsigned kNull = 0xFFFFFFFFu;
constexpr int kChunk = 32;                                // 31 pixels + next pointer
constexpr int kTabDim = 511;                              // (DA, BC) in [-255, 255]

OLF_HD u64 make_prio(int bin_rev, int idx) { return ((u64)(unsigned)bin_rev << 30) | (u64)(unsigned)idx; }
OLF_HD u64 stamp_field(unsigned round) { return (u64)(0xFFFFFFu - round); }

#if defined(__CUDA_ARCH__)
OLF_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
OLF_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
OLF_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
OLF_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
OLF_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
OLF_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
OLF_HD double d_sub(double a, double b) { return __dsub_rn(a, b); }
OLF_HD double d_div(double a, double b) { return __ddiv_rn(a, b); }
OLF_HD double d_sqrt(double a) { return __dsqrt_rn(a); }
OLF_HD u64 atomic_min64(u64* p, u64 v) { return atomicMin(p, v); }
OLF_HD unsigned atomic_inc32(unsigned* p) { return atomicAdd(p, 1u); }
OLF_HD unsigned atomic_add32(unsigned* p, unsigned v) { return atomicAdd(p, v); }
#else   // host build: compile with -ffp-contract=off
OLF_HD float f_mul(float a, float b) { return a * b; }
OLF_HD float f_add(float a, float b) { return a + b; }
OLF_HD float f_sub(float a, float b) { return a - b; }
OLF_HD float f_div(float a, float b) { return a / b; }
OLF_HD double d_mul(double a, double b) { return a * b; }
OLF_HD double d_add(double a, double b) { return a + b; }
OLF_HD double d_sub(double a, double b) { return a - b; }
OLF_HD double d_div(double a, double b) { return a / b; }
OLF_HD double d_sqrt(double a) { return std::sqrt(a); }
OLF_HD u64 atomic_min64(u64* p, u64 v) { u64 o = *p; if (v < o) *p = v; return o; }
OLF_HD unsigned atomic_inc32(unsigned* p) { return (*p)++; }
OLF_HD unsigned atomic_add32(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
#endif

// cv::fastAtan2 (degrees), SURVEY A.4
OLF_HD float fast_atan2_deg(float y, float x) {
    const float p1 = 0x1.ca44dep+5f, p3 = -0x1.2aaddcp+4f, p5 = 0x1.1d3f7ep+3f, p7 = -0x1.4515b2p+1f, eps = 0x1p-52f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = f_div(ay, f_add(ax, eps));
        c2 = f_mul(c, c);
        a = f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = f_div(ax, f_add(ay, eps));
        c2 = f_mul(c, c);
        a = f_sub(90.f, f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = f_sub(180.f, a);
    if (y < 0) a = f_sub(360.f, a);
    return a;
}

struct short2_t { short x, y; };
struct float2_t { float x, y; };

// Everything a neighbour test needs sits in ONE 32-byte sector: both claim words, the level-line angle, the (cos, sin)
// the reference accumulates ((float)cos((double)(float)a), host libm table keyed by (DA, BC)) and the priority bin.
struct alignas(32) PxRec { u64 claim[2]; float ang; float cx; float cy; unsigned binrev; };
// per seed and round parity: its pixel list (head chunk + count) and the list of refused candidates (same chunked layout;
// a pixel refused from several queue entries appears several times)
struct alignas(16) SeedRec { unsigned head; int cnt; unsigned bchunk; int bcnt; };

#if defined(__CUDA_ARCH__)
// claims are read with a strong (relaxed, gpu-scope) load: program order + coherence make a thread's own earlier
// atomicMin on the same word visible to it, which is how "already mine" is decided without any side table
OLF_HD void ld_claims(const PxRec* r, u64& c0, u64& c1) {
#ifdef OLF_EXP_WEAK
    asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "l"(r) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "l"(r) : "memory");
#endif
}
OLF_HD u64 ld_claim(const PxRec* r, int parity) {
    u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&r->claim[parity]) : "memory"); return v;
}
OLF_HD void ld_lo(const PxRec* r, float& ang, float& cx, float& cy, unsigned& binrev) {
    const float4 f = __ldcg(reinterpret_cast<const float4*>(r) + 1);
    ang = f.x; cx = f.y; cy = f.z; binrev = __float_as_uint(f.w);
}
OLF_HD void red_min64(u64* p, u64 v) { atomicMin(p, v); }          // result unused -> RED (fire and forget)
#else
OLF_HD void ld_claims(const PxRec* r, u64& c0, u64& c1) { c0 = r->claim[0]; c1 = r->claim[1]; }
OLF_HD u64 ld_claim(const PxRec* r, int parity) { return r->claim[parity]; }
OLF_HD void ld_lo(const PxRec* r, float& ang, float& cx, float& cy, unsigned& binrev) { ang = r->ang; cx = r->cx; cy = r->cy; binrev = r->binrev; }
OLF_HD void red_min64(u64* p, u64 v) { if (v < *p) *p = v; }
#endif

struct GrowCtx {
    int W, H;
    PxRec* px;
    const short2_t* dabc;         // (DA, BC) per pixel; gx = DA+BC, gy = DA-BC
    const float2_t* tab_seed;     // [(DA+255)*511 + BC+255] -> (float(cos(a)), float(sin(a))), a = double angle (seed pixel only)
    unsigned* pool;               // chunked pixel lists: one bump pool per wave (lists are carried over rounds)
    unsigned* pool_ctr; unsigned pool_chunks;
    SeedRec* srec[2];             // per seed, per round parity
    double* regang;               // per seed: region angle of its latest growth
    const int* seed_pix; const u64* seed_prio;
    double prec;                  // pi * ang_th / 180
    int fast_align; float c_hi2, c_lo2;   // lazy alignment test: cos^2(prec -/+ 0.1 deg)
};

OLF_HD int tab_index(short2_t d) { return ((int)d.x + 255) * kTabDim + ((int)d.y + 255); }

// is a pixel with claim words (e_prev, e_cur) unavailable to a seed of priority `prio`?
OLF_HD bool blocked_vals(u64 e_prev, u64 e_cur, u64 sf_prev, u64 sf_cur, u64 prio) {
    u64 sf = e_prev >> 40;
    if (sf == 0) return true;
    if (sf == sf_prev && (e_prev & kPrioMask) < prio) return true;
    sf = e_cur >> 40;
    if (sf == 0) return true;
    if (sf == sf_cur && (e_cur & kPrioMask) <= prio) return true;       // == prio: already mine
    return false;
}

// ---- pass 1: scan ------------------------------------------------------------------------------------------------
// Is the seed alive at the start of round `round`?  (non-authoritative: the verify pass claims the seed pixel)
OLF_HD bool seed_alive(const GrowCtx& C, unsigned round, int seed, u64 prio) {
    u64 c0, c1; ld_claims(&C.px[seed], c0, c1);
    const int cur = round & 1;
    return !blocked_vals(cur ? c0 : c1, cur ? c1 : c0, stamp_field(round - 1), stamp_field(round), prio);
}
// A seed whose pixel belongs to a finalised region is dead for good: the first scan of a wave drops it from the wave's
// candidate list, later rounds only look at the candidates.
OLF_HD bool seed_final(const GrowCtx& C, int seed) {
    u64 c0, c1; ld_claims(&C.px[seed], c0, c1);
    return (c0 >> 40) == 0 || (c1 >> 40) == 0;
}
// Work saver for the FIRST round of a wave (does not change the fixed point): a seed that has a live, higher-priority,
// aligned 8-neighbour will almost surely be absorbed by that neighbour's region, so it sits the round out; from the
// second round on every live seed grows as usual.
OLF_HD bool seed_deferred(const GrowCtx& C, unsigned round, int seed, u64 prio) {
    const int W = C.W, H = C.H, py = seed / W, px = seed - py * W, prv = (round - 1) & 1;
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
            if ((ld_claim(&C.px[q], prv) >> 40) == 0) continue;                 // finalised long ago: cannot absorb us now
            float d = fabsf(a_s - a_q); if (d > 180.f) d = 360.f - d;
            if (d <= 20.f) return true;
        }
    return false;
}

// ---- chunked pixel lists (31 pixels + next pointer per 128-byte chunk) -----------------------------------------------
struct ListWriter {
    unsigned head, chunk; int off; bool overflow;
    OLF_HD void init() { head = kNull; chunk = kNull; off = kChunk - 1; overflow = false; }
    OLF_HD void push(const GrowCtx& C, unsigned pix) {
        if (off == kChunk - 1) {
            const unsigned nc = atomic_inc32(C.pool_ctr);
            if (nc >= C.pool_chunks) { overflow = true; return; }
            C.pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
            if (chunk == kNull) head = nc; else C.pool[(size_t)chunk * kChunk + kChunk - 1] = nc;
            chunk = nc; off = 0;
        }
        C.pool[(size_t)chunk * kChunk + off++] = pix;
    }
};
struct ListReader {
    unsigned chunk; int off;
    OLF_HD void init(unsigned head) { chunk = head; off = 0; }
    OLF_HD unsigned next(const unsigned* pool) {
        if (off == kChunk - 1) { chunk = pool[(size_t)chunk * kChunk + kChunk - 1]; off = 0; }
        return pool[(size_t)chunk * kChunk + off++];
    }
};

#if defined(__CUDACC__)
#define OLF_UNROLL _Pragma("unroll")
#else
#define OLF_UNROLL
#endif
#if defined(__CUDA_ARCH__)
struct Quad { unsigned v0, v1, v2, v3; };
OLF_HD Quad ld_u32x4(const unsigned* p) { const uint4 q = *reinterpret_cast<const uint4*>(p); Quad r; r.v0 = q.x; r.v1 = q.y; r.v2 = q.z; r.v3 = q.w; return r; }
#else
struct Quad { unsigned v0, v1, v2, v3; };
OLF_HD Quad ld_u32x4(const unsigned* p) { Quad r; r.v0 = p[0]; r.v1 = p[1]; r.v2 = p[2]; r.v3 = p[3]; return r; }
#endif
OLF_HD unsigned ld_u32(const unsigned* p) { return *p; }
// Walk a chunked list FOUR entries per step: one 16-byte load of pixel indices (chunks are 128-byte aligned, entries
// 28..30 share their quad with the link), then four independent claim loads -- two memory round trips per four pixels
// instead of eight.  `what` 0: all pixels free of higher-priority claims of round-1?  1: all pixels still held?
// 2: re-stamp for this round (always true).
OLF_HD bool walk_bad(u64 ep, int what, u64 sf_prev, u64 prio) {
    const u64 sf = ep >> 40;
    const bool higher = sf == sf_prev && (ep & kPrioMask) < prio;
    return what == 0 ? higher : !(sf == 0 || higher);
}
OLF_HD bool walk_list4(const GrowCtx& C, unsigned head, int cnt, int what, int prv, int cur, u64 sf_prev, u64 prio, u64 mine, unsigned skip) {
    unsigned chunk = head;
    for (int k = 0; k < cnt; k += kChunk - 1) {
        const int nc = cnt - k < kChunk - 1 ? cnt - k : kChunk - 1;
        const unsigned* base = &C.pool[(size_t)chunk * kChunk];
        for (int j = 0; j < nc; j += 4) {
            const Quad q = ld_u32x4(base + j);
            const int m = nc - j;                                   // entries of this quad that belong to the list: min(m, 4)
            if (what == 2) {
                if (q.v0 != skip) red_min64(&C.px[q.v0].claim[cur], mine);
                if (m > 1 && q.v1 != skip) red_min64(&C.px[q.v1].claim[cur], mine);
                if (m > 2 && q.v2 != skip) red_min64(&C.px[q.v2].claim[cur], mine);
                if (m > 3 && q.v3 != skip) red_min64(&C.px[q.v3].claim[cur], mine);
            } else {
                const u64 e0 = ld_claim(&C.px[q.v0], prv);
                const u64 e1 = m > 1 ? ld_claim(&C.px[q.v1], prv) : 0;
                const u64 e2 = m > 2 ? ld_claim(&C.px[q.v2], prv) : 0;
                const u64 e3 = m > 3 ? ld_claim(&C.px[q.v3], prv) : 0;
                bool bad = walk_bad(e0, what, sf_prev, prio);
                bad |= m > 1 && walk_bad(e1, what, sf_prev, prio);
                bad |= m > 2 && walk_bad(e2, what, sf_prev, prio);
                bad |= m > 3 && walk_bad(e3, what, sf_prev, prio);
                if (bad) return false;
            }
        }
        if (k + nc < cnt) chunk = ld_u32(base + kChunk - 1);
    }
    return true;
}

// ---- pass 2: verify -----------------------------------------------------------------------------------------------
enum VerifyResult { kSeedDead = 0, kSeedCarried = 1, kSeedGrow = 2, kSeedLong = 3 };
// One alive seed (index i in the seed arrays).  `have_prev` = the seed may own a list from round-1 (false in the first
// round of a wave).  Returns what happened; *changed is set when the seed's outcome differs from the previous round.
// Lists longer than `long_cnt` are left to the caller (kSeedLong: the seed pixel is claimed, the lists are unchecked) --
// the kernel walks those with a whole warp.
OLF_HD VerifyResult verify_seed(const GrowCtx& C, unsigned round, int i, bool have_prev, bool* changed, int long_cnt = 0x7fffffff) {
    const int cur = round & 1, prv = cur ^ 1;
    const u64 sf_prev = stamp_field(round - 1), sf_cur = stamp_field(round);
    const int seed = C.seed_pix[i];
    const u64 prio = C.seed_prio[i];
    const u64 mine = (sf_cur << 40) | prio;
    SeedRec pr; pr.head = kNull; pr.cnt = 0; pr.bchunk = kNull; pr.bcnt = 0;
    if (have_prev) pr = C.srec[prv][i];
    // claim the seed pixel (authoritative check through the atomic's return value)
    bool dead;
    {
        const u64 ep = ld_claim(&C.px[seed], prv);
        u64 sf = ep >> 40;
        dead = (sf == 0) || (sf == sf_prev && (ep & kPrioMask) < prio);
        if (!dead) {
            const u64 old = atomic_min64(&C.px[seed].claim[cur], mine);
            sf = old >> 40;
            dead = (sf == 0) || (sf == sf_cur && (old & kPrioMask) < prio);
        }
    }
    if (dead) {
        SeedRec z; z.head = kNull; z.cnt = 0; z.bchunk = kNull; z.bcnt = 0;
        C.srec[cur][i] = z;
        if (pr.cnt != 0) *changed = true;
        return kSeedDead;
    }
    if (pr.cnt <= 0) return kSeedGrow;
    if (pr.cnt + pr.bcnt > long_cnt) return kSeedLong;
    // every pixel accepted last round must still be free of higher-priority claims of that round
    if (!walk_list4(C, pr.head, pr.cnt, 0, prv, cur, sf_prev, prio, mine, kNull)) return kSeedGrow;
    // every aligned candidate refused last round (by a non-final claim) must still be held
    if (!walk_list4(C, pr.bchunk, pr.bcnt, 1, prv, cur, sf_prev, prio, mine, kNull)) return kSeedGrow;
    // carry the region over: re-stamp its claims for this round
    walk_list4(C, pr.head, pr.cnt, 2, prv, cur, sf_prev, prio, mine, (unsigned)seed);
    C.srec[cur][i] = pr;
    return kSeedCarried;
}

// ---- pass 3: grow -------------------------------------------------------------------------------------------------
// Resumable breadth-first growth of one seed: grow_begin(), then grow_step() once per queue entry while it returns true,
// then grow_end().  The order of every decision is the reference's: queue order, 3x3 neighbourhood row-major
// (yy outer, xx inner), the region angle updated after every accepted pixel.
struct GrowSt {
    int i, seed; u64 prio;
    ListWriter wr; ListReader rd, pv;
    int count, done;                      // pixels in the list / queue entries processed
    int prev_cnt; bool same;
    ListWriter bw; int bcnt;              // refused candidates
    float sumdx, sumdy, u2; double reg_angle; bool dirty;
    bool overflow;
};

OLF_HD void grow_push(const GrowCtx& C, GrowSt& s, unsigned pix) {
    s.wr.push(C, pix);
    if (s.wr.overflow) { s.overflow = true; return; }
    if (s.same) {                                                    // lock-step comparison with last round's list
        if (s.count >= s.prev_cnt) s.same = false;
        else if (s.pv.next(C.pool) != pix) s.same = false;
    }
    ++s.count;
}
OLF_HD void grow_record_blocked(const GrowCtx& C, GrowSt& s, unsigned pix) {
    s.bw.push(C, pix);
    if (s.bw.overflow) { s.overflow = true; return; }
    ++s.bcnt;
}

// the seed pixel has already been claimed for this round by verify_seed()
OLF_HD void grow_begin(const GrowCtx& C, unsigned round, int i, bool have_prev, GrowSt& s) {
    const int prv = (round - 1) & 1;
    s.i = i; s.seed = C.seed_pix[i]; s.prio = C.seed_prio[i];
    s.wr.init(); s.count = 0; s.done = 0; s.overflow = false;
    s.bw.init(); s.bcnt = 0;
    SeedRec pr; pr.head = kNull; pr.cnt = 0;
    if (have_prev) pr = C.srec[prv][i];
    s.prev_cnt = pr.cnt > 0 ? pr.cnt : 0;
    s.same = s.prev_cnt > 0;
    s.pv.init(pr.head);
    grow_push(C, s, (unsigned)s.seed);
    s.rd.init(s.wr.head);
    float a, cx, cy; unsigned b;
    ld_lo(&C.px[s.seed], a, cx, cy, b);
    s.reg_angle = d_mul((double)a, kDegToRads);
    const float2_t t0 = C.tab_seed[tab_index(C.dabc[s.seed])];
    s.sumdx = t0.x; s.sumdy = t0.y;
    s.u2 = f_add(f_mul(s.sumdx, s.sumdx), f_mul(s.sumdy, s.sumdy));
    s.dirty = false;
}

// isAligned(reg_angle, a) with reg_angle = fastAtan2(sumdy, sumdx) evaluated lazily: the sign of
// cos(angular distance) - cos(prec -/+ 0.1 deg) (float dot product, no division) decides all but the candidates within
// 0.1 deg of the threshold -- 10x the worst error of the fastAtan2 polynomial -- which take the reference's exact
// double arithmetic.
OLF_HD bool grow_aligned(const GrowCtx& C, GrowSt& s, float aq, float cx, float cy) {
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

// one queue entry; returns false when the queue is exhausted (or on pool overflow)
OLF_HD bool grow_step(const GrowCtx& C, unsigned round, GrowSt& s) {
    const int cur = round & 1;
    const u64 sf_prev = stamp_field(round - 1), sf_cur = stamp_field(round);
    const u64 mine = (sf_cur << 40) | s.prio;
    const int p = (int)s.rd.next(C.pool);
    ++s.done;
    const int py = p / C.W, px = p - py * C.W;
    // fetch the whole neighbourhood first (independent loads), then decide in scan order
    u64 c0[9], c1[9]; float ang[9], ncx[9], ncy[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 9; ++k) {
        const int xx = px + (k % 3) - 1, yy = py + (k / 3) - 1;
        ang[k] = -1.f; c0[k] = 0; c1[k] = 0; ncx[k] = 0.f; ncy[k] = 0.f;
        if (k != 4 && xx >= 0 && xx < C.W && yy >= 0 && yy < C.H) {
            const PxRec* r = &C.px[yy * C.W + xx];
            unsigned b;
            ld_lo(r, ang[k], ncx[k], ncy[k], b);
            ld_claims(r, c0[k], c1[k]);
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 9; ++k) {
        if (ang[k] < 0.f) continue;                                   // NOTDEF, outside the image, or the entry itself
        const u64 ep = cur ? c0[k] : c1[k], ec = cur ? c1[k] : c0[k];
        const u64 sp = ep >> 40, sc = ec >> 40;
        if (sp == 0 || sc == 0) continue;                             // finalised region (USED for good)
        if (sc == sf_cur && (ec & kPrioMask) == s.prio) continue;     // already in this region
        if (!grow_aligned(C, s, ang[k], ncx[k], ncy[k])) continue;
        const int q = (py + (k / 3) - 1) * C.W + px + (k % 3) - 1;
        if ((sp == sf_prev && (ep & kPrioMask) < s.prio) || (sc == sf_cur && (ec & kPrioMask) < s.prio)) {
            grow_record_blocked(C, s, (unsigned)q);                   // aligned but held by a higher-priority seed
            if (s.overflow) return false;
            continue;
        }
        red_min64(&C.px[q].claim[cur], mine);
        grow_push(C, s, (unsigned)q);
        if (s.overflow) return false;
        s.sumdx = f_add(s.sumdx, ncx[k]);
        s.sumdy = f_add(s.sumdy, ncy[k]);
        s.u2 = f_add(f_mul(s.sumdx, s.sumdx), f_mul(s.sumdy, s.sumdy));
        s.dirty = true;
    }
    return s.done < s.count;
}

// returns true when the seed's outcome differs from the previous round
OLF_HD bool grow_end(const GrowCtx& C, unsigned round, GrowSt& s) {
    const int cur = round & 1;
    SeedRec r;
    if (s.overflow) { r.head = kNull; r.cnt = 0; r.bchunk = kNull; r.bcnt = 0; C.srec[cur][s.i] = r; return true; }
    if (s.dirty) { s.reg_angle = d_mul((double)fast_atan2_deg(s.sumdy, s.sumdx), kDegToRads); s.dirty = false; }
    r.head = s.wr.head; r.cnt = s.count; r.bchunk = s.bw.head; r.bcnt = s.bcnt;
    C.srec[cur][s.i] = r;
    C.regang[s.i] = s.reg_angle;
    return !(s.same && s.count == s.prev_cnt);
}

// ---- finalise ------------------------------------------------------------------------------------------------------
// One live seed of a converged wave: its pixels are stamped for good (stamp 0 wins every later atomicMin); regions of
// at least min_reg_size pixels are copied to the contiguous final pool for the rectangle fit.
struct LsdRegion { u64 prio; unsigned off; int count; double reg_angle; };
struct FinalOut { unsigned* final_pool; unsigned* final_ctr; LsdRegion* regs; unsigned* nreg; unsigned reg_cap; int min_reg_size; };
// returns false if the region table is full
OLF_HD bool finalize_seed(const GrowCtx& C, unsigned round, int i, const FinalOut& F) {
    const SeedRec r = C.srec[round & 1][i];
    if (r.cnt <= 0) return true;
    const u64 prio = C.seed_prio[i];
    const bool accept = r.cnt >= F.min_reg_size;
    unsigned off = 0;
    if (accept) off = atomic_add32(F.final_ctr, (unsigned)r.cnt);
    ListReader rd; rd.init(r.head);
    for (int k = 0; k < r.cnt; ++k) {
        const unsigned v = rd.next(C.pool);
        C.px[v].claim[0] = prio; C.px[v].claim[1] = prio;
        if (accept) F.final_pool[off + k] = v;
    }
    if (!accept) return true;
    const unsigned slot = atomic_inc32(F.nreg);
    if (slot >= F.reg_cap) return false;
    LsdRegion R; R.prio = prio; R.off = off; R.count = r.cnt; R.reg_angle = C.regang[i];
    F.regs[slot] = R;
    return true;
}

OLF_HD double angle_diff(double a, double b) {
    double diff = d_sub(a, b);
    while (diff <= -M_PI) diff = d_add(diff, k2Pi);
    while (diff > M_PI) diff = d_sub(diff, k2Pi);
    if (diff < 0) diff = -diff;
    return diff;
}

// region2rect part 1 (centroid + orientation), sequential double sums in region order.
// pix = contiguous list of the region's pixels in BFS order.
struct RectA { double x, y, theta; };
OLF_HD RectA region_rect_a(const unsigned* pix, int n, const short2_t* dabc, int W, double reg_angle, double prec) {
    double x = 0, y = 0, sum = 0;
    for (int i = 0; i < n; ++i) {
        const int p = (int)pix[i];
        const short2_t d = dabc[p];
        const int gx = d.x + d.y, gy = d.x - d.y;
        const double w = d_sqrt(d_div((double)(gx * gx + gy * gy), 4.0));
        x = d_add(x, d_mul((double)(p % W), w));
        y = d_add(y, d_mul((double)(p / W), w));
        sum = d_add(sum, w);
    }
    x = d_div(x, sum); y = d_div(y, sum);
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i = 0; i < n; ++i) {
        const int p = (int)pix[i];
        const short2_t d = dabc[p];
        const int gx = d.x + d.y, gy = d.x - d.y;
        const double w = d_sqrt(d_div((double)(gx * gx + gy * gy), 4.0));
        const double dx = d_sub((double)(p % W), x), dy = d_sub((double)(p / W), y);
        Ixx = d_add(Ixx, d_mul(d_mul(dy, dy), w));
        Iyy = d_add(Iyy, d_mul(d_mul(dx, dx), w));
        Ixy = d_sub(Ixy, d_mul(d_mul(dx, dy), w));
    }
    const double dI = d_sub(Ixx, Iyy);
    const double lambda = d_mul(0.5, d_sub(d_add(Ixx, Iyy), d_sqrt(d_add(d_mul(dI, dI), d_mul(d_mul(4.0, Ixy), Ixy)))));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)d_sub(lambda, Ixx), (float)Ixy)
                                            : (double)fast_atan2_deg((float)Ixy, (float)d_sub(lambda, Iyy));
    theta = d_mul(theta, kDegToRads);
    if (angle_diff(theta, reg_angle) > prec) theta = d_add(theta, M_PI);
    RectA r; r.x = x; r.y = y; r.theta = theta;
    return r;
}
// per-pixel projection on the region direction (order-free min / max over the region)
OLF_HD double region_proj(unsigned p, int W, double cx, double cy, double dx, double dy) {
    const double regdx = d_sub((double)((int)p % W), cx), regdy = d_sub((double)((int)p / W), cy);
    return d_add(d_mul(regdx, dx), d_mul(regdy, dy));
}

}  // namespace lsd
}  // namespace olf
