// Core of the parallel LSD region growing (replaces the sequential region_grow loop inside
// cv::LineSegmentDetector, reached from Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:262).
//
// The reference visits seeds in priority order (gradient bin descending, then row-major) and each region marks
// its pixels USED for every later seed.  That is the least fixed point of
//     region(s) = dead                                 if s is inside region(q) for some q of higher priority
//               = grow(s, used = U_{q<s} region(q))    otherwise
// which is well-founded on the priority order, so ANY iteration of that operator that reaches a fixed point
// reaches the sequential result.  We iterate it in parallel: in round t every live seed of the current wave
// re-grows against the claims higher-priority seeds made in round t-1 (complete) and so far in round t (partial,
// same regions once those seeds have stabilised).  A round in which no seed's pixel sequence changed proves the
// fixed point (induction on priority: the highest-priority wrong seed would have seen only correct claims).
// Waves are priority prefixes (whole bins), so each wave is finalised before lower-priority seeds are considered.
//
// Claims: one 64-bit word per pixel and round parity, [stamp:24 | prio:40], updated with atomicMin.
//   stamp = 0xFFFFFF - round  (a later round always wins the min, so stale claims need no clearing)
//   stamp = 0                 pixel belongs to a finalised region (wins forever)
//   prio  = (n_bins-1-bin) << 30 | pixel index   (smaller = visited earlier by the reference)
//
// The same source compiles for the device (lsd.cu) and for the host (tests/emul: sequential emulation of the
// rounds, used to validate the algorithm against the oracle without a GPU).
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

struct GrowArgs {
    int W, H;
    const float* ang;             // gradient angle in degrees (cv::fastAtan2(gx,-gy)); < 0 = NOTDEF
    const short2_t* dabc;         // (DA, BC) per pixel; gx = DA+BC, gy = DA-BC
    const float2_t* tab_seed;     // [(DA+255)*511 + BC+255] -> (float(cos(a)), float(sin(a))), a = double angle
    const float2_t* tab_acc;      //                         -> (float(cos((double)(float)a)), float(sin(...)))
    u64* claim[2];                // claim words per round parity
    unsigned* pool[2];            // chunked pixel lists per round parity
    unsigned* pool_ctr[2];        // chunks allocated
    unsigned pool_chunks;         // capacity in chunks
    double prec;                  // pi * ang_th / 180
};

OLF_HD int tab_index(short2_t d) { return ((int)d.x + 255) * kTabDim + ((int)d.y + 255); }

// is pixel q unavailable to a seed of priority `prio` in round `round`?
OLF_HD bool blocked(const GrowArgs& A, unsigned round, int q, u64 prio) {
    const u64 sf_prev = stamp_field(round - 1), sf_cur = stamp_field(round);
    u64 e = A.claim[(round - 1) & 1][q];
    u64 sf = e >> 40;
    if (sf == 0) return true;
    if (sf == sf_prev && (e & kPrioMask) < prio) return true;
    e = A.claim[round & 1][q];
    sf = e >> 40;
    if (sf == 0) return true;
    if (sf == sf_cur && (e & kPrioMask) <= prio) return true;       // == prio: already mine
    return false;
}

struct ListWriter {
    unsigned* pool; unsigned* ctr; unsigned cap;
    unsigned head, chunk; int off; int count; bool overflow;
    OLF_HD void init(unsigned* pool_, unsigned* ctr_, unsigned cap_) {
        pool = pool_; ctr = ctr_; cap = cap_; head = kNull; chunk = kNull; off = kChunk - 1; count = 0; overflow = false;
    }
    OLF_HD void push(unsigned pix) {
        if (off == kChunk - 1) {
            const unsigned nc = atomic_inc32(ctr);
            if (nc >= cap) { overflow = true; return; }
            pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
            if (chunk == kNull) head = nc; else pool[(size_t)chunk * kChunk + kChunk - 1] = nc;
            chunk = nc; off = 0;
        }
        pool[(size_t)chunk * kChunk + off++] = pix;
        ++count;
    }
};
struct ListReader {
    const unsigned* pool; unsigned chunk; int off;
    OLF_HD void init(const unsigned* pool_, unsigned head) { pool = pool_; chunk = head; off = 0; }
    OLF_HD unsigned next() {
        if (off == kChunk - 1) { chunk = pool[(size_t)chunk * kChunk + kChunk - 1]; off = 0; }
        return pool[(size_t)chunk * kChunk + off++];
    }
};

struct GrowResult { unsigned head; int count; double reg_angle; bool same_as_prev; bool overflow; };

// One seed, one round.  prev_head/prev_count describe the seed's list of round-1 (count 0 = was dead / first round).
OLF_HD GrowResult grow_seed(const GrowArgs& A, unsigned round, int seed, u64 prio, unsigned prev_head, int prev_count) {
    GrowResult R;
    R.head = kNull; R.count = 0; R.reg_angle = 0; R.overflow = false;
    if (blocked(A, round, seed, prio)) { R.same_as_prev = (prev_count == 0); return R; }
    const u64 mine = (stamp_field(round) << 40) | prio;
    ListWriter wr; wr.init(A.pool[round & 1], A.pool_ctr[round & 1], A.pool_chunks);
    ListReader rd; rd.init(A.pool[round & 1], kNull);
    ListReader pv; pv.init(A.pool[(round - 1) & 1], prev_head);
    bool same = prev_count > 0;
    atomic_min64(&A.claim[round & 1][seed], mine);
    wr.push((unsigned)seed);
    if (wr.overflow) { R.overflow = true; R.same_as_prev = false; return R; }
    rd.chunk = wr.head;
    if (same && pv.next() != (unsigned)seed) same = false;
    double reg_angle = d_mul((double)A.ang[seed], kDegToRads);
    const float2_t t0 = A.tab_seed[tab_index(A.dabc[seed])];
    float sumdx = t0.x, sumdy = t0.y;
    for (int i = 0; i < wr.count; ++i) {
        const int p = (int)rd.next();
        const int px = p % A.W, py = p / A.W;
        const int x0 = px > 0 ? px - 1 : 0, x1 = px < A.W - 1 ? px + 1 : A.W - 1;
        const int y0 = py > 0 ? py - 1 : 0, y1 = py < A.H - 1 ? py + 1 : A.H - 1;
        for (int yy = y0; yy <= y1; ++yy)
            for (int xx = x0; xx <= x1; ++xx) {
                const int q = yy * A.W + xx;
                const float aq = A.ang[q];
                if (aq < 0.f) continue;                               // NOTDEF
                // isAligned(xx, yy, reg_angle, prec)
                double n_theta = d_sub(reg_angle, d_mul((double)aq, kDegToRads));
                if (n_theta < 0) n_theta = -n_theta;
                if (n_theta > k3_2Pi) { n_theta = d_sub(n_theta, k2Pi); if (n_theta < 0) n_theta = -n_theta; }
                if (!(n_theta <= A.prec)) continue;
                if (blocked(A, round, q, prio)) continue;
                atomic_min64(&A.claim[round & 1][q], mine);
                wr.push((unsigned)q);
                if (wr.overflow) { R.overflow = true; R.same_as_prev = false; return R; }
                if (same) { if (wr.count > prev_count || pv.next() != (unsigned)q) same = false; }
                const float2_t t = A.tab_acc[tab_index(A.dabc[q])];
                sumdx = f_add(sumdx, t.x);
                sumdy = f_add(sumdy, t.y);
                reg_angle = d_mul((double)fast_atan2_deg(sumdy, sumdx), kDegToRads);
            }
    }
    R.head = wr.head; R.count = wr.count; R.reg_angle = reg_angle;
    R.same_as_prev = same && wr.count == prev_count;
    return R;
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
