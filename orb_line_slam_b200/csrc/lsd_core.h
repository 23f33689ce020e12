// Shared primitives of the parallel LSD region growing (replaces the sequential region_grow loop inside
// cv::LineSegmentDetector, reached from Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:262): exact float helpers,
// cv::fastAtan2, the per-pixel record, chunked pixel lists, and the rectangle fit (region2rect).  The growing scheme itself
// -- claims, rounds, waves, the scan / verify / grow passes -- is lsd_sticky.h.
//
// The same source compiles for the device (line.cu) and for the host (tests/emul: sequential emulation of the passes with
// random schedules, used to validate the algorithm against the oracle without a GPU).
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

// Everything a neighbour test needs sits in ONE 32-byte sector: the claim word (claim[0]; claim[1] is spare), the level-line
// angle, the (cos, sin) the reference accumulates ((float)cos((double)(float)a), host libm table keyed by (DA, BC)) and the
// priority bin.
struct alignas(32) PxRec { u64 claim[2]; float ang; float cx; float cy; unsigned binrev; };
#if defined(__CUDA_ARCH__)
OLF_HD void ld_lo(const PxRec* r, float& ang, float& cx, float& cy, unsigned& binrev) {
    const float4 f = __ldcg(reinterpret_cast<const float4*>(r) + 1);
    ang = f.x; cx = f.y; cy = f.z; binrev = __float_as_uint(f.w);
}
OLF_HD void red_min64(u64* p, u64 v) { atomicMin(p, v); }          // result unused -> RED (fire and forget)
#else
OLF_HD void ld_lo(const PxRec* r, float& ang, float& cx, float& cy, unsigned& binrev) { ang = r->ang; cx = r->cx; cy = r->cy; binrev = r->binrev; }
OLF_HD void red_min64(u64* p, u64 v) { if (v < *p) *p = v; }
#endif

OLF_HD int tab_index(short2_t d) { return ((int)d.x + 255) * kTabDim + ((int)d.y + 255); }

// ---- chunked pixel lists (31 pixels + next pointer per 128-byte chunk) -----------------------------------------------
struct ListWriter {
    unsigned head, chunk; int off; bool overflow;
    OLF_HD void init() { head = kNull; chunk = kNull; off = kChunk - 1; overflow = false; }
    OLF_HD void push(unsigned* pool, unsigned* pool_ctr, unsigned pool_chunks, unsigned pix) {
        if (off == kChunk - 1) {
            const unsigned nc = atomic_inc32(pool_ctr);
            if (nc >= pool_chunks) { overflow = true; return; }
            pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
            if (chunk == kNull) head = nc; else pool[(size_t)chunk * kChunk + kChunk - 1] = nc;
            chunk = nc; off = 0;
        }
        pool[(size_t)chunk * kChunk + off++] = pix;
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

// ---- output of a finalised wave: regions of at least min_reg_size pixels, copied to a contiguous pool for the rectangle fit
struct LsdRegion { u64 prio; unsigned off; int count; double reg_angle; };
struct FinalOut { unsigned* final_pool; unsigned* final_ctr; LsdRegion* regs; unsigned* nreg; unsigned reg_cap; int min_reg_size; };
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
