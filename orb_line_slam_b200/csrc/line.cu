// Line half of the front end on sm_100a: LSD (cv::LineSegmentDetector, refine NONE) + LBD binary descriptors.
// Replaces Lineextractor::operator() (reference src/LineExtractor.cc:31-67), LSDDetectorC::detectImpl
// (Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:227-308, which calls the un-vendored cv LSD) and
// BinaryDescriptor::compute (binary_descriptor_custom.cpp:350-412, 539-687, 1026-1372).
//
// LSD stages: blur 7x7 sigma 0.6 -> resize x1.2 (INTER_LINEAR_EXACT) -> gradient/angle -> gradient-bin histogram ->
// seeds grouped by bin -> parallel region growing (fixed-point iteration, lsd_core.h) -> rectangle fit.
// Host steps between GPU phases: libm cos/sin of the O(#regions) rectangle angles, KeyLine construction
// (atan2, sort by response, top-N) -- all O(#segments), the per-pixel work never leaves the device.
#include "common.cuh"
#include "img_kernels.cuh"
#include "lsd_core.h"
#include "line.h"
#include <algorithm>
#include <cmath>
#include <numeric>


namespace olf {
using namespace lsd;

// ---- cv::resize INTER_LINEAR_EXACT 8UC1 -------------------------------------------------------
struct ExCoef { int s; int c0, c1; };
__global__ void k_resize_exact(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                               uint8_t* __restrict__ dst, int dw, int dh, int dpitch,
                               const ExCoef* __restrict__ cx, const ExCoef* __restrict__ cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const ExCoef a = cx[x], b = cy[y];
    const int x1 = min(a.s + 1, sw - 1), y1 = min(b.s + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)b.s * spitch;
    const uint8_t* r1 = src + (size_t)y1 * spitch;
    const unsigned t0 = (unsigned)(r0[a.s] * a.c0 + r0[x1] * a.c1);
    const unsigned t1 = (unsigned)(r1[a.s] * a.c0 + r1[x1] * a.c1);
    const unsigned acc = t0 * (unsigned)b.c0 + t1 * (unsigned)b.c1;
    dst[(size_t)y * dpitch + x] = (uint8_t)((acc + (1u << 15)) >> 16);
}

// Everything the grow kernel needs to know about a pixel sits in ONE 32-byte sector: both claim words, the level-line
// angle, the (cos, sin) the reference accumulates and the priority bin.  A neighbour test costs one sector instead of four
// scattered ones, which is what keeps many frames in flight from thrashing L2 / HBM with random 32-B accesses.
struct __align__(32) PxRec { u64 claim[2]; float ang; float cx; float cy; unsigned binrev; };
struct __align__(16) PxLo { float ang; float cx; float cy; unsigned binrev; };
__device__ __forceinline__ u64 px_claim(const PxRec* px, int q, int parity) { return __ldcg(&px[q].claim[parity]); }
__device__ __forceinline__ void px_load(const PxRec* px, int q, u64& c0, u64& c1, PxLo& lo) {
    const ulonglong2 c = __ldcg(reinterpret_cast<const ulonglong2*>(&px[q]));
    const float4 f = __ldcg(reinterpret_cast<const float4*>(&px[q]) + 1);
    c0 = c.x; c1 = c.y; lo.ang = f.x; lo.cx = f.y; lo.cy = f.z; lo.binrev = __float_as_uint(f.w);
}
#define GW_HASH_BITS 8
#define GW_HASH (1 << GW_HASH_BITS)

// ---- ll_angle: gradient, level-line angle, max gradient ---------------------------------
// n2_thresh = smallest gx^2+gy^2 whose norm sqrt(n2/4.0) exceeds rho (computed exactly on the host).
__global__ void __launch_bounds__(256) k_lsd_grad(const uint8_t* __restrict__ img, int W, int H, int pitch, int n2_thresh,
                                                  float* __restrict__ ang, short2_t* __restrict__ dabc,
                                                  const float2_t* __restrict__ tab_acc, PxRec* __restrict__ px, int* __restrict__ n2max) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int n2 = 0;
    if (x < W && y < H) {
        const int q = y * W + x;
        float a = -1.f;
        short2_t d; d.x = 0; d.y = 0;
        if (x < W - 1 && y < H - 1) {
            const uint8_t* r0 = img + (size_t)y * pitch + x;
            const uint8_t* r1 = r0 + pitch;
            const int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
            const int gx = DA + BC, gy = DA - BC;
            d.x = (short)DA; d.y = (short)BC;
            const int v = gx * gx + gy * gy;
            if (v >= n2_thresh) { a = olf::lsd::fast_atan2_deg((float)gx, (float)-gy); n2 = v; }
        }
        float2_t c; c.x = 0.f; c.y = 0.f;
        if (a >= 0.f) c = tab_acc[tab_index(d)];
        ang[q] = a; dabc[q] = d;
        PxRec r; r.claim[0] = kClaimNone; r.claim[1] = kClaimNone; r.ang = a; r.cx = c.x; r.cy = c.y; r.binrev = 0;
        px[q] = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 = max(n2, __shfl_xor_sync(0xffffffffu, n2, o));
    if ((threadIdx.x & 31) == 0 && n2 > 0) atomicMax(n2max, n2);
}

__device__ __forceinline__ int lsd_bin(short2_t d, double bin_coef) {
    const int gx = d.x + d.y, gy = d.x - d.y;
    return (int)__dmul_rn(__dsqrt_rn(__ddiv_rn((double)(gx * gx + gy * gy), 4.0)), bin_coef);
}
__device__ __forceinline__ double lsd_bin_coef(int n2max, int n_bins) {
    const double max_grad = __dsqrt_rn(__ddiv_rn((double)n2max, 4.0));
    return max_grad > 0 ? __ddiv_rn((double)(n_bins - 1), max_grad) : 0.0;
}

// histogram of gradient bins over defined pixels
__global__ void __launch_bounds__(256) k_lsd_hist(const float* __restrict__ ang, const short2_t* __restrict__ dabc, int S,
                                                  const int* __restrict__ n2max, int n_bins, unsigned* __restrict__ hist) {
    extern __shared__ unsigned sh[];
    for (int i = threadIdx.x; i < n_bins; i += 256) sh[i] = 0;
    __syncthreads();
    const double coef = lsd_bin_coef(*n2max, n_bins);
    for (int q = blockIdx.x * 256 + threadIdx.x; q < S; q += gridDim.x * 256)
        if (ang[q] >= 0.f) atomicAdd(&sh[lsd_bin(dabc[q], coef)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += 256) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one block: bin offsets (descending bin order) and the wave plan (whole bins, cumulative targets doubling)
struct LsdPlan { int n_seeds; int n_waves; int wave_start[64]; };
__global__ void __launch_bounds__(1024) k_lsd_plan(const unsigned* __restrict__ hist, int n_bins, int first_wave, int wave_growth, unsigned* __restrict__ bin_start,
                                                    unsigned* __restrict__ cursor, LsdPlan* __restrict__ plan) {
    __shared__ unsigned sh[1024], st[1024];
    for (int b = threadIdx.x; b < 1024; b += blockDim.x) sh[b] = b < n_bins ? hist[b] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        int nw = 0;
        long long target = first_wave;
        plan->wave_start[0] = 0;
        for (int b = n_bins - 1; b >= 0; --b) {
            st[b] = acc;
            acc += sh[b];
            if ((long long)acc - plan->wave_start[nw] >= target && nw < 61) { plan->wave_start[++nw] = (int)acc; target *= wave_growth; }
        }
        if (plan->wave_start[nw] != (int)acc) plan->wave_start[++nw] = (int)acc;
        plan->n_seeds = (int)acc; plan->n_waves = nw;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) { bin_start[b] = st[b]; cursor[b] = 0; }
}

__global__ void __launch_bounds__(256) k_lsd_scatter(const float* __restrict__ ang, const short2_t* __restrict__ dabc, int S,
                                                     const int* __restrict__ n2max, int n_bins, const unsigned* __restrict__ bin_start,
                                                     unsigned* __restrict__ cursor, int* __restrict__ seed_pix, u64* __restrict__ seed_prio,
                                                     PxRec* __restrict__ px) {
    const double coef = lsd_bin_coef(*n2max, n_bins);
    for (int q = blockIdx.x * 256 + threadIdx.x; q < S; q += gridDim.x * 256)
        if (ang[q] >= 0.f) {
            const int b = lsd_bin(dabc[q], coef);
            px[q].binrev = (unsigned)(n_bins - 1 - b);
            const unsigned pos = bin_start[b] + atomicAdd(&cursor[b], 1u);
            seed_pix[pos] = q;
            seed_prio[pos] = make_prio(n_bins - 1 - b, q);
        }
}

// ---- region growing state shared by the phase kernel (semantics: grow_seed() in lsd_core.h) ------------------
struct LsdRegion { u64 prio; unsigned off; int count; double reg_angle; };
struct GrowState {
    GrowArgs A;
    const int* seed_pix; const u64* seed_prio;
    unsigned* head[2]; int* cnt[2]; double* regang;
    const LsdPlan* plan;
    unsigned* changed;          // [max_rounds] zero-initialised
    unsigned max_rounds;
    int min_reg_size;
    unsigned* final_pool; unsigned* final_ctr;
    LsdRegion* regs; unsigned* nreg; unsigned reg_cap;
    int* status;                // [0] error flag, [1] rounds used, [2] waves
};

// ---- region growing, warp-cooperative: one WARP per seed per round ----------------------------------------------
// Same operator as grow_seed() in lsd_core.h (which stays the executable specification, emulated on the host by
// tests/emul), restructured so that nothing on the sequential critical path of a region is a dependent global load:
//   * up to 3 queue entries x 9 neighbours are examined per step; each of 27 lanes fetches its neighbour's angle, both
//     claim words and (cos, sin) in parallel (one L2 latency per step instead of ~30);
//   * the order-dependent accept decisions (running region angle) are then replayed in the reference's scan order on
//     warp-uniform registers, visiting only lanes whose prefetched data make them potential accepts;
//   * pixel lists live in 31-entry chunks: the chunk being written / read / compared sits in one register per lane and
//     moves to and from global memory as one coalesced 128-byte transaction;
//   * claims are atomicMin by lane 0; their return values are awaited once per step, which makes the warp's own claims
//     visible to the next step's (L2) loads -- no duplicate can enter a list.
#define GW_WARPS 4
struct GrowStateW {
    GrowState G;
    PxRec* px;                  // packed per-pixel record (claims + angle + (cos, sin) + bin)
    unsigned* work_ctr;         // [2*max_rounds + 64] zero-initialised work counters (one per round / finalise pass)
    // per seed and round parity: the aligned candidates that were refused because a NON-final higher-priority claim held
    // them (one chunk at most; count 255 = too many, always re-grow).  Together with the pixel list they are the complete
    // set of external facts a growth depended on, which is what lets an unchanged region be verified instead of re-grown.
    unsigned* blk_chunk[2]; int* blk_cnt[2];
    int fast_align; float c_hi2, c_lo2;   // lazy alignment test: cos^2(prec -/+ 0.1 deg)
    int defer;                            // first round of a wave: seeds with a live higher-priority aligned neighbour wait
    int* dbg;                   // optional per-round trace (see OLF_LSD_TRACE)
};

__device__ __forceinline__ bool blocked_vals(u64 e_prev, u64 e_cur, u64 sf_prev, u64 sf_cur, u64 prio) {
    u64 sf = e_prev >> 40;
    if (sf == 0) return true;
    if (sf == sf_prev && (e_prev & kPrioMask) < prio) return true;
    sf = e_cur >> 40;
    if (sf == 0) return true;
    if (sf == sf_cur && (e_cur & kPrioMask) <= prio) return true;
    return false;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_f64(double v, int src) { return __longlong_as_double((long long)shfl_u64((u64)__double_as_longlong(v), src)); }

// returns true if the seed's outcome differs from the previous round
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TRACE_REC 8
__device__ bool grow_seed_warp(const GrowStateW& S, unsigned round, int i, int seed, u64 prio, int lane, volatile unsigned* hs) {
    const GrowState& G = S.G;
    const GrowArgs& A = G.A;
    const int cur = round & 1, prv = (round - 1) & 1;
    const u64 sf_prev = stamp_field(round - 1), sf_cur = stamp_field(round);
    const u64 mine = (sf_cur << 40) | prio;
    PxRec* px = S.px;
    unsigned* pool = A.pool[cur];
    const unsigned* ppool = A.pool[prv];
    const unsigned prev_head = __ldcg(&G.head[prv][i]);
    const int prev_cnt = __ldcg(&G.cnt[prv][i]);
    // claim the seed pixel (authoritative check through the atomic's return value)
    int dead = 0;
    unsigned first_chunk = 0;
    if (lane == 0) {
        const u64 ep = px_claim(px, seed, prv);
        u64 sf = ep >> 40;
        dead = (sf == 0) || (sf == sf_prev && (ep & kPrioMask) < prio);
        if (!dead) {
            const u64 old = atomicMin(&px[seed].claim[cur], mine);
            sf = old >> 40;
            dead = (sf == 0) || (sf == sf_cur && (old & kPrioMask) < prio);
        }
    }
    dead = __shfl_sync(0xffffffffu, dead, 0);
    if (dead) {
        if (lane == 0) { G.cnt[cur][i] = 0; G.head[cur][i] = kNull; }
        return prev_cnt != 0;
    }
    // ---- verify instead of re-grow: last round's run is still exact iff every pixel it accepted is still free of
    // higher-priority claims and every aligned candidate it was refused (by a non-final claim) is still held.
    if (prev_cnt > 0) {
        const int bc = __ldcg(&S.blk_cnt[prv][i]);
        bool ok = bc != 255;
        if (ok) {
            unsigned chunk = prev_head;
            for (int k = 0; k < prev_cnt && ok; k += kChunk - 1) {
                const unsigned v = __ldcg(&ppool[(size_t)chunk * kChunk + lane]);
                bool bad = false;
                if (lane < kChunk - 1 && k + lane < prev_cnt) {
                    const u64 ep = px_claim(px, (int)v, prv);
                    bad = ((ep >> 40) == sf_prev) && ((ep & kPrioMask) < prio);
                }
                ok = !__any_sync(0xffffffffu, bad);
                chunk = __shfl_sync(0xffffffffu, v, kChunk - 1);
            }
        }
        if (ok && bc > 0) {
            const unsigned v = __ldcg(&ppool[(size_t)__ldcg(&S.blk_chunk[prv][i]) * kChunk + lane]);
            bool bad = false;
            if (lane < bc) {
                const u64 ep = px_claim(px, (int)v, prv);
                const u64 sf = ep >> 40;
                bad = !((sf == 0) || (sf == sf_prev && (ep & kPrioMask) < prio));      // no longer held -> must re-grow
            }
            ok = !__any_sync(0xffffffffu, bad);
        }
        if (ok) {                                                   // carry the region over: re-stamp its claims for this round
            unsigned chunk = prev_head;
            for (int k = 0; k < prev_cnt; k += kChunk - 1) {
                const unsigned v = __ldcg(&ppool[(size_t)chunk * kChunk + lane]);
                if (lane < kChunk - 1 && k + lane < prev_cnt && v != (unsigned)seed) atomicMin(&px[v].claim[cur], mine);
                chunk = __shfl_sync(0xffffffffu, v, kChunk - 1);
            }
            if (S.dbg && lane == 0) atomicAdd(&S.dbg[round * TRACE_REC + 4], 1);
            if (lane == 0) {
                G.head[cur][i] = prev_head; G.cnt[cur][i] = prev_cnt;
                S.blk_chunk[cur][i] = __ldcg(&S.blk_chunk[prv][i]); S.blk_cnt[cur][i] = bc;
            }
            return false;
        }
    }
    if (lane == 0) first_chunk = atomicAdd(A.pool_ctr[cur], 1u);
    first_chunk = __shfl_sync(0xffffffffu, first_chunk, 0);
    if (first_chunk >= A.pool_chunks) { if (lane == 0) { G.status[0] = OLF_ERR_CAPACITY; G.cnt[cur][i] = 0; G.head[cur][i] = kNull; } return true; }
    // writer / reader / comparer state (warp-uniform scalars + one payload register per lane)
    const unsigned w_head = first_chunk;
    unsigned w_chunk = first_chunk; int w_n = 0; unsigned w_val = 0;
    unsigned r_chunk = first_chunk; int r_n = 0; unsigned r_val = 0;
    unsigned pv_val = (prev_cnt > 0 && prev_head != kNull) ? __ldcg(&ppool[(size_t)prev_head * kChunk + lane]) : kNull;
    int pv_n = 0;
    bool same = prev_cnt > 0;
    int count = 0;
    bool overflow = false;
    unsigned b_chunk = kNull; int b_n = 0;                           // refused-candidate record (lazily allocated chunk)
    auto record_blocked = [&](unsigned pix) {
        if (b_n == 255) return;
        if (b_chunk == kNull) {
            unsigned nc = 0;
            if (lane == 0) nc = atomicAdd(A.pool_ctr[cur], 1u);
            nc = __shfl_sync(0xffffffffu, nc, 0);
            if (nc >= A.pool_chunks) { overflow = true; return; }
            b_chunk = nc;
        }
        if (b_n >= kChunk - 1) { b_n = 255; return; }
        if (lane == 0) pool[(size_t)b_chunk * kChunk + b_n] = pix;
        ++b_n;
    };
    auto push = [&](unsigned pix) {
        if (w_n == kChunk - 1) {                                   // flush the full chunk, chain a new one
            unsigned nc = 0;
            if (lane == 0) nc = atomicAdd(A.pool_ctr[cur], 1u);
            nc = __shfl_sync(0xffffffffu, nc, 0);
            if (nc >= A.pool_chunks) { overflow = true; return; }
            const unsigned outv = (lane == kChunk - 1) ? nc : w_val;
            pool[(size_t)w_chunk * kChunk + lane] = outv;
            if (r_chunk == w_chunk) r_val = outv;                  // the reader keeps the chunk it is still consuming
            w_chunk = nc; w_n = 0;
        }
        if (lane == w_n) w_val = pix;
        ++w_n;
        if (same) {                                                // lock-step comparison with last round's list
            if (count >= prev_cnt) same = false;
            else {
                if (__shfl_sync(0xffffffffu, pv_val, pv_n) != pix) same = false;
                if (++pv_n == kChunk - 1) {
                    const unsigned nxt = __shfl_sync(0xffffffffu, pv_val, kChunk - 1);
                    pv_val = (nxt != kNull) ? __ldcg(&ppool[(size_t)nxt * kChunk + lane]) : kNull;
                    pv_n = 0;
                }
            }
        }
        ++count;
    };
    push((unsigned)seed);
    double reg_angle = d_mul((double)__ldg(&A.ang[seed]), kDegToRads);
    const float2_t t0 = A.tab_seed[tab_index(A.dabc[seed])];
    float sumdx = t0.x, sumdy = t0.y;
    float u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
    bool dirty = false;                                              // reg_angle is stale w.r.t. (sumdx, sumdy)
    // ---- software-pipelined BFS: while the accept decisions of step k are replayed, the neighbour data of step k+1 are
    // already in flight.  Claims are fire-and-forget reductions (RED): nothing on the critical path waits for L2.  The warp's
    // own accepts since the last fence live in a small shared-memory hash set `hs`; loads are issued only after a fence has
    // made every older own claim visible, so "already mine" is always decidable from (loaded claim word) OR (hash set hit).
    struct Pf { int q; float aq; u64 ep, ec; float cx, cy; };
    int n_since_fence = 0;
    int i_issue = 0;                                                 // queue entries whose neighbour loads have been issued
    for (int k = lane; k < GW_HASH; k += 32) hs[k] = kNull;
    __syncwarp();
    auto hs_insert = [&](unsigned pix) {                             // lane 0 only
        unsigned slot = (pix * 2654435761u) >> (32 - GW_HASH_BITS);
        while (hs[slot] != kNull) slot = (slot + 1) & (GW_HASH - 1);
        hs[slot] = pix;
    };
    auto hs_contains = [&](unsigned pix) -> bool {
        unsigned slot = (pix * 2654435761u) >> (32 - GW_HASH_BITS);
        for (;;) {
            const unsigned v = hs[slot];
            if (v == pix) return true;
            if (v == kNull) return false;
            slot = (slot + 1) & (GW_HASH - 1);
        }
    };
    auto issue = [&](int nb) -> Pf {
        int ent[3] = {-1, -1, -1};
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            if (e < nb) {
                if (r_n == kChunk - 1) {                           // lazily step to the next chunk (the old one is flushed by now)
                    r_chunk = __shfl_sync(0xffffffffu, r_val, kChunk - 1); r_n = 0;
                    if (r_chunk != w_chunk) r_val = __ldcg(&pool[(size_t)r_chunk * kChunk + lane]);
                }
                ent[e] = (int)((r_chunk == w_chunk) ? __shfl_sync(0xffffffffu, w_val, r_n) : __shfl_sync(0xffffffffu, r_val, r_n));
                ++r_n;
            }
        }
        // lane -> (entry, neighbour) in the reference's scan order: yy outer, xx inner
        const int e = lane / 9, nidx = lane - e * 9;
        const int p = e == 0 ? ent[0] : (e == 1 ? ent[1] : ent[2]);
        Pf f; f.q = -1 - lane; f.aq = -1.f; f.ep = kClaimNone; f.ec = kClaimNone; f.cx = 0.f; f.cy = 0.f;
        if (lane < 27 && e < nb) {
            const int ey = p / A.W, ex = p - ey * A.W;
            const int xx = ex + (nidx % 3) - 1, yy = ey + (nidx / 3) - 1;
            if (xx >= 0 && xx < A.W && yy >= 0 && yy < A.H) {
                f.q = yy * A.W + xx;
                u64 c0, c1; PxLo lo;
                px_load(px, f.q, c0, c1, lo);
                f.ep = prv ? c1 : c0; f.ec = cur ? c1 : c0;
                f.aq = lo.ang; f.cx = lo.cx; f.cy = lo.cy;
            }
        }
        i_issue += nb;
        return f;
    };
    Pf cf = issue(1);                                                // step 0: the seed
    bool have_cur = true;
#ifdef OLF_LSD_PROFILE
#define PCLK() clock64()
#else
#define PCLK() 0ll
#endif
    long long tA = 0, tB0 = 0, tB1 = 0, tC = 0, tE = 0, tS = 0, tL = 0, tP = 0; int n_steps = 0, n_pot = 0, n_pipe = 0, n_accepts = 0;
    while (have_cur && !overflow) {
        const long long c0 = PCLK();
        // [A] issue the next step's loads if the queue already holds its entries (not across a pending fence)
        const bool need_fence = n_since_fence > GW_HASH / 2 - 32;
        Pf nf; bool have_next = false;
        if (count - i_issue > 0 && !need_fence) {
            nf = issue(min(3, count - i_issue));
            have_next = true;
        }
        const long long c1 = PCLK();
        // [B] replay the accept decisions of the current step in scan order.  Every lane evaluates isAligned() for its own
        // candidate against the warp-uniform region state; the lowest aligned lane is the next pixel the reference would take;
        // after an accept the state changes and the remaining (higher) lanes are re-evaluated.  Candidates that are not aligned
        // at their turn never come back (lanes below the last decision are dropped), exactly as in the sequential scan.
        bool pot = false, held = false;
        if (cf.q >= 0 && cf.aq >= 0.f) {
            const u64 sp = cf.ep >> 40, sc = cf.ec >> 40;
            const bool fin = sp == 0 || sc == 0;                                            // finalised region: never comes back
            const bool own = sc == sf_cur && (cf.ec & kPrioMask) == prio;                   // already in this region
            held = (sp == sf_prev && (cf.ep & kPrioMask) < prio) || (sc == sf_cur && (cf.ec & kPrioMask) < prio);
            pot = !fin && !own && !hs_contains((unsigned)cf.q);
        }
        const long long c2 = PCLK();
        n_pot += __popc(__ballot_sync(0xffffffffu, pot)); ++n_steps; n_pipe += have_next;
        const double a_rad = d_mul((double)cf.aq, kDegToRads);
        unsigned live = 0xffffffffu;                                                        // lanes not yet passed by the scan
        bool state_changed = true;
        unsigned al = 0;
        for (;;) {
            const long long e0 = PCLK();
            if (state_changed) {
                // isAligned(reg_angle, a) with reg_angle = fastAtan2(sumdy, sumdx) evaluated lazily: the sign
                // of cos(angular distance) - cos(prec -/+ 0.1 deg) decides all but the candidates within 0.1 deg of the threshold
                int st = 0;
                if (pot && ((live >> lane) & 1u)) {
                    const float dot = f_add(f_mul(sumdx, cf.cx), f_mul(sumdy, cf.cy)), d2 = f_mul(dot, dot);
                    if (S.fast_align && u2 > 1e-3f && dot > 0.f && d2 >= f_mul(S.c_hi2, u2)) st = 1;
                    else if (S.fast_align && u2 > 1e-3f && (dot <= 0.f || d2 <= f_mul(S.c_lo2, u2))) st = 0;
                    else st = 2;
                }
                if (__any_sync(0xffffffffu, st == 2)) {
                    if (dirty) { reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads); dirty = false; }
                    if (st == 2) {
                        double n_theta = d_sub(reg_angle, a_rad);
                        if (n_theta < 0) n_theta = -n_theta;
                        if (n_theta > k3_2Pi) { n_theta = d_sub(n_theta, k2Pi); if (n_theta < 0) n_theta = -n_theta; }
                        st = (n_theta <= A.prec) ? 1 : 0;
                    }
                }
                al = __ballot_sync(0xffffffffu, st == 1);
                state_changed = false;
            }
            al &= live;
            const long long e1 = PCLK(); tE += e1 - e0;
            if (!al) break;
            const int c = __ffs(al) - 1;
            live = (c == 31) ? 0u : (0xffffffffu << (c + 1));                               // the scan has passed lanes <= c
            const int qc = __shfl_sync(0xffffffffu, cf.q, c);
            if (__shfl_sync(0xffffffffu, (int)held, c)) { record_blocked((unsigned)qc); if (overflow) break; continue; }   // aligned but held by a higher-priority seed
            const float cx = __shfl_sync(0xffffffffu, cf.cx, c), cy = __shfl_sync(0xffffffffu, cf.cy, c);
            if (cf.q == qc) pot = false;                                                    // the same pixel seen from another entry
            const long long e2 = PCLK(); tS += e2 - e1;
            if (lane == 0) { atomicMin(&px[qc].claim[cur], mine); hs_insert((unsigned)qc); }   // RED: no return value is consumed
            ++n_since_fence;
            const long long e3 = PCLK(); tL += e3 - e2;
            push((unsigned)qc);
            const long long e4 = PCLK(); tP += e4 - e3; ++n_accepts;
            if (overflow) break;
            sumdx = f_add(sumdx, cx);
            sumdy = f_add(sumdy, cy);
            u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
            dirty = true;
            state_changed = true;
        }
        __syncwarp();                                                             // hash-set inserts visible to all lanes
        const long long c3 = PCLK();
        // [C] rotate; if nothing was prefetched, fence when due (every own claim reaches L2, the hash set restarts) and issue now
        if (have_next) cf = nf;
        else if (count - i_issue > 0 && !overflow) {
            if (need_fence) {
                __threadfence();
                for (int k = lane; k < GW_HASH; k += 32) hs[k] = kNull;
                __syncwarp();
                n_since_fence = 0;
            }
            cf = issue(min(3, count - i_issue));
        }
        else have_cur = false;
        const long long c4 = PCLK();
        tA += c1 - c0; tB0 += c2 - c1; tB1 += c3 - c2; tC += c4 - c3;
    }
#ifdef OLF_LSD_PROFILE
    if (S.dbg && lane == 0 && count >= 900) {
        int* d = S.dbg + 200 * TRACE_REC;
        d[0] = count; d[1] = n_steps; d[2] = n_pot; d[3] = n_pipe; d[4] = (int)tA; d[5] = (int)tB0; d[6] = (int)tB1; d[7] = (int)tC;
        d[8] = n_accepts; d[9] = (int)tE; d[10] = (int)tS; d[11] = (int)tL; d[12] = (int)tP;
    }
#endif
    const unsigned pend = 0;
    if (dirty) reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads);
#ifdef OLF_LSD_PROFILE
    if (S.dbg && lane == 0 && count >= 900) {
        int* d = S.dbg + 200 * TRACE_REC;
        d[0] = count; d[1] = n_steps; d[2] = n_pot; d[3] = n_pipe; d[4] = (int)tA; d[5] = (int)tB0; d[6] = (int)tB1; d[7] = (int)tC;
        d[8] = n_accepts; d[9] = (int)tE; d[10] = (int)tS; d[11] = (int)tL; d[12] = (int)tP;
    }
#endif
    if (overflow) { if (lane == 0) { G.status[0] = OLF_ERR_CAPACITY; G.cnt[cur][i] = 0; G.head[cur][i] = kNull; } return true; }
    // flush the partial chunk
    if (lane < w_n || lane == kChunk - 1) pool[(size_t)w_chunk * kChunk + lane] = (lane == kChunk - 1) ? kNull : w_val;
    if (lane == 0) { G.head[cur][i] = w_head; G.cnt[cur][i] = count; G.regang[i] = reg_angle; S.blk_chunk[cur][i] = b_chunk; S.blk_cnt[cur][i] = b_n; }
    if (S.dbg && lane == 0) { atomicAdd(&S.dbg[round * TRACE_REC + 5], 1); atomicAdd(&S.dbg[round * TRACE_REC + 6], count); atomicMax(&S.dbg[round * TRACE_REC + 7], count); }
    return !(same && count == prev_cnt) || (pend == 0xFFFFFFFFu);
}

// One launch = one phase of the wave/round state machine (ROUND: every live seed of the wave verifies or re-grows;
// FINALIZE: the converged wave is stamped for good).  The last block to finish advances the state, so a fixed batch of
// back-to-back launches walks through all phases without any host round trip; launches after `done` return at once.
// Ordinary (non-cooperative) launches: no co-residency requirement, so the grow phases of the left/right eyes and of
// several frames in flight interleave freely with every other kernel on the device.
struct PhaseState { int wave; unsigned round; int mode; int done; unsigned pass; unsigned ticket; int launches; unsigned wave_first_round; };

__global__ void __launch_bounds__(GW_WARPS * 32, 6) k_lsd_phase(const GrowStateW S, PhaseState* __restrict__ st) {
    __shared__ unsigned hash_sets[GW_WARPS][GW_HASH];
    __shared__ bool s_last;
    const GrowState& G = S.G;
    const int wv = st->wave; const unsigned round = st->round; const int mode = st->mode; const unsigned pass = st->pass;
    const bool first_round = round == st->wave_first_round;
    if (st->done) return;
    if (wv >= G.plan->n_waves) {                                   // no seeds at all (flat image)
        if (blockIdx.x == 0 && threadIdx.x == 0) { st->done = 1; G.status[1] = (int)round; G.status[2] = G.plan->n_waves; G.status[3] = 1; }
        return;
    }
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int lo = G.plan->wave_start[wv], hi = G.plan->wave_start[wv + 1];
    const int cur = round & 1, prv = (round - 1) & 1;
    if (mode == 0) {
        const u64 sf_prev = stamp_field(round - 1), sf_cur = stamp_field(round);
        bool any_change = false;
        if (S.dbg && tid == 0) { S.dbg[round * TRACE_REC + 0] = wv; S.dbg[round * TRACE_REC + 1] = hi - lo; S.dbg[round * TRACE_REC + 2] = (int)(gtime() & 0x7fffffff); }
        // seeds are handed out in batches of `bsz` (one per lane for the cheap liveness test); small waves use small
        // batches so that the live seeds of a batch do not queue up behind each other inside one warp
        const int nwarps = (int)(nth >> 5);
        const int bsz = max(1, min(32, (hi - lo) / (2 * nwarps)));
        for (;;) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&S.work_ctr[pass], (unsigned)bsz);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lo + (int)base >= hi) break;
            const int i = lo + (int)base + lane;
            int seed = 0; u64 prio = 0; bool alive = false, deferred = false;
            if (i < hi && lane < bsz) {
                seed = G.seed_pix[i]; prio = G.seed_prio[i];
                alive = !blocked_vals(px_claim(S.px, seed, prv), px_claim(S.px, seed, cur), sf_prev, sf_cur, prio);
                if (alive && first_round && S.defer) {
                    // Work saver (does not change the fixed point): in the first round of a wave a seed that has a live,
                    // higher-priority, aligned 8-neighbour will almost surely be absorbed by that neighbour's region, so it
                    // sits this round out; from the second round on every live seed grows as usual.
                    const int W = G.A.W, H = G.A.H, py = seed / W, px = seed - py * W;
                    const float a_s = __ldg(&G.A.ang[seed]);
                    for (int dy = -1; dy <= 1 && alive; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int xx = px + dx, yy = py + dy;
                            if ((dx | dy) == 0 || xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
                            const int q = yy * W + xx;
                            u64 c0, c1; PxLo lo;
                            px_load(S.px, q, c0, c1, lo);
                            const float a_q = lo.ang;
                            if (a_q < 0.f) continue;
                            const u64 pq = make_prio((int)lo.binrev, q);
                            if (pq >= prio) continue;
                            if (((prv ? c1 : c0) >> 40) == 0) continue;                       // finalised long ago: cannot absorb us now
                            float d = fabsf(a_s - a_q); if (d > 180.f) d = 360.f - d;
                            if (d <= 20.f) { alive = false; deferred = true; break; }
                        }
                }
                if (!alive) { if (deferred || __ldcg(&G.cnt[prv][i]) != 0) any_change = true; G.cnt[cur][i] = 0; G.head[cur][i] = kNull; }
            }
            unsigned m = __ballot_sync(0xffffffffu, alive);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int s_seed = __shfl_sync(0xffffffffu, seed, src);
                const u64 s_prio = shfl_u64(prio, src);
                if (grow_seed_warp(S, round, lo + (int)base + src, s_seed, s_prio, lane, hash_sets[threadIdx.x >> 5])) any_change = true;
            }
        }
        if (__syncthreads_or(any_change) && threadIdx.x == 0) G.changed[round] = 1;
    } else {
        // finalise the wave (warp per live seed): stamp the regions for good, keep the lists of accepted regions
        const unsigned* pool = G.A.pool[cur];
        for (;;) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&S.work_ctr[pass], 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lo + (int)base >= hi) break;
            const int i = lo + (int)base + lane;
            const int c_l = (i < hi) ? __ldcg(&G.cnt[cur][i]) : 0;
            unsigned m = __ballot_sync(0xffffffffu, c_l > 0);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int si = lo + (int)base + src;
                const int c = __shfl_sync(0xffffffffu, c_l, src);
                const u64 prio = G.seed_prio[si];
                const bool accept = c >= G.min_reg_size;
                unsigned off = 0;
                if (accept && lane == 0) off = atomicAdd(G.final_ctr, (unsigned)c);
                off = __shfl_sync(0xffffffffu, off, 0);
                unsigned chunk = __ldcg(&G.head[cur][si]);
                for (int k = 0; k < c; k += kChunk - 1) {
                    const unsigned v = __ldcg(&pool[(size_t)chunk * kChunk + lane]);
                    if (lane < kChunk - 1 && k + lane < c) {
                        S.px[v].claim[0] = prio; S.px[v].claim[1] = prio;
                        if (accept) G.final_pool[off + k + lane] = v;
                    }
                    chunk = __shfl_sync(0xffffffffu, v, kChunk - 1);
                }
                if (accept && lane == 0) {
                    const unsigned r = atomicAdd(G.nreg, 1u);
                    if (r < G.reg_cap) { LsdRegion R; R.prio = prio; R.off = off; R.count = c; R.reg_angle = __ldcg(&G.regang[si]); G.regs[r] = R; }
                    else G.status[0] = OLF_ERR_CAPACITY;
                }
            }
        }
    }
    // last block to finish advances the state machine
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    const int err = __ldcg(&G.status[0]);
    st->ticket = 0; st->launches += 1; st->pass = pass + 1;
    if (mode == 0) {
        if (S.dbg) S.dbg[round * TRACE_REC + 3] = (int)(gtime() & 0x7fffffff);
        const bool changed = __ldcg(&G.changed[round]) != 0;
        if (!changed || round + 2 >= G.max_rounds || err != 0) st->mode = 1; else st->round = round + 1;
    } else {
        st->mode = 0; st->round = round + 1; st->wave = wv + 1; st->wave_first_round = round + 1;
        *G.A.pool_ctr[0] = 0;                                       // one bump pool per wave (lists may be carried over rounds)
        if (wv + 1 >= G.plan->n_waves || err != 0) { st->done = 1; G.status[1] = (int)(round + 1); G.status[2] = G.plan->n_waves; G.status[3] = 1; }
    }
    __threadfence();
}

// ---- region2rect ---------------------------------------------------------------------------
struct RectRec { double x, y, theta; u64 prio; };
// Warp per region.  The reference's sums are sequential double additions in BFS order (not associative), so the additions
// stay a single chain -- but everything that feeds them (pixel fetch, gradient weight sqrt/div, products) is computed by
// the 32 lanes in parallel and staged in shared memory; lane 0 only walks the three add chains.
__global__ void __launch_bounds__(256) k_lsd_rect_a(const LsdRegion* __restrict__ regs, const unsigned* __restrict__ nreg, unsigned cap,
                                                    const unsigned* __restrict__ final_pool, const short2_t* __restrict__ dabc, int W,
                                                    double prec, RectRec* __restrict__ out_host) {
    __shared__ double sh[8][3][32];
    const unsigned n = min(*nreg, cap);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (unsigned r = blockIdx.x * 8 + warp; r < n; r += gridDim.x * 8) {
        const LsdRegion R = regs[r];
        const unsigned* pix = final_pool + R.off;
        // pass 1: centroid weighted by the gradient norm
        double x = 0, y = 0, sum = 0;
        for (int base = 0; base < R.count; base += 32) {
            const int k = base + lane;
            if (k < R.count) {
                const int p = (int)pix[k];
                const short2_t d = dabc[p];
                const int gx = d.x + d.y, gy = d.x - d.y;
                const double w = d_sqrt(d_div((double)(gx * gx + gy * gy), 4.0));
                sh[warp][0][lane] = d_mul((double)(p % W), w);
                sh[warp][1][lane] = d_mul((double)(p / W), w);
                sh[warp][2][lane] = w;
            }
            __syncwarp();
            if (lane == 0) {
                const int m = min(32, R.count - base);
                for (int j = 0; j < m; ++j) { x = d_add(x, sh[warp][0][j]); y = d_add(y, sh[warp][1][j]); sum = d_add(sum, sh[warp][2][j]); }
            }
            __syncwarp();
        }
        x = shfl_f64(d_div(x, sum), 0); y = shfl_f64(d_div(y, sum), 0);
        // pass 2: inertia matrix
        double Ixx = 0, Iyy = 0, Ixy = 0;
        for (int base = 0; base < R.count; base += 32) {
            const int k = base + lane;
            if (k < R.count) {
                const int p = (int)pix[k];
                const short2_t d = dabc[p];
                const int gx = d.x + d.y, gy = d.x - d.y;
                const double w = d_sqrt(d_div((double)(gx * gx + gy * gy), 4.0));
                const double dx = d_sub((double)(p % W), x), dy = d_sub((double)(p / W), y);
                sh[warp][0][lane] = d_mul(d_mul(dy, dy), w);
                sh[warp][1][lane] = d_mul(d_mul(dx, dx), w);
                sh[warp][2][lane] = d_mul(d_mul(dx, dy), w);
            }
            __syncwarp();
            if (lane == 0) {
                const int m = min(32, R.count - base);
                for (int j = 0; j < m; ++j) { Ixx = d_add(Ixx, sh[warp][0][j]); Iyy = d_add(Iyy, sh[warp][1][j]); Ixy = d_sub(Ixy, sh[warp][2][j]); }
            }
            __syncwarp();
        }
        if (lane == 0) {
            const double dI = d_sub(Ixx, Iyy);
            const double lambda = d_mul(0.5, d_sub(d_add(Ixx, Iyy), d_sqrt(d_add(d_mul(dI, dI), d_mul(d_mul(4.0, Ixy), Ixy)))));
            double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)olf::lsd::fast_atan2_deg((float)d_sub(lambda, Ixx), (float)Ixy)
                                                    : (double)olf::lsd::fast_atan2_deg((float)Ixy, (float)d_sub(lambda, Iyy));
            theta = d_mul(theta, kDegToRads);
            if (angle_diff(theta, R.reg_angle) > prec) theta = d_add(theta, M_PI);
            RectRec o; o.x = x; o.y = y; o.theta = theta; o.prio = R.prio;
            out_host[r] = o;
        }
    }
}
// warp per region: extreme projections on the (host-libm) direction -> segment end points (Vec4f)
__global__ void __launch_bounds__(256) k_lsd_rect_b(const LsdRegion* __restrict__ regs, int n, const unsigned* __restrict__ final_pool, int W,
                                                    const RectRec* __restrict__ rect, const double2* __restrict__ dir, double scale,
                                                    float4* __restrict__ seg_host) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n) return;
    const LsdRegion R = regs[r];
    const double cx = rect[r].x, cy = rect[r].y, dx = dir[r].x, dy = dir[r].y;
    double l_min = 0, l_max = 0;
    for (int k = lane; k < R.count; k += 32) {
        const double l = region_proj(final_pool[R.off + k], W, cx, cy, dx, dy);
        l_max = fmax(l_max, l); l_min = fmin(l_min, l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l_max = fmax(l_max, __shfl_xor_sync(0xffffffffu, l_max, o));
        l_min = fmin(l_min, __shfl_xor_sync(0xffffffffu, l_min, o));
    }
    if (lane == 0) {
        double v[4] = {__dadd_rn(cx, __dmul_rn(l_min, dx)), __dadd_rn(cy, __dmul_rn(l_min, dy)),
                       __dadd_rn(cx, __dmul_rn(l_max, dx)), __dadd_rn(cy, __dmul_rn(l_max, dy))};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = __dadd_rn(v[k], 0.5); if (scale != 1.0) v[k] = __ddiv_rn(v[k], scale); }
        seg_host[r] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    }
}

// ---- LBD (binary_descriptor_custom.cpp:350-412, 1026-1372) ---------------------------------------------------------
// cv::Sobel 8U->16S ksize 3, BORDER_REFLECT_101; dx and dy packed as short2 per pixel
__global__ void __launch_bounds__(256) k_sobel3(const uint8_t* __restrict__ img, int w, int h, int pitch, short2_t* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const uint8_t* a = img + (size_t)reflect101(y - 1, h) * pitch;
    const uint8_t* b = img + (size_t)y * pitch;
    const uint8_t* c = img + (size_t)reflect101(y + 1, h) * pitch;
    short2_t o;
    o.x = (short)((a[xp] - a[xm]) + 2 * (b[xp] - b[xm]) + (c[xp] - c[xm]));
    o.y = (short)((c[xm] + 2 * c[x] + c[xp]) - (a[xm] + 2 * a[x] + a[xp]));
    out[(size_t)y * w + x] = o;
}

struct LbdLine { float sx, sy, ex, ey, dL0, dL1; int num_px; };      // dL = (float)cos/sin((double)angle), host libm
__constant__ float c_gaussG[63];
__constant__ float c_gaussL[21];

// thread per (line, row of the line support region): the reference's sequential float sums along the row (:1146-1186)
__global__ void __launch_bounds__(256) k_lbd_rows(const LbdLine* __restrict__ lines, int n, const short2_t* __restrict__ grad,
                                                  int imw, int imh, float4* __restrict__ rowsum) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 63) return;
    const int li = t / 63, hID = t % 63;
    const LbdLine L = lines[li];
    const short imageWidth = (short)(imw - 1), imageHeight = (short)(imh - 1);
    const short lengthOfLSP = (short)L.num_px;
    const short halfWidth = (lengthOfLSP - 1) / 2, halfHeight = 31;
    const float midX = fmul(fadd(L.sx, L.ex), 0.5f), midY = fmul(fadd(L.sy, L.ey), 0.5f);
    const float dO0 = -L.dL1, dO1 = L.dL0;
    float sCorX0 = fadd(fadd(fmul(-L.dL0, (float)halfWidth), fmul(L.dL1, (float)halfHeight)), midX);
    float sCorY0 = fadd(fsub(fmul(-L.dL1, (float)halfWidth), fmul(L.dL0, (float)halfHeight)), midY);
    for (int r = 0; r < hID; ++r) { sCorX0 = fsub(sCorX0, L.dL1); sCorY0 = fadd(sCorY0, L.dL0); }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pL = 0, nL = 0, pO = 0, nO = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
        short tc = (short)(int)roundf(sCorX);
        const short xCor = (tc < 0) ? 0 : (tc > imageWidth) ? imageWidth : tc;
        tc = (short)(int)roundf(sCorY);
        const short yCor = (tc < 0) ? 0 : (tc > imageHeight) ? imageHeight : tc;
        const short2_t g = grad[(size_t)yCor * imw + xCor];
        const float gDL = fadd(fmul((float)g.x, L.dL0), fmul((float)g.y, L.dL1));
        const float gDO = fadd(fmul((float)g.x, dO0), fmul((float)g.y, dO1));
        if (gDL > 0) pL = fadd(pL, gDL); else nL = fsub(nL, gDL);
        if (gDO > 0) pO = fadd(pO, gDO); else nO = fsub(nO, gDO);
        sCorX = fadd(sCorX, L.dL0);
        sCorY = fadd(sCorY, L.dL1);
    }
    rowsum[t] = make_float4(pL, nL, pO, nO);
}

__constant__ unsigned char c_lbd_comb[64];
// thread per line: fold the 63 rows into 9 bands in reference order, mean/std, normalise, clamp, binarise (:1188-1341, :401-412)
__global__ void __launch_bounds__(64) k_lbd_fold(const float4* __restrict__ rowsum, int n, uint8_t* __restrict__ desc_host) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n) return;
    float band[8][9];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int b = 0; b < 9; ++b) band[k][b] = 0.f;
    auto fold = [&](int b, float c, float pL, float nL, float pL2, float nL2, float pO, float nO, float pO2, float nO2) {
        const float cc = fmul(c, c);
        band[0][b] = fadd(band[0][b], fmul(c, pL));  band[1][b] = fadd(band[1][b], fmul(c, nL));
        band[2][b] = fadd(band[2][b], fmul(cc, pL2)); band[3][b] = fadd(band[3][b], fmul(cc, nL2));
        band[4][b] = fadd(band[4][b], fmul(c, pO));  band[5][b] = fadd(band[5][b], fmul(c, nO));
        band[6][b] = fadd(band[6][b], fmul(cc, pO2)); band[7][b] = fadd(band[7][b], fmul(cc, nO2));
    };
#pragma unroll 1
    for (int hID = 0; hID < 63; ++hID) {
        const float4 rs = rowsum[(size_t)li * 63 + hID];
        const float cg_ = c_gaussG[hID];
        const float pL = fmul(cg_, rs.x), nL = fmul(cg_, rs.y), pO = fmul(cg_, rs.z), nO = fmul(cg_, rs.w);
        const float pL2 = fmul(pL, pL), nL2 = fmul(nL, nL), pO2 = fmul(pO, pO), nO2 = fmul(nO, nO);
        const int b = hID / 7, m = hID % 7;
        // dynamic band index -> keep the register array addressable with a switch-free loop
#pragma unroll
        for (int bb = 0; bb < 9; ++bb) {
            if (bb == b) fold(bb, c_gaussL[m + 7], pL, nL, pL2, nL2, pO, nO, pO2, nO2);
        }
#pragma unroll
        for (int bb = 0; bb < 9; ++bb) {
            if (bb == b - 1) fold(bb, c_gaussL[m + 14], pL, nL, pL2, nL2, pO, nO, pO2, nO2);
        }
#pragma unroll
        for (int bb = 0; bb < 9; ++bb) {
            if (bb == b + 1) fold(bb, c_gaussL[m], pL, nL, pL2, nL2, pO, nO, pO2, nO2);
        }
    }
    float des[72];
    const float invN2 = (float)(1.0 / 14.0), invN3 = (float)(1.0 / 21.0);
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const float invN = (b == 0 || b == 8) ? invN2 : invN3;
        float t = fmul(band[0][b], invN);
        des[b * 8] = t;     des[b * 8 + 4] = __fsqrt_rn(fsub(fmul(band[2][b], invN), fmul(t, t)));
        t = fmul(band[1][b], invN);
        des[b * 8 + 1] = t; des[b * 8 + 5] = __fsqrt_rn(fsub(fmul(band[3][b], invN), fmul(t, t)));
        t = fmul(band[4][b], invN);
        des[b * 8 + 2] = t; des[b * 8 + 6] = __fsqrt_rn(fsub(fmul(band[6][b], invN), fmul(t, t)));
        t = fmul(band[5][b], invN);
        des[b * 8 + 3] = t; des[b * 8 + 7] = __fsqrt_rn(fsub(fmul(band[7][b], invN), fmul(t, t)));
    }
    float tM = 0, tS = 0;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tM = fadd(tM, fmul(des[b * 8 + k], des[b * 8 + k]));
#pragma unroll
        for (int k = 4; k < 8; ++k) tS = fadd(tS, fmul(des[b * 8 + k], des[b * 8 + k]));
    }
    tM = fdiv(1.f, __fsqrt_rn(tM));
    tS = fdiv(1.f, __fsqrt_rn(tS));
#pragma unroll
    for (int b = 0; b < 9; ++b) {
#pragma unroll
        for (int k = 0; k < 4; ++k) des[b * 8 + k] = fmul(des[b * 8 + k], tM);
#pragma unroll
        for (int k = 4; k < 8; ++k) des[b * 8 + k] = fmul(des[b * 8 + k], tS);
    }
#pragma unroll
    for (int i = 0; i < 72; ++i) if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float t = 0;
#pragma unroll
    for (int i = 0; i < 72; ++i) t = fadd(t, fmul(des[i], des[i]));
    t = fdiv(1.f, __fsqrt_rn(t));
#pragma unroll
    for (int i = 0; i < 72; ++i) des[i] = fmul(des[i], t);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const int a = c_lbd_comb[2 * c], b = c_lbd_comb[2 * c + 1];
        unsigned r = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // a, b are compile-time unknown: select through unrolled compares to keep des[] in registers
            float fa = 0.f, fb = 0.f;
#pragma unroll
            for (int bb = 0; bb < 9; ++bb) { if (bb == a) fa = des[bb * 8 + i]; if (bb == b) fb = des[bb * 8 + i]; }
            if (fa > fb) r += (1u << i);
        }
        desc_host[(size_t)li * 32 + c] = (uint8_t)r;
    }
}

// ======================================================================================================
// host side
// ======================================================================================================
static const unsigned char LBD_COMB[64] = {0,1, 0,2, 0,3, 0,4, 0,5, 0,6, 1,2, 1,3, 1,4, 1,5, 1,6, 2,3, 2,4, 2,5, 2,6, 2,7,
                                           2,8, 3,4, 3,5, 3,6, 3,7, 3,8, 4,5, 4,6, 4,7, 4,8, 5,6, 5,7, 5,8, 6,7, 6,8, 7,8};

struct LineImpl {
    int device = 0;
    olf_line_params P;
    cudaStream_t stream = nullptr;
    // derived LSD constants
    double prec, rho; int n2_thresh; int blur_k; int blur_q[4];
    // size-dependent
    int img_w = 0, img_h = 0, ipitch = 0, W = 0, H = 0, wpitch = 0, S = 0, min_reg_size = 0;
    PinBuf<uint8_t> img_stage;
    DevBuf<uint8_t> img, blurred, scaled, lbd_blur;
    DevBuf<ExCoef> coef; size_t coef_y_off = 0;
    DevBuf<float> ang; DevBuf<short2_t> dabc;
    DevBuf<u64> claim0, claim1, seed_prio;
    DevBuf<int> seed_pix, cnt0, cnt1, n2max, status;
    DevBuf<unsigned> head0, head1, hist, bin_start, cursor, pool0, pool1, ctrs, changed, final_pool;
    DevBuf<double> regang;
    DevBuf<LsdPlan> plan;
    DevBuf<LsdRegion> regs;
    DevBuf<float2_t> tab_seed, tab_acc, cs;
    DevBuf<unsigned> work_ctr, blk_chunk0, blk_chunk1;
    DevBuf<PhaseState> phase;
    DevBuf<PxRec> px;
    int phase_batch = 40;
    DevBuf<int> blk_cnt0, blk_cnt1;
    bool scalar_grow = false, trace = false;
    int first_wave = 4096, wave_growth = 16;
    DevBuf<int> dbg;
    unsigned pool_chunks = 0, reg_cap = 0, max_rounds = 4096;
    int grow_blocks = 0;
    PinBuf<RectRec> rect_host; PinBuf<double2> dir_host; PinBuf<float4> seg_host; PinBuf<int> status_host; PinBuf<unsigned> nreg_host;
    // LBD
    DevBuf<short2_t> grad;
    PinBuf<LbdLine> lbd_lines; DevBuf<float4> rowsum; PinBuf<uint8_t> desc_host;
    int lbd_cap = 0;
    cudaEvent_t ev_grow0 = nullptr, ev_grow1 = nullptr;
    int last_stats[8] = {0};
};

static void build_exact_coefs(int src, int dst, double scale, std::vector<ExCoef>& out) {
    out.resize(dst);
    for (int d = 0; d < dst; ++d) {
        const double s = (d + 0.5) / scale - 0.5;
        if (s < 0) { out[d] = {0, 256, 0}; continue; }
        if (s >= src - 1) { out[d] = {src - 1, 256, 0}; continue; }
        const int o = (int)std::floor(s);
        const int c1 = (int)lrint((s - o) * 256.0);
        out[d] = {o, 256 - c1, c1};
    }
}
static void gauss_kernel_q8(int n, double sigma, int* q) {
    std::vector<double> k(n);
    double sum = 0, c = (n - 1) * 0.5;
    for (int i = 0; i < n; ++i) { const double x = i - c; k[i] = std::exp(-(x * x) / (2.0 * sigma * sigma)); sum += k[i]; }
    double carry = 0; int acc = 0;
    for (int i = 0; i < n / 2; ++i) {
        const double adj = k[i] / sum * 256.0 + carry;
        const int v = (int)lrint(adj);
        carry = adj - v; q[i] = v; acc += 2 * v;
    }
    q[n / 2] = 256 - acc;
}

LineImpl* line_create(const olf_line_params* p, int device) {
    if (!p || p->lsd_refine != 0 || p->lsd_n_bins < 1 || p->lsd_n_bins > 1024 || p->lsd_scale <= 0 || p->lsd_ang_th <= 0 || p->lsd_ang_th >= 180) {
        set_last_error("olf_line_create: unsupported parameters (refine must be 0, n_bins <= 1024)"); return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_last_error("olf_line_create: no such CUDA device (this library has no CPU path)"); return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return nullptr; }
    LineImpl* h = new LineImpl();
    h->device = device; h->P = *p;
    h->prec = M_PI * p->lsd_ang_th / 180.0;
    h->rho = p->lsd_quant / std::sin(h->prec);
    // smallest integer n2 with sqrt(n2/4.0) > rho  (norm <= rho -> NOTDEF)
    { int v = 0; while (v < (1 << 20) && std::sqrt(v / 4.0) <= h->rho) ++v; h->n2_thresh = std::max(v, 1); }
    if (p->lsd_scale != 1.0) {
        const double sigma = (p->lsd_scale < 1) ? (p->lsd_sigma_scale / p->lsd_scale) : p->lsd_sigma_scale;
        const unsigned hk = (unsigned)(std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0))));
        h->blur_k = 1 + 2 * (int)hk;
        if (h->blur_k != 7 && h->blur_k != 5 && h->blur_k != 3) { set_last_error("olf_line_create: unsupported LSD pre-blur kernel size"); delete h; return nullptr; }
        int q[8] = {0}; gauss_kernel_q8(h->blur_k, sigma, q);
        for (int i = 0; i < 4; ++i) h->blur_q[i] = q[i];
    } else h->blur_k = 0;
    // trig tables keyed by (DA, BC), from host libm; 2 x 2 MB
    std::vector<float2_t> ts((size_t)kTabDim * kTabDim), ta((size_t)kTabDim * kTabDim);
    for (int DA = -255; DA <= 255; ++DA)
        for (int BC = -255; BC <= 255; ++BC) {
            const int gx = DA + BC, gy = DA - BC;
            // host restatement of cv::fastAtan2 (identical source to the device one)
            const double a = (double)olf::lsd::fast_atan2_deg((float)gx, (float)-gy) * kDegToRads;
            const size_t i = (size_t)(DA + 255) * kTabDim + (BC + 255);
            ts[i] = float2_t{(float)std::cos(a), (float)std::sin(a)};
            ta[i] = float2_t{(float)std::cos((double)(float)a), (float)std::sin((double)(float)a)};
        }
    // LBD weights (BinaryDescriptor ctor :217-259, integer divisions preserved)
    float gG[63], gL[21];
    {
        const int wb = 7, nb = 9;
        double u = (wb * 3 - 1) / 2, sigma = (wb * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < wb * 3; i++) { const double dis = i - u; gL[i] = (float)std::exp(dis * dis * inv); }
        u = (nb * wb - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < nb * wb; i++) { const double dis = i - u; gG[i] = (float)std::exp(dis * dis * inv); }
    }
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_grow0) == cudaSuccess && cudaEventCreate(&h->ev_grow1) == cudaSuccess;
    ok = ok && h->tab_seed.ensure(ts.size()) == OLF_OK && h->tab_acc.ensure(ta.size()) == OLF_OK;
    ok = ok && cudaMemcpy(h->tab_seed.p, ts.data(), ts.size() * sizeof(float2_t), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->tab_acc.p, ta.data(), ta.size() * sizeof(float2_t), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpyToSymbol(c_gaussG, gG, sizeof(gG)) == cudaSuccess && cudaMemcpyToSymbol(c_gaussL, gL, sizeof(gL)) == cudaSuccess;
    ok = ok && cudaMemcpyToSymbol(c_lbd_comb, LBD_COMB, sizeof(LBD_COMB)) == cudaSuccess;
    int coop = 0, sms = 0, per_sm = 0;
    ok = ok && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && coop;
    ok = ok && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    if (const char* e = getenv("OLF_LSD_FIRST_WAVE")) h->first_wave = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_WAVE_GROWTH")) h->wave_growth = std::max(2, atoi(e));
    h->trace = getenv("OLF_LSD_TRACE") != nullptr;                // per-round trace of the grow kernel (tools/lsd_trace.py)
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lsd_phase, GW_WARPS * 32, 0) == cudaSuccess && per_sm > 0;
    ok = ok && h->phase.ensure(1) == OLF_OK;
    if (const char* e = getenv("OLF_LSD_PHASE_BATCH")) h->phase_batch = std::max(4, atoi(e));
    if (!ok) { set_last_error(std::string("olf_line_create: ") + cudaGetErrorString(cudaGetLastError())); delete h; return nullptr; }
    {   // blocks per SM of the persistent grow kernel: the kernel is latency-bound (one warp walks the longest region), so a
        // small grid loses little and lets the left/right eyes and several frames in flight share the GPU
        // half a block (4 warps) per SM: measured optimum with 16 rigs in flight (larger grids only fight for block slots)
        h->grow_blocks = std::max(1, sms / 2);
        if (const char* e = getenv("OLF_LSD_BPS")) h->grow_blocks = sms * std::max(1, std::min(per_sm, atoi(e)));
        if (const char* e = getenv("OLF_LSD_BLOCKS")) h->grow_blocks = std::max(1, std::min(atoi(e), sms * per_sm));
    }
    return h;
}

void line_destroy(LineImpl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    if (h->ev_grow0) cudaEventDestroy(h->ev_grow0);
    if (h->ev_grow1) cudaEventDestroy(h->ev_grow1);
    h->img_stage.release(); h->img.release(); h->blurred.release(); h->scaled.release(); h->lbd_blur.release(); h->coef.release();
    h->ang.release(); h->dabc.release(); h->claim0.release(); h->claim1.release(); h->seed_prio.release(); h->seed_pix.release();
    h->cnt0.release(); h->cnt1.release(); h->n2max.release(); h->status.release(); h->head0.release(); h->head1.release();
    h->hist.release(); h->bin_start.release(); h->cursor.release(); h->pool0.release(); h->pool1.release(); h->ctrs.release();
    h->changed.release(); h->final_pool.release(); h->regang.release(); h->plan.release(); h->regs.release();
    h->tab_seed.release(); h->tab_acc.release(); h->cs.release(); h->work_ctr.release();
    h->blk_chunk0.release(); h->blk_chunk1.release(); h->blk_cnt0.release(); h->blk_cnt1.release(); h->phase.release(); h->px.release(); h->rect_host.release(); h->dir_host.release(); h->seg_host.release();
    h->status_host.release(); h->nreg_host.release(); h->grad.release(); h->lbd_lines.release(); h->rowsum.release(); h->desc_host.release();
    delete h;
}

static int line_ensure_size(LineImpl* h, int w, int hgt) {
    if (h->img_w == w && h->img_h == hgt) return OLF_OK;
    int rc;
    const double sc = h->P.lsd_scale;
    h->ipitch = align_up(w, 64);
    h->W = (sc != 1.0) ? (int)lrint(w * sc) : w;
    h->H = (sc != 1.0) ? (int)lrint(hgt * sc) : hgt;
    if (h->W < 3 || h->H < 3) { set_last_error("image too small for LSD"); return OLF_ERR_ARG; }
    h->wpitch = align_up(h->W, 64);
    h->S = h->W * h->H;
    const double logNT = 5 * (std::log10((double)h->W) + std::log10((double)h->H)) / 2 + std::log10(11.0);
    h->min_reg_size = (int)(unsigned)(-logNT / std::log10(h->P.lsd_ang_th / 180.0));
    const size_t ib = (size_t)h->ipitch * hgt + 256, wb = (size_t)h->wpitch * h->H + 256, S = h->S;
    if ((rc = h->img_stage.ensure((size_t)w * hgt)) || (rc = h->img.ensure(ib)) || (rc = h->blurred.ensure(ib)) || (rc = h->lbd_blur.ensure(ib)) ||
        (rc = h->scaled.ensure(wb)) || (rc = h->grad.ensure((size_t)w * hgt))) return rc;
    std::vector<ExCoef> cx, cy;
    build_exact_coefs(w, h->W, sc, cx); build_exact_coefs(hgt, h->H, sc, cy);
    h->coef_y_off = cx.size();
    cx.insert(cx.end(), cy.begin(), cy.end());
    if ((rc = h->coef.ensure(cx.size()))) return rc;
    OLF_CUDA(cudaMemcpy(h->coef.p, cx.data(), cx.size() * sizeof(ExCoef), cudaMemcpyHostToDevice));
    h->pool_chunks = (unsigned)std::max<size_t>(S / 2, 1u << 16);       // 16 px of list space per image pixel per round
    h->reg_cap = (unsigned)(S / std::max(h->min_reg_size, 1) + 16);
    if ((rc = h->ang.ensure(S)) || (rc = h->dabc.ensure(S)) || 
        (rc = h->seed_prio.ensure(S)) || (rc = h->seed_pix.ensure(S)) || (rc = h->cnt0.ensure(S)) || (rc = h->cnt1.ensure(S)) ||
        (rc = h->head0.ensure(S)) || (rc = h->head1.ensure(S)) || (rc = h->regang.ensure(S)) || (rc = h->final_pool.ensure(S)) ||
        (rc = h->n2max.ensure(1)) || (rc = h->status.ensure(4)) || (rc = h->hist.ensure(1024)) || (rc = h->bin_start.ensure(1024)) ||
        (rc = h->cursor.ensure(1024)) || (rc = h->ctrs.ensure(4)) || (rc = h->changed.ensure(h->max_rounds)) || (rc = h->plan.ensure(1)) ||
        (rc = h->pool0.ensure((size_t)h->pool_chunks * kChunk)) || (rc = h->pool1.ensure((size_t)h->pool_chunks * kChunk)) ||
        (rc = h->dbg.ensure((size_t)h->max_rounds * TRACE_REC)) || (rc = h->px.ensure(S)) || (rc = h->work_ctr.ensure(2 * h->max_rounds + 64)) ||
        (rc = h->blk_chunk0.ensure(S)) || (rc = h->blk_chunk1.ensure(S)) || (rc = h->blk_cnt0.ensure(S)) || (rc = h->blk_cnt1.ensure(S)) ||
        (rc = h->regs.ensure(h->reg_cap)) || (rc = h->rect_host.ensure(h->reg_cap)) || (rc = h->dir_host.ensure(h->reg_cap)) ||
        (rc = h->seg_host.ensure(h->reg_cap)) || (rc = h->status_host.ensure(4)) || (rc = h->nreg_host.ensure(1))) return rc;
    h->img_w = w; h->img_h = hgt;
    return OLF_OK;
}

static LevelTable single_level(int w, int hgt, int pitch) {
    LevelTable T; memset(&T, 0, sizeof(T));
    T.n = 1; T.w[0] = w; T.h[0] = hgt; T.pitch[0] = pitch; T.off[0] = 0;
    T.tiles_x[0] = (w + TILE_W - 1) / TILE_W; T.tile_start[0] = 0;
    T.tile_start[1] = T.tiles_x[0] * ((hgt + TILE_H - 1) / TILE_H);
    return T;
}

static int line_upload(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device) {
    int rc = line_ensure_size(h, w, hgt);
    if (rc) return rc;
    if (on_device) OLF_CUDA(cudaMemcpy2DAsync(h->img.p, h->ipitch, img, stride, w, hgt, cudaMemcpyDeviceToDevice, h->stream));
    else {
        for (int y = 0; y < hgt; ++y) memcpy(h->img_stage.p + (size_t)y * w, img + (size_t)y * stride, w);
        OLF_CUDA(cudaMemcpy2DAsync(h->img.p, h->ipitch, h->img_stage.p, w, w, hgt, cudaMemcpyHostToDevice, h->stream));
    }
    return OLF_OK;
}

// LSD on the uploaded image; returns segments sorted in seed order (host vector)
static int lsd_run(LineImpl* h, std::vector<float4>& segs) {
    cudaStream_t s = h->stream;
    const int w = h->img_w, hgt = h->img_h, W = h->W, H = h->H, S = h->S;
    const uint8_t* work = h->img.p; int wp = h->ipitch;
    if (h->blur_k) {
        const LevelTable T = single_level(w, hgt, h->ipitch);
        const int nt = T.tile_start[1];
        if (h->blur_k == 7) k_blur_q8<7><<<nt, 256, 0, s>>>(h->img.p, h->blurred.p, T, h->blur_q[0], h->blur_q[1], h->blur_q[2], h->blur_q[3]);
        else if (h->blur_k == 5) k_blur_q8<5><<<nt, 256, 0, s>>>(h->img.p, h->blurred.p, T, h->blur_q[0], h->blur_q[1], h->blur_q[2], 0);
        else k_blur_q8<3><<<nt, 256, 0, s>>>(h->img.p, h->blurred.p, T, h->blur_q[0], h->blur_q[1], 0, 0);
        dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
        k_resize_exact<<<g, b, 0, s>>>(h->blurred.p, w, hgt, h->ipitch, h->scaled.p, W, H, h->wpitch, h->coef.p, h->coef.p + h->coef_y_off);
        work = h->scaled.p; wp = h->wpitch;
    }
    OLF_CUDA(cudaMemsetAsync(h->n2max.p, 0, sizeof(int), s));
    OLF_CUDA(cudaMemsetAsync(h->hist.p, 0, 1024 * sizeof(unsigned), s));
    OLF_CUDA(cudaMemsetAsync(h->ctrs.p, 0, 4 * sizeof(unsigned), s));
    OLF_CUDA(cudaMemsetAsync(h->status.p, 0, 4 * sizeof(int), s));
    OLF_CUDA(cudaMemsetAsync(h->changed.p, 0, h->max_rounds * sizeof(unsigned), s));
    OLF_CUDA(cudaMemsetAsync(h->work_ctr.p, 0, (2 * h->max_rounds + 64) * sizeof(unsigned), s));
    {
        dim3 g((W + 31) / 32, (H + 7) / 8);
        k_lsd_grad<<<g, 256, 0, s>>>(work, W, H, wp, h->n2_thresh, h->ang.p, h->dabc.p, h->tab_acc.p, h->px.p, h->n2max.p);
    }
    const int nb = h->P.lsd_n_bins;
    k_lsd_hist<<<296, 256, nb * sizeof(unsigned), s>>>(h->ang.p, h->dabc.p, S, h->n2max.p, nb, h->hist.p);
    k_lsd_plan<<<1, 1024, 0, s>>>(h->hist.p, nb, h->first_wave, h->wave_growth, h->bin_start.p, h->cursor.p, h->plan.p);
    k_lsd_scatter<<<296, 256, 0, s>>>(h->ang.p, h->dabc.p, S, h->n2max.p, nb, h->bin_start.p, h->cursor.p, h->seed_pix.p, h->seed_prio.p, h->px.p);
    GrowState G;
    G.A.W = W; G.A.H = H; G.A.ang = h->ang.p; G.A.dabc = h->dabc.p; G.A.tab_seed = h->tab_seed.p; G.A.tab_acc = h->tab_acc.p;
    G.A.claim[0] = nullptr; G.A.claim[1] = nullptr; G.A.pool[0] = h->pool0.p; G.A.pool[1] = h->pool1.p;
    G.A.pool_ctr[0] = h->ctrs.p; G.A.pool_ctr[1] = h->ctrs.p + 1; G.A.pool_chunks = h->pool_chunks; G.A.prec = h->prec;
    G.seed_pix = h->seed_pix.p; G.seed_prio = h->seed_prio.p;
    G.head[0] = h->head0.p; G.head[1] = h->head1.p; G.cnt[0] = h->cnt0.p; G.cnt[1] = h->cnt1.p; G.regang = h->regang.p;
    G.plan = h->plan.p; G.changed = h->changed.p; G.max_rounds = h->max_rounds; G.min_reg_size = h->min_reg_size;
    G.final_pool = h->final_pool.p; G.final_ctr = h->ctrs.p + 2; G.regs = h->regs.p; G.nreg = h->ctrs.p + 3; G.reg_cap = h->reg_cap;
    G.status = h->status.p;
    GrowStateW GW; GW.G = G; GW.px = h->px.p; GW.work_ctr = h->work_ctr.p;
    GW.G.A.pool[1] = GW.G.A.pool[0]; GW.G.A.pool_ctr[1] = GW.G.A.pool_ctr[0];      // warp kernel: one bump pool per wave
    GW.blk_chunk[0] = h->blk_chunk0.p; GW.blk_chunk[1] = h->blk_chunk1.p; GW.blk_cnt[0] = h->blk_cnt0.p; GW.blk_cnt[1] = h->blk_cnt1.p;
    {
        const double margin = 0.1 * M_PI / 180.0;
        GW.fast_align = (h->prec + margin < 80.0 * M_PI / 180.0) && !getenv("OLF_LSD_EXACT_ALIGN");
        GW.c_hi2 = (float)(std::cos(h->prec - margin) * std::cos(h->prec - margin));
        GW.c_lo2 = (float)(std::cos(h->prec + margin) * std::cos(h->prec + margin));
    }
    GW.defer = getenv("OLF_LSD_NO_DEFER") ? 0 : 1;
    GW.dbg = h->trace ? h->dbg.p : nullptr;
    if (h->trace) OLF_CUDA(cudaMemsetAsync(h->dbg.p, 0, (size_t)h->max_rounds * TRACE_REC * sizeof(int), s));
    OLF_CUDA(cudaEventRecord(h->ev_grow0, s));
    {
        PhaseState init; memset(&init, 0, sizeof(init)); init.round = 1; init.wave_first_round = 1;
        // seeds of later waves must start with "no previous list"
        OLF_CUDA(cudaMemsetAsync(h->cnt0.p, 0, (size_t)S * sizeof(int), s));
        OLF_CUDA(cudaMemsetAsync(h->cnt1.p, 0, (size_t)S * sizeof(int), s));
        OLF_CUDA(cudaMemcpyAsync(h->phase.p, &init, sizeof(init), cudaMemcpyHostToDevice, s));
        for (int k = 0; k < h->phase_batch; ++k) k_lsd_phase<<<h->grow_blocks, GW_WARPS * 32, 0, s>>>(GW, h->phase.p);
        count_launches(h->phase_batch);
    }
    OLF_CUDA(cudaEventRecord(h->ev_grow1, s));
    count_launches((h->blur_k ? 2 : 0) + 6);
    k_lsd_rect_a<<<296, 256, 0, s>>>(h->regs.p, h->ctrs.p + 3, h->reg_cap, h->final_pool.p, h->dabc.p, W, h->prec, h->rect_host.d);
    OLF_CUDA(cudaMemcpyAsync(h->nreg_host.p, h->ctrs.p + 3, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaMemcpyAsync(h->status_host.p, h->status.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    {
        // the fixed batch of phase launches normally covers all rounds; otherwise keep going (rare)
        for (int guard = 0; guard < 400; ++guard) {
            if (h->status_host.p[3] || h->status_host.p[0]) break;
            for (int k = 0; k < h->phase_batch; ++k) k_lsd_phase<<<h->grow_blocks, GW_WARPS * 32, 0, s>>>(GW, h->phase.p);
            count_launches(h->phase_batch + 1);
            OLF_CUDA(cudaEventRecord(h->ev_grow1, s));
            k_lsd_rect_a<<<296, 256, 0, s>>>(h->regs.p, h->ctrs.p + 3, h->reg_cap, h->final_pool.p, h->dabc.p, W, h->prec, h->rect_host.d);
            OLF_CUDA(cudaMemcpyAsync(h->nreg_host.p, h->ctrs.p + 3, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
            OLF_CUDA(cudaMemcpyAsync(h->status_host.p, h->status.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            OLF_CUDA(stream_sync(s));
        }
    }
    h->last_stats[0] = h->status_host.p[1]; h->last_stats[1] = h->status_host.p[2];
    { float ms = 0; if (cudaEventElapsedTime(&ms, h->ev_grow0, h->ev_grow1) == cudaSuccess) h->last_stats[3] = (int)(ms * 1000.f); }
    if (h->status_host.p[0] != 0) { set_last_error("LSD region growing: internal pool overflow"); return OLF_ERR_CAPACITY; }
    if (h->status_host.p[1] + 2 >= (int)h->max_rounds) { set_last_error("LSD region growing did not converge"); return OLF_ERR_INTERNAL; }
    const int n = (int)*h->nreg_host.p;
    h->last_stats[2] = n;
    segs.clear();
    if (n == 0) return OLF_OK;
    // libm cos/sin of the O(#regions) rectangle angles, then the projection pass on the device
    for (int i = 0; i < n; ++i) { const double t = h->rect_host.p[i].theta; h->dir_host.p[i] = make_double2(std::cos(t), std::sin(t)); }
    k_lsd_rect_b<<<(n + 7) / 8, 256, 0, s>>>(h->regs.p, n, h->final_pool.p, W, h->rect_host.d, h->dir_host.d, h->P.lsd_scale, h->seg_host.d);
    count_launches(1);
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    // seed order = ascending priority key
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    const RectRec* rr = h->rect_host.p;
    std::sort(order.begin(), order.end(), [rr](int a, int b) { return rr[a].prio < rr[b].prio; });
    segs.resize(n);
    for (int i = 0; i < n; ++i) segs[i] = h->seg_host.p[order[i]];
    return OLF_OK;
}

// checkLineExtremes + KeyLine construction (LSDDetector_custom.cpp:76-102, 264-308), numOctaves = 1
static void make_keylines(const std::vector<float4>& segs, int w, int hgt, double min_length, std::vector<olf_keyline>& out) {
    out.clear();
    int class_counter = -1;
    for (const float4& sg : segs) {
        float e[4] = {sg.x, sg.y, sg.z, sg.w};
        if (e[0] < 0) e[0] = 0;
        if (e[0] >= w) e[0] = (float)w - 1.0f;
        if (e[2] < 0) e[2] = 0;
        if (e[2] >= w) e[2] = (float)w - 1.0f;
        if (e[1] < 0) e[1] = 0;
        if (e[1] >= hgt) e[1] = (float)hgt - 1.0f;
        if (e[3] < 0) e[3] = 0;
        if (e[3] >= hgt) e[3] = (float)hgt - 1.0f;
        const double length = (float)std::sqrt(std::pow((double)(e[0] - e[2]), 2) + std::pow((double)(e[1] - e[3]), 2));
        if (!(length > min_length)) continue;
        olf_keyline kl;
        kl.startPointX = e[0]; kl.startPointY = e[1]; kl.endPointX = e[2]; kl.endPointY = e[3];
        kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1]; kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
        kl.lineLength = (float)length;
        const int x1 = (int)lrintf(e[0]), y1 = (int)lrintf(e[1]), x2 = (int)lrintf(e[2]), y2 = (int)lrintf(e[3]);
        kl.numOfPixels = std::max(std::abs(x2 - x1), std::abs(y2 - y1)) + 1;        // cv::LineIterator(...).count, 8-connected
        kl.angle = (float)std::atan2((double)(kl.endPointY - kl.startPointY), (double)(kl.endPointX - kl.startPointX));
        kl.class_id = ++class_counter;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, hgt);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2; kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        out.push_back(kl);
    }
}

// LBD of `n` keylines on the uploaded image: blur 5x5 sigma 1 + Sobel, row sums, fold + binarise
static int lbd_run(LineImpl* h, const olf_keyline* kls, int n, uint8_t* desc) {
    if (n == 0) return OLF_OK;
    cudaStream_t s = h->stream;
    const int w = h->img_w, hgt = h->img_h;
    int rc;
    if (n > h->lbd_cap) {
        const int cap = std::max(2 * n, 1024);
        if ((rc = h->lbd_lines.ensure(cap)) || (rc = h->rowsum.ensure((size_t)cap * 63)) || (rc = h->desc_host.ensure((size_t)cap * 32))) return rc;
        h->lbd_cap = cap;
    }
    for (int i = 0; i < n; ++i) {
        const olf_keyline& k = kls[i];
        LbdLine L;
        L.sx = k.sPointInOctaveX; L.sy = k.sPointInOctaveY; L.ex = k.ePointInOctaveX; L.ey = k.ePointInOctaveY;
        L.dL0 = (float)std::cos((double)k.angle); L.dL1 = (float)std::sin((double)k.angle);       // :1130-1131
        L.num_px = k.numOfPixels;
        h->lbd_lines.p[i] = L;
    }
    const LevelTable T = single_level(w, hgt, h->ipitch);
    k_blur_q8<5><<<T.tile_start[1], 256, 0, s>>>(h->img.p, h->lbd_blur.p, T, 14, 62, 104, 0);       // 5x5 sigma 1 (:358)
    dim3 g((w + 31) / 32, (hgt + 7) / 8);
    k_sobel3<<<g, 256, 0, s>>>(h->lbd_blur.p, w, hgt, h->ipitch, h->grad.p);
    k_lbd_rows<<<(n * 63 + 255) / 256, 256, 0, s>>>(h->lbd_lines.d, n, h->grad.p, w, hgt, h->rowsum.p);
    k_lbd_fold<<<(n + 63) / 64, 64, 0, s>>>(h->rowsum.p, n, h->desc_host.d);
    count_launches(4);
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(desc, h->desc_host.p, (size_t)n * 32);
    return OLF_OK;
}

int line_lsd_detect(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, float* segs, int cap, int* n) {
    if (!h || !img || !n || w <= 0 || hgt <= 0 || stride < w) { set_last_error("olf_lsd_detect: bad arguments"); return OLF_ERR_ARG; }
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = line_upload(h, img, w, hgt, stride, on_device);
    if (rc) return rc;
    std::vector<float4> sg;
    if ((rc = lsd_run(h, sg))) return rc;
    *n = (int)sg.size();
    if (*n > cap) { set_last_error("olf_lsd_detect: segment capacity too small"); return OLF_ERR_CAPACITY; }
    if (*n) memcpy(segs, sg.data(), sg.size() * sizeof(float4));
    return OLF_OK;
}

int line_lbd_compute(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, uint8_t* desc) {
    if (!h || !img || w <= 0 || hgt <= 0 || stride < w || n < 0) { set_last_error("olf_lbd_compute: bad arguments"); return OLF_ERR_ARG; }
    if (n == 0) return OLF_OK;                           // "keypoint list is empty": descriptors untouched (:556-560)
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = line_upload(h, img, w, hgt, stride, false);
    if (rc) return rc;
    return lbd_run(h, kls, n, desc);
}

// Lineextractor::operator() (src/LineExtractor.cc:31-67)
int line_extract(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keyline* kls, uint8_t* desc, int cap, int* n) {
    if (!h || !img || !n || w <= 0 || hgt <= 0 || stride < w) { set_last_error("olf_line_extract: bad arguments"); return OLF_ERR_ARG; }
    *n = 0;
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = line_upload(h, img, w, hgt, stride, on_device);
    if (rc) return rc;
    std::vector<float4> sg;
    if ((rc = lsd_run(h, sg))) return rc;
    std::vector<olf_keyline> k;
    make_keylines(sg, w, hgt, h->P.min_line_length * std::min(w, hgt), k);
    const int nf = h->P.lsd_nfeatures;
    if ((int)k.size() > nf && nf != 0) {
        // canonical: stable sort by response
        std::stable_sort(k.begin(), k.end(), [](const olf_keyline& a, const olf_keyline& b) { return a.response > b.response; });
        k.resize(nf);
        for (int i = 0; i < nf; i++) k[i].class_id = i;
    }
    if ((int)k.size() > cap) { set_last_error("olf_line_extract: keyline capacity too small"); return OLF_ERR_CAPACITY; }
    *n = (int)k.size();
    if (k.empty()) return OLF_OK;
    memcpy(kls, k.data(), k.size() * sizeof(olf_keyline));
    return lbd_run(h, k.data(), (int)k.size(), desc);
}

// [0] rounds, [1] waves, [2] accepted regions, [3] k_lsd_grow device time in microseconds (CUDA events on its stream)
void line_last_stats(const LineImpl* h, int* out8) { for (int i = 0; i < 8; ++i) out8[i] = h->last_stats[i]; }
cudaStream_t line_stream(const LineImpl* h) { return h ? h->stream : nullptr; }
int line_trace(LineImpl* h, int* out, int max_rounds) {
    if (!h || !h->trace || !h->dbg.p) return OLF_ERR_ARG;
    const int n = std::min<int>(max_rounds, (int)h->max_rounds);
    OLF_CUDA(cudaMemcpy(out, h->dbg.p, (size_t)n * TRACE_REC * sizeof(int), cudaMemcpyDeviceToHost));
    return OLF_OK;
}

}  // namespace olf
