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
#include "lsd_sticky.h"
#include "line.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <chrono>
#include <cstddef>


namespace olf {
using namespace lsd;

// ---- cv::resize INTER_LINEAR_EXACT 8UC1 -------------------------------------------------------
struct ExCoef { int s; int c0, c1; };
// Everything before region growing runs ONE launch per stage for all images of a call (blockIdx.z / .y = image): the images of a rig have the
// same size and parameters, and every one of these stages is launch-bound on a single 0.9-MB image.
#define LSD_PRE_MAX 16
struct LsdPlan;
struct PhaseState;
struct PreBatch {
    int n;
    const uint8_t* work_src[LSD_PRE_MAX]; uint8_t* scaled[LSD_PRE_MAX];          // resize: blurred (or raw) image -> scaled image
    const uint8_t* grad_src[LSD_PRE_MAX];                                        // gradient input (scaled image, or the raw one at scale 1)
    float* ang[LSD_PRE_MAX]; short2_t* dabc[LSD_PRE_MAX]; PxRec* px[LSD_PRE_MAX];
    int* n2max[LSD_PRE_MAX]; unsigned* hist[LSD_PRE_MAX]; unsigned* bin_start[LSD_PRE_MAX]; unsigned* cursor[LSD_PRE_MAX];
    int* seed_pix[LSD_PRE_MAX]; u64* seed_prio[LSD_PRE_MAX]; LsdPlan* plan[LSD_PRE_MAX];
    unsigned* ctrs[LSD_PRE_MAX]; int* status[LSD_PRE_MAX]; PhaseState* phase[LSD_PRE_MAX];
};
__global__ void k_resize_exact(const __grid_constant__ PreBatch PB, int sw, int sh, int spitch, int dw, int dh, int dpitch,
                               const ExCoef* __restrict__ cx, const ExCoef* __restrict__ cy) {
    const uint8_t* __restrict__ src = PB.work_src[blockIdx.z];
    uint8_t* __restrict__ dst = PB.scaled[blockIdx.z];
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

// ---- ll_angle: gradient, level-line angle, max gradient ---------------------------------
// n2_thresh = smallest gx^2+gy^2 whose norm sqrt(n2/4.0) exceeds rho (computed exactly on the host).
__global__ void __launch_bounds__(256) k_lsd_grad(const __grid_constant__ PreBatch PB, int W, int H, int pitch, int n2_thresh,
                                                  const float2_t* __restrict__ tab_acc) {
    const uint8_t* __restrict__ img = PB.grad_src[blockIdx.z];
    float* __restrict__ ang = PB.ang[blockIdx.z]; short2_t* __restrict__ dabc = PB.dabc[blockIdx.z];
    PxRec* __restrict__ px = PB.px[blockIdx.z]; int* __restrict__ n2max = PB.n2max[blockIdx.z];
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
        // NOTDEF pixels are never accepted by isAligned(): they are born with a final claim (stamp 0), so the grow pass needs no angle test
        PxRec r; r.claim[0] = a >= 0.f ? kClaimNone : 0ull; r.claim[1] = r.claim[0]; r.ang = a; r.cx = c.x; r.cy = c.y; r.binrev = 0;
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
__global__ void __launch_bounds__(256) k_lsd_hist(const __grid_constant__ PreBatch PB, int S, int n_bins) {
    const float* __restrict__ ang = PB.ang[blockIdx.y]; const short2_t* __restrict__ dabc = PB.dabc[blockIdx.y];
    const int* __restrict__ n2max = PB.n2max[blockIdx.y]; unsigned* __restrict__ hist = PB.hist[blockIdx.y];
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
__global__ void __launch_bounds__(1024) k_lsd_plan(const __grid_constant__ PreBatch PB, int n_bins, int first_wave, int wave_growth) {
    const unsigned* __restrict__ hist = PB.hist[blockIdx.x]; unsigned* __restrict__ bin_start = PB.bin_start[blockIdx.x];
    unsigned* __restrict__ cursor = PB.cursor[blockIdx.x]; LsdPlan* __restrict__ plan = PB.plan[blockIdx.x];
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

__global__ void __launch_bounds__(256) k_lsd_scatter(const __grid_constant__ PreBatch PB, int S, int n_bins) {
    const float* __restrict__ ang = PB.ang[blockIdx.y]; const short2_t* __restrict__ dabc = PB.dabc[blockIdx.y];
    const int* __restrict__ n2max = PB.n2max[blockIdx.y]; const unsigned* __restrict__ bin_start = PB.bin_start[blockIdx.y];
    unsigned* __restrict__ cursor = PB.cursor[blockIdx.y]; int* __restrict__ seed_pix = PB.seed_pix[blockIdx.y];
    u64* __restrict__ seed_prio = PB.seed_prio[blockIdx.y]; PxRec* __restrict__ px = PB.px[blockIdx.y];
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

// ---- region growing: three passes per round, one thread per seed (semantics: lsd_sticky.h) ----------------------------
// The wave / round state machine lives on the device: the host enqueues a fixed batch of (scan, verify, grow) triples,
// the last block of every grow launch advances {wave, round, mode}, launches after `done` return at once.  Ordinary
// (non-cooperative) launches: no co-residency requirement, so the passes of several rigs in flight interleave freely with
// every other kernel on the device.
#define TRACE_REC 8
// EXPERIMENT (-DOLF_PDL=1, tools/try_pdl.sh; off in the shipped library): programmatic dependent launch between the three pass kernels of the
// plain-launch chain.  A pass waits for the grid before it (griddepcontrol.wait) and then lets the grid after it be launched; that grid's CTAs
// sit at their own wait until this one has completed, so the start latency of a dependent launch is paid while the predecessor still runs.
#ifndef OLF_PDL
#define OLF_PDL 0
#endif
__device__ __forceinline__ void pdl_enter() {
#if OLF_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
struct PhaseState {
    int wave; unsigned round; int mode; int done;           // mode 0: round passes, 2: converged, waiting for the batch, 3: full verification, 1: finalise
    int launches; unsigned wave_first_round;
    unsigned wl0_cnt, wl1_cnt, wl2_cnt, wl2_pop, changed, ticket;
    unsigned recheck, force, pad0, pad1;                     // full verification failed / the next round verifies every region
    // grow pass in instalments: regions that used up a launch's step budget wait in cont[cont_par] for the next launch
    unsigned cont_cnt[2], cont_par, growing;
};
// a paused region: everything k_lsd_grow keeps in registers for a seed (64 bytes)
struct alignas(16) GrowCont {
    int i; unsigned w_head, w_chunk, r_chunk, b_head, b_chunk;
    unsigned offs;                                           // w_off | r_off << 8 | b_off << 16 | dirty << 24
    int count, done, bcnt;
    unsigned short bx0, by0, bx1, by1;
    float sumdx, sumdy; double reg_angle;
};
struct GrowDev {
    Ctx3 C;
    const LsdPlan* plan;
    int* wl0; int* wl1; int* wl2;   // candidates of the wave (not finalised at its start) / seeds to verify (~i: died) / seeds that must grow
    FinalOut F;
    int* status;                 // [0] error flag, [1] rounds used, [2] waves, [3] done
    unsigned max_rounds;
    int defer;                   // first round of a wave: seeds with a live higher-priority aligned neighbour wait
    int dirty_words;             // words of one dirty bitmap
    int test_recheck;            // (test hook, OLF_LSD_TEST_RECHECK) pretend the full verification of every wave fails once
    int* dbg;                    // optional per-round trace (OLF_LSD_TRACE)
    GrowCont* cont[2];           // paused regions, by launch parity (capacity: threads of one grow launch per image)
    int budget;                  // queue entries a thread may expand per grow launch
    // Event-driven bookkeeping (what the passes read every round is 9 bytes per candidate instead of ~100):
    unsigned char* sstate;       // per seed: 1 = alive at the last evaluation.  A seed's state can only change where an event dirtied
                                 // the tile of its pixel (takeover / release), or in the verify pass (which writes a death back)
    unsigned* tbox;              // per seed: the tile-coordinate bounding box of its region packed 4 x 8 bits (tx0 | ty0 << 8 | tx1 << 16 | ty1 << 24),
                                 // kNull = no region, kTboxWide = look at the seed record (an image wider than 255 tiles)
    int event_scan;              // 0: evaluate every candidate every round (OLF_LSD_FULL_SCAN, the previous behaviour)
};
constexpr unsigned kTboxWide = 0xFFFFFFFEu;
__device__ __forceinline__ unsigned tbox_pack(int x0, int y0, int x1, int y1) {
    const int tx0 = x0 >> kTileShift, ty0 = y0 >> kTileShift, tx1 = x1 >> kTileShift, ty1 = y1 >> kTileShift;
    if ((tx1 | ty1) > 254) return kTboxWide;
    return (unsigned)tx0 | ((unsigned)ty0 << 8) | ((unsigned)tx1 << 16) | ((unsigned)ty1 << 24);
}
// the same test as bbox_dirty() (lsd_sticky.h) on the packed tile box
__device__ __forceinline__ bool tbox_dirty(const Ctx3& C, unsigned round, unsigned tb) {
    const unsigned* d = C.dirty[(round - 1) & 1];
    const int tx0 = tb & 255, ty0 = (tb >> 8) & 255, tx1 = (tb >> 16) & 255, ty1 = tb >> 24;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int w = tx0 >> 5; w <= tx1 >> 5; ++w) {
            const int lo = tx0 > w * 32 ? tx0 - w * 32 : 0, hi = tx1 < w * 32 + 31 ? tx1 - w * 32 : 31;
            const unsigned mask = (hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            if (d[ty * C.tile_wpr + w] & mask) return true;
        }
    return false;
}
__device__ __forceinline__ bool pixel_tile_dirty(const Ctx3& C, unsigned round, int pix) {
    const int y = pix / C.W, x = pix - y * C.W;
    const int tx = x >> kTileShift, ty = y >> kTileShift;
    return (C.dirty[(round - 1) & 1][ty * C.tile_wpr + (tx >> 5)] >> (tx & 31)) & 1u;
}
// One launch serves a BATCH of images (the two eyes of a stereo frame, several frames): blockIdx.y selects the image, every
// image has its own state machine.  The passes are latency-bound with small grids, so a batch costs one chain of launches
// on one stream instead of one chain per image -- that is what keeps many frames in flight within the 32 hardware queues.
#define LSD_MAX_BATCH 16
#define LSD_MAX_WAVES 64
// conv[w]: images of the batch that have converged in wave w; conv[LSD_MAX_WAVES]: images that are done.
// cond: handle of the graph's WHILE node when the chain runs as a CUDA graph (0: plain launches)
struct GrowBatch { GrowDev d[LSD_MAX_BATCH]; PhaseState* st[LSD_MAX_BATCH]; unsigned* conv; int n; cudaGraphConditionalHandle cond; };
__device__ __forceinline__ const GrowDev& batch_image(const GrowBatch& B, GrowDev* sh, PhaseState*& st) {
    // block-uniform copy of this image's descriptor into shared memory (the by-value batch lives in parameter space);
    // the claim stamp of the image's current wave is filled in here
    const int* src = reinterpret_cast<const int*>(&B.d[blockIdx.y]);
    int* dst = reinterpret_cast<int*>(sh);
    for (int k = threadIdx.x; k < (int)(sizeof(GrowDev) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
    st = B.st[blockIdx.y];
    __syncthreads();
    if (threadIdx.x == 0) sh->C.stamp = (u64)(0xFFFFFFu - (unsigned)(st->wave + 1)) << 40;
    __syncthreads();
    return *sh;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_f64(double v, int src) { return __longlong_as_double((long long)shfl_u64((u64)__double_as_longlong(v), src)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
// warp-aggregated append of the lanes with `take` to a work list
__device__ __forceinline__ void wl_append(bool take, int value, int* __restrict__ list, unsigned* __restrict__ counter) {
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    unsigned base = 0;
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) list[base + __popc(m & lanemask_lt())] = value;
}

// an image has finished all its waves: the last one ends the WHILE loop of the graph
__device__ __forceinline__ void image_done(const GrowBatch& B) {
    if (atomicAdd(&B.conv[LSD_MAX_WAVES], 1u) + 1u == (unsigned)B.n && B.cond) cudaGraphSetConditional(B.cond, 0);
}
// start of a call: counters, histogram, status and the state machine of every image of the batch (one launch instead of five stream operations per image)
__global__ void __launch_bounds__(256) k_lsd_reset(const __grid_constant__ PreBatch PB) {
    const int k = blockIdx.x;
    for (int i = threadIdx.x; i < 1024; i += 256) PB.hist[k][i] = 0u;
    if (threadIdx.x < 4) { PB.ctrs[k][threadIdx.x] = 0u; PB.status[k][threadIdx.x] = 0; }
    if (threadIdx.x == 0) {
        *PB.n2max[k] = 0;
        PhaseState z; memset(&z, 0, sizeof(z)); z.round = 1; z.wave_first_round = 1;
        *PB.phase[k] = z;
    }
}

// pass 1: every candidate of the wave -- dead or alive (+ first-round deferral); alive seeds and seeds that died owning a
// region go to work list 1; the bitmap that collects THIS round's events is cleared
__global__ void __launch_bounds__(256) k_lsd_scan(const __grid_constant__ GrowBatch B) {
    pdl_enter();
    __shared__ GrowDev s_dev;
    PhaseState* st;
    const GrowDev& D = batch_image(B, &s_dev, st);
    if (st->done || st->mode != 0 || st->growing) return;
    const int wv = st->wave;
    if (wv >= D.plan->n_waves) return;
    const unsigned round = st->round;
    const bool first_round = round == st->wave_first_round;
    const int lo = D.plan->wave_start[wv], hi = D.plan->wave_start[wv + 1];
    if (D.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        D.dbg[round * TRACE_REC + 0] = wv; D.dbg[round * TRACE_REC + 1] = hi - lo; D.dbg[round * TRACE_REC + 2] = (int)(gtime() & 0x7fffffff);
    }
    if (blockIdx.x == 0) for (int k = threadIdx.x; k < D.dirty_words; k += blockDim.x) D.C.dirty[round & 1][k] = 0u;
    bool chg = false;
    // first round of a wave: all its seeds, those not yet swallowed by a finalised region become the wave's candidates;
    // later rounds: the candidates only.  From the second round on a candidate is re-evaluated only where the previous round's
    // events dirtied the tile of its pixel (or after a failed full verification: everything); otherwise its recorded state stands.
    const int first = first_round ? lo : 0, last = first_round ? hi : (int)st->wl0_cnt;
    const bool everything = first_round || st->force != 0 || !D.event_scan;
    for (int base = first + blockIdx.x * blockDim.x; base < last; base += gridDim.x * blockDim.x) {
        const int j = base + threadIdx.x;
        bool cand = false, list = false;
        int i = 0, entry = 0;
        if (j < last) {
            i = first_round ? j : D.wl0[j];
            const int seed = D.C.seed_pix[i];
            if (first_round) {
                cand = !s3_final(D.C, seed);
                if (cand) { SeedRec3 z; z.head = kNull; z.cnt = 0; z.bhead = kNull; z.bcnt = 0; z.x0 = z.y0 = z.x1 = z.y1 = 0; z.pad0 = z.pad1 = 0; D.C.srec[i] = z; D.tbox[i] = kNull; }
            }
            if (cand || !first_round) {
                bool alive;
                if (everything || pixel_tile_dirty(D.C, round, seed)) {
                    alive = s3_alive(D.C, seed, D.C.seed_prio[i]);
                    D.sstate[i] = alive ? 1 : 0;
                    if (!alive && !first_round && D.C.srec[i].cnt > 0) { list = true; entry = ~i; }                 // died owning a region
                } else alive = D.sstate[i] != 0;
                if (alive && first_round && D.defer && s3_deferred(D.C, seed, D.C.seed_prio[i])) chg = true;       // sits this round out
                else if (alive) { list = true; entry = i; }
            }
        }
        if (first_round) wl_append(cand, i, D.wl0, &st->wl0_cnt);
        wl_append(list, entry, D.wl1, &st->wl1_cnt);
    }
    if (__syncthreads_or(chg) && threadIdx.x == 0) st->changed = 1;
}

// Warp-cooperative check / release of ONE long list (31 entries of a chunk per step).
// what 0: every pixel still mine?  1: every refused candidate still held (or final)?
__device__ bool walk3_warp(const Ctx3& C, unsigned head, int cnt, int what, u64 mine, int lane) {
    bool ok = true;
    unsigned chunk = head;
    for (int k = 0; k < cnt && ok; k += kChunk - 1) {
        const unsigned v = C.pool[(size_t)chunk * kChunk + lane];
        bool bad = false;
        if (lane < kChunk - 1 && k + lane < cnt) {
            const u64 c = ld_claim0(&C.px[v]);
            bad = what == 0 ? c != mine : !(c < mine);
        }
        ok = !__any_sync(0xffffffffu, bad);
        chunk = __shfl_sync(0xffffffffu, v, kChunk - 1);
    }
    return ok;
}
__device__ void release3_warp(const Ctx3& C, unsigned round, unsigned head, int cnt, u64 mine, unsigned keep, int lane) {
    unsigned chunk = head;
    for (int k = 0; k < cnt; k += kChunk - 1) {
        const unsigned v = C.pool[(size_t)chunk * kChunk + lane];
        if (lane < kChunk - 1 && k + lane < cnt && v != keep)
            if (atomicCAS(&C.px[v].claim[0], mine, kClaimNone) == mine) mark_dirty(C, round, v);
        chunk = __shfl_sync(0xffffffffu, v, kChunk - 1);
    }
}

// pass 2: every listed seed: dead -> give its region back; alive without a region -> claim the seed pixel, go to work list 2;
// alive with a region -> nothing at all unless a tile its bounding box touches was dirtied in the previous round, then the
// lists are checked (long ones by a whole warp) and a region that no longer holds is released and re-grown.
// In finalise mode: every alive seed of the converged wave is stamped for good.
#define VERIFY_LONG 48
__global__ void __launch_bounds__(128, 6) k_lsd_verify(const __grid_constant__ GrowBatch B) {
    pdl_enter();
    __shared__ GrowDev s_dev;
    PhaseState* st;
    const GrowDev& D = batch_image(B, &s_dev, st);
    if (st->done || st->growing) return;
    const int wv = st->wave;
    if (wv >= D.plan->n_waves) return;
    const unsigned round = st->round;
    const int n = (int)st->wl1_cnt;
    const int nth = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const Ctx3& C = D.C;
    if (st->mode == 0) {
        bool chg = false;
        int carried = 0;
        const bool force = st->force != 0;
        // entry k of the list -> warp k % nwarps, lane k / nwarps: the long lists (walked by a whole warp, one after the other)
        // are spread over as many warps as possible
        const int nwarps = nth >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        for (int it = 0; it * 32 * nwarps + gw < n; ++it) {
            const int k = (it * 32 + lane) * nwarps + gw;
            bool grow = false, is_long = false; int i = 0;
            if (k < n) {
                const int e = D.wl1[k];
                const bool alive = e >= 0;
                i = alive ? e : ~e;
                // the common case first, on 4 bytes: a live region whose tile box no event touched is carried as it is
                const unsigned tb = D.tbox[i];
                if (alive && !force && tb < kTboxWide && !tbox_dirty(C, round, tb)) ++carried;
                else {
                    const SeedRec3 r = C.srec[i];
                    if (alive && r.cnt > 0 && r.cnt + r.bcnt > VERIFY_LONG) is_long = force || bbox_dirty(C, round, r);      // else: carried, O(1)
                    else if (!alive && r.cnt > VERIFY_LONG) is_long = true;
                    else {
                        const Verify3 v = s3_verify(C, round, i, alive, &chg, force);
                        grow = v == kV3Grow; carried += v == kV3Carried;
                        if (v != kV3Carried) D.tbox[i] = kNull;                      // the region is gone (released, or never was)
                        if (alive && v == kV3Dead) D.sstate[i] = 0;                   // lost the seed pixel meanwhile: the scan's record follows
                    }
                    if (alive && r.cnt > 0 && r.cnt + r.bcnt > VERIFY_LONG && !is_long) ++carried;
                }
            }
            unsigned lm = __ballot_sync(0xffffffffu, is_long);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const int si = __shfl_sync(0xffffffffu, i, src);
                const int se = __shfl_sync(0xffffffffu, k < n ? D.wl1[k] : 0, src);
                const bool alive = se >= 0;
                const int seed = C.seed_pix[si];
                const u64 mine = key_of(C, C.seed_prio[si]);
                const SeedRec3 r = C.srec[si];
                bool ok = false;
                if (alive) ok = walk3_warp(C, r.head, r.cnt, 0, mine, lane) && walk3_warp(C, r.bhead, r.bcnt, 1, mine, lane);
                if (!ok) release3_warp(C, round, r.head, r.cnt, mine, alive ? (unsigned)seed : kNull, lane);
                if (lane == src) {
                    if (ok) ++carried;
                    else {
                        SeedRec3 z = r; z.cnt = 0; z.bcnt = 0; z.head = kNull; z.bhead = kNull; C.srec[si] = z; D.tbox[si] = kNull;
                        chg = true;
                        if (alive) {                                      // still owns the seed pixel? (see s3_verify)
                            bool dead = false;
                            if (ld_claim0(&C.px[seed]) != mine) {
                                const u64 old = atomicMin(&C.px[seed].claim[0], mine);
                                dead = old < mine;
                                if (!dead && old != mine && claim_valid(C, old)) mark_dirty(C, round, (unsigned)seed);
                            }
                            grow = !dead;
                            if (dead) D.sstate[si] = 0;
                        }
                    }
                }
            }
            wl_append(grow, i, D.wl2, &st->wl2_cnt);
        }
        if (D.dbg && carried) atomicAdd(&D.dbg[round * TRACE_REC + 4], carried);
        if (__syncthreads_or(chg) && threadIdx.x == 0) st->changed = 1;
    } else if (st->mode == 3) {
        // full verification of the converged wave (the grow pass claims without looking at return values): every live region,
        // whatever the dirty tiles say; nothing is released here -- a failure sends the image back to the rounds
        bool bad = false;
        const int nwarps = nth >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        for (int it = 0; it * 32 * nwarps + gw < n; ++it) {
            const int k = (it * 32 + lane) * nwarps + gw;
            bool is_long = false; int i = 0;
            if (k < n && D.wl1[k] >= 0) {
                i = D.wl1[k];
                const SeedRec3 r = C.srec[i];
                const u64 mine = key_of(C, C.seed_prio[i]);
                if (r.cnt + r.bcnt > VERIFY_LONG) is_long = true;
                else if (r.cnt > 0) {
                    ListReader rd; rd.init(r.head);
                    for (int j = 0; j < r.cnt; ++j) bad |= ld_claim0(&C.px[rd.next(C.pool)]) != mine;
                    ListReader rb; rb.init(r.bhead);
                    for (int j = 0; j < r.bcnt; ++j) bad |= !(ld_claim0(&C.px[rb.next(C.pool)]) < mine);
                }
            }
            unsigned lm = __ballot_sync(0xffffffffu, is_long);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const int si = __shfl_sync(0xffffffffu, i, src);
                const u64 mine = key_of(C, C.seed_prio[si]);
                const SeedRec3 r = C.srec[si];
                if (!(walk3_warp(C, r.head, r.cnt, 0, mine, lane) && walk3_warp(C, r.bhead, r.bcnt, 1, mine, lane))) bad = true;
            }
        }
        if (__syncthreads_or(bad) && threadIdx.x == 0) st->recheck = 1;
    } else if (st->mode == 1) {
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += nth) {
            const int e = D.wl1[k];
            if (e >= 0 && !s3_finalize(C, e, D.F)) D.status[0] = OLF_ERR_CAPACITY;
        }
    }
}

// pass 3: the seeds of work list 2 grow, ONE THREAD PER SEED.  Same decisions as s3_begin / s3_step / s3_end in lsd_sticky.h
// (the executable specification, emulated on the host by tests/emul).  A single GPU thread walks a region's sequential
// critical path no faster than a warp does (the path is instruction latency, not memory), but 32 regions per warp cost 32x
// fewer issue slots -- so the kernel is written for few, branch-free instructions per queue entry:
//   * the eight neighbour records (one 32-byte sector each) are fetched together; with ONE claim word per pixel "free" /
//     "mine" / "held" are single 64-bit compares against the key (stamp|prio), evaluated for all eight before any decision;
//   * only the candidates that survive are visited, in the reference's scan order, by a compact loop that reads their
//     (cos, sin) from a per-thread shared-memory slot;
//   * the recent part of the queue lives in a per-thread shared-memory ring; a thread's own claims are visible to its own
//     later (strong) loads, so no side table is needed;
//   * a thread whose region is complete takes the next seed of the list: the lanes of a warp stay busy.
#ifndef GROW_THREADS
#define GROW_THREADS 64
#endif
// The static half of a pixel record (angle, cos, sin, bin: written before the chain starts, never during it) may be read through L1; the claim
// half of the same sector changes under atomics and is always read with a strong load that bypasses L1.  (Measured: .cg / .ca / .nc make no
// difference to the frame rate and the L1 hit rate of this kernel stays at 6 % -- profiles/r02_notes.md.)
#ifndef OLF_LO_LD
#define OLF_LO_LD "ld.global.ca.v4.f32"
#endif
#define GROW_RING 16
struct GrowSmem { float2 nb[8][GROW_THREADS]; unsigned ring[GROW_RING][GROW_THREADS]; };   // (cos, sin) of the 8 neighbours; recent queue

// The last block to finish advances the wave / round state machine.
#ifndef GROW_MIN_BLOCKS
#define GROW_MIN_BLOCKS 1
#endif
// PIPE: software-pipelined walk -- the sixteen loads of the NEXT queue entry travel while the candidates of the current one are decided.
// Same decisions (a region's own claims made meanwhile are patched into the loaded flags); measured on a B200: one chain alone 8 % shorter
// (24.97 -> 23.13 ms for 8 images), no difference at 20 rigs (1041 vs 1037 frames/s) -- so a single frame (latency) runs the pipelined
// instance, a batch (throughput) the plain one with its smaller register footprint (114 vs 135).
template <bool PIPE>
__global__ void __launch_bounds__(GROW_THREADS, GROW_MIN_BLOCKS) k_lsd_grow(const __grid_constant__ GrowBatch B) {
    pdl_enter();
    __shared__ GrowSmem sm;
    __shared__ GrowDev s_dev;
    __shared__ bool s_last;
    PhaseState* st;
    const GrowDev& D = batch_image(B, &s_dev, st);
    if (st->done) return;
    const int wv = st->wave; const unsigned round = st->round; const int mode = st->mode;
    if (wv >= D.plan->n_waves) {                                   // no seeds at all (flat image)
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->done = 1; D.status[1] = (int)round; D.status[2] = D.plan->n_waves; D.status[3] = 1;
            for (int w = wv; w < LSD_MAX_WAVES; ++w) atomicAdd(&B.conv[w], 1u);       // never holds the batch back
            image_done(B);
        }
        return;
    }
    if (mode == 0 && st->wl2_cnt > 0) {
        const Ctx3& C = D.C;
        const unsigned n = st->wl2_cnt;
        const int t = threadIdx.x;
        const int W = C.W, H = C.H;
        const float inv_w = 1.0f / (float)W;
        unsigned* const pool = C.pool;
        bool active = false;
        // per-seed state
        int i = 0; u64 mine = 0;
        unsigned w_head = 0, w_chunk = 0, r_chunk = 0, b_head = kNull, b_chunk = kNull;
        int w_off = 0, r_off = 0, b_off = 0, count = 0, done = 0, bcnt = 0;
        int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
        bool overflow = false, dirty = false;
        float sumdx = 0.f, sumdy = 0.f, u2 = 0.f; double reg_angle = 0.0;
        // The pass runs in instalments: a thread expands at most `budget` queue entries per launch; a region that is not complete
        // then is parked (GrowCont) and the NEXT launch resumes the parked regions densely packed, thread k <- region k.  The few
        // long regions of a round therefore end up together in a few warps instead of keeping one lane busy in every warp
        // (and its block resident) for milliseconds; the decisions of a region are the same, only who executes them changes.
        const unsigned par = st->cont_par;
        const unsigned n_in = st->cont_cnt[par];
        const GrowCont* const c_in = D.cont[par];
        GrowCont* const c_out = D.cont[par ^ 1];
        const unsigned gt = blockIdx.x * GROW_THREADS + t;
        int steps = 0, ring_base = 0;
        const int budget = D.budget;
        if (gt < n_in) {
            const GrowCont c = c_in[gt];
            i = c.i; mine = key_of(C, C.seed_prio[i]);
            w_head = c.w_head; w_chunk = c.w_chunk; r_chunk = c.r_chunk; b_head = c.b_head; b_chunk = c.b_chunk;
            w_off = (int)(c.offs & 0xffu); r_off = (int)((c.offs >> 8) & 0xffu); b_off = (int)((c.offs >> 16) & 0xffu); dirty = (c.offs >> 24) != 0;
            count = c.count; done = c.done; bcnt = c.bcnt; bx0 = c.bx0; by0 = c.by0; bx1 = c.bx1; by1 = c.by1;
            sumdx = c.sumdx; sumdy = c.sumdy; reg_angle = c.reg_angle;
            u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
            ring_base = count;                                   // the shared-memory ring only holds what this thread pushes from now on
            active = true;
        }
        auto push = [&](unsigned pix) {
            if (w_off == kChunk - 1) {
                const unsigned nc = atomicAdd(C.pool_ctr, 1u);
                if (nc >= C.pool_chunks) { overflow = true; return; }
                pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
                pool[(size_t)w_chunk * kChunk + kChunk - 1] = nc;
                w_chunk = nc; w_off = 0;
            }
            pool[(size_t)w_chunk * kChunk + w_off++] = pix;
            sm.ring[count & (GROW_RING - 1)][t] = pix;
            ++count;
        };
        auto record_blocked = [&](unsigned pix) {
            if (b_chunk == kNull || b_off == kChunk - 1) {
                const unsigned nc = atomicAdd(C.pool_ctr, 1u);
                if (nc >= C.pool_chunks) { overflow = true; return; }
                pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
                if (b_chunk == kNull) b_head = nc; else pool[(size_t)b_chunk * kChunk + kChunk - 1] = nc;
                b_chunk = nc; b_off = 0;
            }
            pool[(size_t)b_chunk * kChunk + b_off++] = pix;
            ++bcnt;
        };
        if constexpr (PIPE) {
        // software-pipelined walk; the plain one (one memory round trip per entry ON the critical path) is the else branch
        bool have_cur = false;
        int p_c = 0, px_c = 0, py_c = 0, p_n = 0, px_n = 0, py_n = 0;
        unsigned mask_n = 0, m_free = 0, m_held = 0, m_low = 0;
        u64 cl[8], cl_c = 0; float4 lo[8];
        for (;;) {
            if (!active) {
                if (steps >= budget) break;                           // the rest of the list waits for the next launch
                // ---- s3_begin: the seed pixel has been claimed by the verify pass
                const unsigned k = atomicAdd(&st->wl2_pop, 1u);
                if (k >= n) break;
                ring_base = 0;
                i = D.wl2[k];
                const int seed = C.seed_pix[i];
                mine = key_of(C, C.seed_prio[i]);
                count = 0; done = 0; b_head = kNull; b_chunk = kNull; b_off = 0; bcnt = 0; overflow = false;
                {
                    int sy = __float2int_rd(__fmul_rn((float)seed, inv_w)), sx = seed - sy * W;
                    if (sx < 0) { --sy; sx += W; } else if (sx >= W) { ++sy; sx -= W; }
                    bx0 = bx1 = sx; by0 = by1 = sy;
                }
                const unsigned nc = atomicAdd(C.pool_ctr, 1u);
                if (nc >= C.pool_chunks) overflow = true;
                else {
                    pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
                    w_head = w_chunk = r_chunk = nc; w_off = 0; r_off = 0;
                    push((unsigned)seed);
                    float a, cx, cy; unsigned b;
                    ld_lo(&C.px[seed], a, cx, cy, b);
                    reg_angle = d_mul((double)a, kDegToRads);
                    const float2_t t0 = C.tab_seed[tab_index(C.dabc[seed])];
                    sumdx = t0.x; sumdy = t0.y;
                    u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
                    dirty = false;
                }
                have_cur = false;
                active = true;
            }
            // (1) the next queue entry, if the queue holds one: pop it and issue its sixteen loads; they travel while the candidates of the
            //     current entry are decided below.  What this thread claims meanwhile is missing from the loaded claim words: mask_n collects it
            bool have_next = false;
            if (!overflow && done < count && steps < budget) {
                if (r_off == kChunk - 1) { r_chunk = pool[(size_t)r_chunk * kChunk + kChunk - 1]; r_off = 0; }
                p_n = (int)((count - done <= GROW_RING && done >= ring_base) ? sm.ring[done & (GROW_RING - 1)][t] : pool[(size_t)r_chunk * kChunk + r_off]);
                ++r_off; ++done; ++steps;
                py_n = __float2int_rd(__fmul_rn((float)p_n, inv_w));            // p < 2^24: exact after one correction step
                px_n = p_n - py_n * W;
                if (px_n < 0) { --py_n; px_n += W; } else if (px_n >= W) { ++py_n; px_n -= W; }
                mask_n = 0; have_next = true;                               // (all bookkeeping BEFORE the loads: nothing may touch a register between them and the candidates)
                const PxRec* const base = C.px + p_n;
                // all sixteen loads are issued before anything consumes them (one memory round trip per queue entry): ONE asm
                // block per kind; neighbours outside the image read the guard band or the neighbouring row (valid memory) and
                // are masked out by `vm`
                const PxRec* ra[8];                                      // (the record array has a guard band of W+2 records at both ends)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int nk = k < 4 ? k : k + 1;
                    ra[k] = base + ((nk / 3) - 1) * W + ((nk % 3) - 1);
                }
                asm volatile(
                    "ld.relaxed.gpu.global.u64 %0, [%9];\n\t"
                    "ld.relaxed.gpu.global.u64 %1, [%10];\n\t"
                    "ld.relaxed.gpu.global.u64 %2, [%11];\n\t"
                    "ld.relaxed.gpu.global.u64 %3, [%12];\n\t"
                    "ld.relaxed.gpu.global.u64 %4, [%13];\n\t"
                    "ld.relaxed.gpu.global.u64 %5, [%14];\n\t"
                    "ld.relaxed.gpu.global.u64 %6, [%15];\n\t"
                    "ld.relaxed.gpu.global.u64 %7, [%16];\n\t"
                    "ld.relaxed.gpu.global.u64 %8, [%17];"
                    : "=l"(cl[0]), "=l"(cl[1]), "=l"(cl[2]), "=l"(cl[3]), "=l"(cl[4]), "=l"(cl[5]), "=l"(cl[6]), "=l"(cl[7]), "=l"(cl_c)
                    : "l"(ra[0]), "l"(ra[1]), "l"(ra[2]), "l"(ra[3]), "l"(ra[4]), "l"(ra[5]), "l"(ra[6]), "l"(ra[7]), "l"(base)
                    : "memory");
                asm volatile(
                    OLF_LO_LD " {%0, %1, %2, %3}, [%32+16];\n\t"
                    OLF_LO_LD " {%4, %5, %6, %7}, [%33+16];\n\t"
                    OLF_LO_LD " {%8, %9, %10, %11}, [%34+16];\n\t"
                    OLF_LO_LD " {%12, %13, %14, %15}, [%35+16];\n\t"
                    OLF_LO_LD " {%16, %17, %18, %19}, [%36+16];\n\t"
                    OLF_LO_LD " {%20, %21, %22, %23}, [%37+16];\n\t"
                    OLF_LO_LD " {%24, %25, %26, %27}, [%38+16];\n\t"
                    OLF_LO_LD " {%28, %29, %30, %31}, [%39+16];"
                    : "=f"(lo[0].x), "=f"(lo[0].y), "=f"(lo[0].z), "=f"(lo[0].w), "=f"(lo[1].x), "=f"(lo[1].y), "=f"(lo[1].z), "=f"(lo[1].w),
                      "=f"(lo[2].x), "=f"(lo[2].y), "=f"(lo[2].z), "=f"(lo[2].w), "=f"(lo[3].x), "=f"(lo[3].y), "=f"(lo[3].z), "=f"(lo[3].w),
                      "=f"(lo[4].x), "=f"(lo[4].y), "=f"(lo[4].z), "=f"(lo[4].w), "=f"(lo[5].x), "=f"(lo[5].y), "=f"(lo[5].z), "=f"(lo[5].w),
                      "=f"(lo[6].x), "=f"(lo[6].y), "=f"(lo[6].z), "=f"(lo[6].w), "=f"(lo[7].x), "=f"(lo[7].y), "=f"(lo[7].z), "=f"(lo[7].w)
                    : "l"(ra[0]), "l"(ra[1]), "l"(ra[2]), "l"(ra[3]), "l"(ra[4]), "l"(ra[5]), "l"(ra[6]), "l"(ra[7])
                    : "memory");
            }
            // (2) the candidates of the current entry (its flags were set up at the end of the previous turn)
            if (have_cur && !overflow) {
                // the surviving candidates, in scan order; the region angle changes after every accepted pixel
                unsigned m = m_free | m_held;
                const int p = p_c, px = px_c, py = py_c;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const float2 csk = sm.nb[k][t];
                    const float cxk = csk.x, cyk = csk.y;
                    const int nk = k < 4 ? k : k + 1;
                    const int dx = (nk % 3) - 1, dy = (nk / 3) - 1;
                    const int q = p + dy * W + dx;
                    bool al;
                    {   // isAligned(), lazily: see s3_aligned() in lsd_sticky.h
                        const float dot = f_add(f_mul(sumdx, cxk), f_mul(sumdy, cyk)), d2 = f_mul(dot, dot);
                        const bool fast = C.fast_align && u2 > 1e-3f;
                        if (fast && dot > 0.f && d2 >= f_mul(C.c_hi2, u2)) al = true;
                        else if (fast && (dot <= 0.f || d2 <= f_mul(C.c_lo2, u2))) al = false;
                        else {
                            if (dirty) { reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads); dirty = false; }
                            const float ang_k = __ldcg(&C.px[q].ang);               // rare path: the candidate's angle comes from its record
                            double n_theta = d_sub(reg_angle, d_mul((double)ang_k, kDegToRads));
                            if (n_theta < 0) n_theta = -n_theta;
                            if (n_theta > k3_2Pi) { n_theta = d_sub(n_theta, k2Pi); if (n_theta < 0) n_theta = -n_theta; }
                            al = n_theta <= C.prec;
                        }
                    }
                    if (!al) continue;
                    const int qx = px + dx, qy = py + dy;
                    bx0 = min(bx0, qx); bx1 = max(bx1, qx); by0 = min(by0, qy); by1 = max(by1, qy);
                    if ((m_held >> k) & 1u) { record_blocked((unsigned)q); if (overflow) break; continue; }   // aligned but held by a higher-priority seed
                    if ((m_low >> k) & 1u) mark_dirty_xy(C, round, qx, qy);      // taken from a lower-priority region: it must re-verify
                    red_min64(&C.px[q].claim[0], mine);                          // fire and forget
                    push((unsigned)q);
                    if (overflow) break;
                    if (have_next) {                                             // is q one of the eight neighbours of the entry whose claim words are in flight?
                        const int ex = qx - px_n, ey = qy - py_n;
                        if (ex >= -1 && ex <= 1 && ey >= -1 && ey <= 1 && (ex | ey) != 0) { const int nq = (ey + 1) * 3 + ex + 1; mask_n |= 1u << (nq < 4 ? nq : nq - 1); }
                    }
                    sumdx = f_add(sumdx, cxk);
                    sumdy = f_add(sumdy, cyk);
                    u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
                    dirty = true;
                }
            }
            have_cur = false;
            // (3) the loads have arrived: flags of the next entry, which becomes the current one
            if (have_next && !overflow) {
                // Scheduling guard: the key every flag is compared with depends on ALL sixteen loads, so no consumer can be
                // scheduled between the loads (the assembler otherwise serialises them to save registers: several memory
                // round trips per entry instead of 1).  The guard never fires: bits 63..56 of a claim are 0x00 or 0xFF, bits
                // 31..10 of a bin number are zero.
                // (The angle word of each record is part of the guard although nothing else reads it here: a destination register that is
                // dead would be reused right after the loads, and writing it waits for the load -- before the candidates, not after.  Eight
                // angles whose top byte is 0xFF do not exist: they are degrees in [0, 360) or NOTDEF = -1024.)
                unsigned acc_a = 0, acc_b = 0, acc_c = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc_a |= (unsigned)(cl[k] >> 32); acc_b |= __float_as_uint(lo[k].w); acc_c &= __float_as_uint(lo[k].x); }
                acc_a |= (unsigned)(cl_c >> 32);
                const bool never = ((acc_a >> 24) == 0x55u) || ((acc_b >> 16) == 0x55u) || ((acc_c >> 24) == 0xFFu);
                const u64 mine_g = never ? 0ull : mine;
                // free: c > mine (nobody's, a stale claim of an earlier wave, or a lower-priority claim that is taken over);
                // held: c < mine and not final (NOTDEF pixels are born final); mine: c == mine
                // the entry itself was claimed a while ago without looking at the return value: did it go to a higher-priority seed?
                if (cl_c != mine_g) mark_dirty_xy(C, round, px_n, py_n);
                m_free = 0; m_held = 0; m_low = 0;
                // on the 32-bit halves (the stamp is the top 24 bits of the high word): x < 256 <=> same stamp <=> a claim of this wave
                const unsigned mh = (unsigned)(mine_g >> 32), ml = (unsigned)mine_g;
                // ten instructions per neighbour, written out (the compiler's version of the same tests took sixteen):
                //   x = h ^ mh;  fre = c > mine;  low = fre && x < 256;  hld = !fre && (x | (l ^ ml)) != 0 && h >= 256   (not mine, not final)
#define OLF_NB_FLAGS(K)                                                                                                         \
                asm("{\n\t.reg .pred pf, pl, ph;\n\t.reg .b32 x, y;\n\t"                                                        \
                    "xor.b32 x, %3, %5;\n\t"                                                                                    \
                    "setp.gt.u64 pf, %7, %8;\n\t"                                                                               \
                    "setp.lt.and.u32 pl, x, 256, pf;\n\t"                                                                       \
                    "xor.b32 y, %4, %6;\n\t"                                                                                    \
                    "or.b32 y, y, x;\n\t"                                                                                       \
                    "setp.ne.and.u32 ph, y, 0, !pf;\n\t"                                                                        \
                    "setp.ge.and.u32 ph, %3, 256, ph;\n\t"                                                                      \
                    "@pf or.b32 %0, %0, " #K ";\n\t"                                                                            \
                    "@ph or.b32 %1, %1, " #K ";\n\t"                                                                            \
                    "@pl or.b32 %2, %2, " #K ";\n\t}"                                                                           \
                    : "+r"(m_free), "+r"(m_held), "+r"(m_low)                                                                   \
                    : "r"((unsigned)(cl[K2IDX(K)] >> 32)), "r"((unsigned)cl[K2IDX(K)]), "r"(mh), "r"(ml), "l"(cl[K2IDX(K)]), "l"(mine_g))
#define K2IDX(K) ((K) == 1 ? 0 : (K) == 2 ? 1 : (K) == 4 ? 2 : (K) == 8 ? 3 : (K) == 16 ? 4 : (K) == 32 ? 5 : (K) == 64 ? 6 : 7)
                OLF_NB_FLAGS(1); OLF_NB_FLAGS(2); OLF_NB_FLAGS(4); OLF_NB_FLAGS(8); OLF_NB_FLAGS(16); OLF_NB_FLAGS(32); OLF_NB_FLAGS(64); OLF_NB_FLAGS(128);
#undef OLF_NB_FLAGS
#undef K2IDX
                // which of the 8 neighbours exist (bit k = k-th neighbour in scan order, centre skipped)
                unsigned vm_n = 0x18u;
                if (py_n > 0) vm_n |= 0x07u;
                if (py_n < H - 1) vm_n |= 0xE0u;
                if (px_n == 0) vm_n &= ~0x29u;
                if (px_n == W - 1) vm_n &= ~0x94u;
                m_free &= vm_n & ~mask_n; m_held &= vm_n & ~mask_n; m_low &= ~mask_n;       // what was claimed since the loads left is this thread's
#pragma unroll
                for (int k = 0; k < 8; ++k) sm.nb[k][t] = make_float2(lo[k].y, never ? 0.f : lo[k].z);
                p_c = p_n; px_c = px_n; py_c = py_n; have_cur = true;
            }
            if (overflow || (!have_cur && done >= count)) {
                // ---- s3_end
                SeedRec3 r; r.pad0 = r.pad1 = 0;
                if (overflow) { r.head = kNull; r.cnt = 0; r.bhead = kNull; r.bcnt = 0; r.x0 = r.y0 = r.x1 = r.y1 = 0; D.status[0] = OLF_ERR_CAPACITY; }
                else {
                    if (dirty) reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads);
                    r.head = w_head; r.cnt = count; r.bhead = b_head; r.bcnt = bcnt;
                    r.x0 = (unsigned short)bx0; r.y0 = (unsigned short)by0; r.x1 = (unsigned short)bx1; r.y1 = (unsigned short)by1;
                    C.regang[i] = reg_angle;
                }
                C.srec[i] = r;
                D.tbox[i] = overflow ? kNull : tbox_pack(bx0, by0, bx1, by1);
                if (D.dbg) { atomicAdd(&D.dbg[round * TRACE_REC + 5], 1); atomicAdd(&D.dbg[round * TRACE_REC + 6], count); atomicMax(&D.dbg[round * TRACE_REC + 7], count); }
                active = false;
            } else if (!have_cur && steps >= budget) {
                // ---- out of budget in the middle of a region: park it
                GrowCont c;
                c.i = i; c.w_head = w_head; c.w_chunk = w_chunk; c.r_chunk = r_chunk; c.b_head = b_head; c.b_chunk = b_chunk;
                c.offs = (unsigned)w_off | ((unsigned)r_off << 8) | ((unsigned)b_off << 16) | ((unsigned)dirty << 24);
                c.count = count; c.done = done; c.bcnt = bcnt;
                c.bx0 = (unsigned short)bx0; c.by0 = (unsigned short)by0; c.bx1 = (unsigned short)bx1; c.by1 = (unsigned short)by1;
                c.sumdx = sumdx; c.sumdy = sumdy; c.reg_angle = reg_angle;
                c_out[atomicAdd(&st->cont_cnt[par ^ 1], 1u)] = c;
                active = false;
                break;
            }
        }
        } else {
        for (;;) {
            if (!active) {
                if (steps >= budget) break;                           // the rest of the list waits for the next launch
                // ---- s3_begin: the seed pixel has been claimed by the verify pass
                const unsigned k = atomicAdd(&st->wl2_pop, 1u);
                if (k >= n) break;
                ring_base = 0;
                i = D.wl2[k];
                const int seed = C.seed_pix[i];
                mine = key_of(C, C.seed_prio[i]);
                count = 0; done = 0; b_head = kNull; b_chunk = kNull; b_off = 0; bcnt = 0; overflow = false;
                {
                    int sy = __float2int_rd(__fmul_rn((float)seed, inv_w)), sx = seed - sy * W;
                    if (sx < 0) { --sy; sx += W; } else if (sx >= W) { ++sy; sx -= W; }
                    bx0 = bx1 = sx; by0 = by1 = sy;
                }
                const unsigned nc = atomicAdd(C.pool_ctr, 1u);
                if (nc >= C.pool_chunks) overflow = true;
                else {
                    pool[(size_t)nc * kChunk + kChunk - 1] = kNull;
                    w_head = w_chunk = r_chunk = nc; w_off = 0; r_off = 0;
                    push((unsigned)seed);
                    float a, cx, cy; unsigned b;
                    ld_lo(&C.px[seed], a, cx, cy, b);
                    reg_angle = d_mul((double)a, kDegToRads);
                    const float2_t t0 = C.tab_seed[tab_index(C.dabc[seed])];
                    sumdx = t0.x; sumdy = t0.y;
                    u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
                    dirty = false;
                }
                active = true;
            }
            if (!overflow) {
                // ---- s3_step: expand one queue entry
                if (r_off == kChunk - 1) { r_chunk = pool[(size_t)r_chunk * kChunk + kChunk - 1]; r_off = 0; }
                const int p = (int)((count - done <= GROW_RING && done >= ring_base) ? sm.ring[done & (GROW_RING - 1)][t] : pool[(size_t)r_chunk * kChunk + r_off]);
                ++r_off; ++done; ++steps;
                int py = __float2int_rd(__fmul_rn((float)p, inv_w));           // p < 2^24: exact after one correction step
                int px = p - py * W;
                if (px < 0) { --py; px += W; } else if (px >= W) { ++py; px -= W; }
                // which of the 8 neighbours exist (bit k = k-th neighbour in scan order, centre skipped)
                unsigned vm = 0x18u;
                if (py > 0) vm |= 0x07u;
                if (py < H - 1) vm |= 0xE0u;
                if (px == 0) vm &= ~0x29u;
                if (px == W - 1) vm &= ~0x94u;
                const PxRec* const base = C.px + p;
                // all sixteen loads are issued before anything consumes them (one memory round trip per queue entry): ONE asm
                // block per kind; neighbours outside the image read the guard band or the neighbouring row (valid memory) and
                // are masked out by `vm`
                u64 cl[8], cl_c; float4 lo[8];
                const PxRec* ra[8];                                      // (the record array has a guard band of W+2 records at both ends)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int nk = k < 4 ? k : k + 1;
                    ra[k] = base + ((nk / 3) - 1) * W + ((nk % 3) - 1);
                }
                asm volatile(
                    "ld.relaxed.gpu.global.u64 %0, [%9];\n\t"
                    "ld.relaxed.gpu.global.u64 %1, [%10];\n\t"
                    "ld.relaxed.gpu.global.u64 %2, [%11];\n\t"
                    "ld.relaxed.gpu.global.u64 %3, [%12];\n\t"
                    "ld.relaxed.gpu.global.u64 %4, [%13];\n\t"
                    "ld.relaxed.gpu.global.u64 %5, [%14];\n\t"
                    "ld.relaxed.gpu.global.u64 %6, [%15];\n\t"
                    "ld.relaxed.gpu.global.u64 %7, [%16];\n\t"
                    "ld.relaxed.gpu.global.u64 %8, [%17];"
                    : "=l"(cl[0]), "=l"(cl[1]), "=l"(cl[2]), "=l"(cl[3]), "=l"(cl[4]), "=l"(cl[5]), "=l"(cl[6]), "=l"(cl[7]), "=l"(cl_c)
                    : "l"(ra[0]), "l"(ra[1]), "l"(ra[2]), "l"(ra[3]), "l"(ra[4]), "l"(ra[5]), "l"(ra[6]), "l"(ra[7]), "l"(base)
                    : "memory");
                asm volatile(
                    OLF_LO_LD " {%0, %1, %2, %3}, [%32+16];\n\t"
                    OLF_LO_LD " {%4, %5, %6, %7}, [%33+16];\n\t"
                    OLF_LO_LD " {%8, %9, %10, %11}, [%34+16];\n\t"
                    OLF_LO_LD " {%12, %13, %14, %15}, [%35+16];\n\t"
                    OLF_LO_LD " {%16, %17, %18, %19}, [%36+16];\n\t"
                    OLF_LO_LD " {%20, %21, %22, %23}, [%37+16];\n\t"
                    OLF_LO_LD " {%24, %25, %26, %27}, [%38+16];\n\t"
                    OLF_LO_LD " {%28, %29, %30, %31}, [%39+16];"
                    : "=f"(lo[0].x), "=f"(lo[0].y), "=f"(lo[0].z), "=f"(lo[0].w), "=f"(lo[1].x), "=f"(lo[1].y), "=f"(lo[1].z), "=f"(lo[1].w),
                      "=f"(lo[2].x), "=f"(lo[2].y), "=f"(lo[2].z), "=f"(lo[2].w), "=f"(lo[3].x), "=f"(lo[3].y), "=f"(lo[3].z), "=f"(lo[3].w),
                      "=f"(lo[4].x), "=f"(lo[4].y), "=f"(lo[4].z), "=f"(lo[4].w), "=f"(lo[5].x), "=f"(lo[5].y), "=f"(lo[5].z), "=f"(lo[5].w),
                      "=f"(lo[6].x), "=f"(lo[6].y), "=f"(lo[6].z), "=f"(lo[6].w), "=f"(lo[7].x), "=f"(lo[7].y), "=f"(lo[7].z), "=f"(lo[7].w)
                    : "l"(ra[0]), "l"(ra[1]), "l"(ra[2]), "l"(ra[3]), "l"(ra[4]), "l"(ra[5]), "l"(ra[6]), "l"(ra[7])
                    : "memory");
                // Scheduling guard: the key every flag is compared with depends on ALL sixteen loads, so no consumer can be
                // scheduled between the loads (the assembler otherwise serialises them to save registers: several memory
                // round trips per entry instead of 1).  The guard never fires: bits 63..56 of a claim are 0x00 or 0xFF, bits
                // 31..10 of a bin number are zero.
                unsigned acc_a = 0, acc_b = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { acc_a |= (unsigned)(cl[k] >> 32); acc_b |= __float_as_uint(lo[k].w); }
                acc_a |= (unsigned)(cl_c >> 32);
                const bool never = ((acc_a >> 24) == 0x55u) || ((acc_b >> 16) == 0x55u);
                const u64 mine_g = never ? 0ull : mine;
                // free: c > mine (nobody's, a stale claim of an earlier wave, or a lower-priority claim that is taken over);
                // held: c < mine and not final (NOTDEF pixels are born final); mine: c == mine
                // the entry itself was claimed a while ago without looking at the return value: did it go to a higher-priority seed?
                if (cl_c != mine_g) mark_dirty_xy(C, round, px, py);
                unsigned m_free = 0, m_held = 0, m_low = 0;
                // on the 32-bit halves (the stamp is the top 24 bits of the high word): x < 256 <=> same stamp <=> a claim of this wave
                const unsigned mh = (unsigned)(mine_g >> 32), ml = (unsigned)mine_g;
                // ten instructions per neighbour, written out (the compiler's version of the same tests took sixteen):
                //   x = h ^ mh;  fre = c > mine;  low = fre && x < 256;  hld = !fre && (x | (l ^ ml)) != 0 && h >= 256   (not mine, not final)
#define OLF_NB_FLAGS(K)                                                                                                         \
                asm("{\n\t.reg .pred pf, pl, ph;\n\t.reg .b32 x, y;\n\t"                                                        \
                    "xor.b32 x, %3, %5;\n\t"                                                                                    \
                    "setp.gt.u64 pf, %7, %8;\n\t"                                                                               \
                    "setp.lt.and.u32 pl, x, 256, pf;\n\t"                                                                       \
                    "xor.b32 y, %4, %6;\n\t"                                                                                    \
                    "or.b32 y, y, x;\n\t"                                                                                       \
                    "setp.ne.and.u32 ph, y, 0, !pf;\n\t"                                                                        \
                    "setp.ge.and.u32 ph, %3, 256, ph;\n\t"                                                                      \
                    "@pf or.b32 %0, %0, " #K ";\n\t"                                                                            \
                    "@ph or.b32 %1, %1, " #K ";\n\t"                                                                            \
                    "@pl or.b32 %2, %2, " #K ";\n\t}"                                                                           \
                    : "+r"(m_free), "+r"(m_held), "+r"(m_low)                                                                   \
                    : "r"((unsigned)(cl[K2IDX(K)] >> 32)), "r"((unsigned)cl[K2IDX(K)]), "r"(mh), "r"(ml), "l"(cl[K2IDX(K)]), "l"(mine_g))
#define K2IDX(K) ((K) == 1 ? 0 : (K) == 2 ? 1 : (K) == 4 ? 2 : (K) == 8 ? 3 : (K) == 16 ? 4 : (K) == 32 ? 5 : (K) == 64 ? 6 : 7)
                OLF_NB_FLAGS(1); OLF_NB_FLAGS(2); OLF_NB_FLAGS(4); OLF_NB_FLAGS(8); OLF_NB_FLAGS(16); OLF_NB_FLAGS(32); OLF_NB_FLAGS(64); OLF_NB_FLAGS(128);
#undef OLF_NB_FLAGS
#undef K2IDX
                m_free &= vm; m_held &= vm;
#pragma unroll
                for (int k = 0; k < 8; ++k) sm.nb[k][t] = make_float2(lo[k].y, never ? 0.f : lo[k].z);
                // the surviving candidates, in scan order; the region angle changes after every accepted pixel
                unsigned m = m_free | m_held;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const float2 csk = sm.nb[k][t];
                    const float cxk = csk.x, cyk = csk.y;
                    const int nk = k < 4 ? k : k + 1;
                    const int dx = (nk % 3) - 1, dy = (nk / 3) - 1;
                    const int q = p + dy * W + dx;
                    bool al;
                    {   // isAligned(), lazily: see s3_aligned() in lsd_sticky.h
                        const float dot = f_add(f_mul(sumdx, cxk), f_mul(sumdy, cyk)), d2 = f_mul(dot, dot);
                        const bool fast = C.fast_align && u2 > 1e-3f;
                        if (fast && dot > 0.f && d2 >= f_mul(C.c_hi2, u2)) al = true;
                        else if (fast && (dot <= 0.f || d2 <= f_mul(C.c_lo2, u2))) al = false;
                        else {
                            if (dirty) { reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads); dirty = false; }
                            const float ang_k = __ldcg(&C.px[q].ang);               // rare path: the candidate's angle comes from its record
                            double n_theta = d_sub(reg_angle, d_mul((double)ang_k, kDegToRads));
                            if (n_theta < 0) n_theta = -n_theta;
                            if (n_theta > k3_2Pi) { n_theta = d_sub(n_theta, k2Pi); if (n_theta < 0) n_theta = -n_theta; }
                            al = n_theta <= C.prec;
                        }
                    }
                    if (!al) continue;
                    const int qx = px + dx, qy = py + dy;
                    bx0 = min(bx0, qx); bx1 = max(bx1, qx); by0 = min(by0, qy); by1 = max(by1, qy);
                    if ((m_held >> k) & 1u) { record_blocked((unsigned)q); if (overflow) break; continue; }   // aligned but held by a higher-priority seed
                    if ((m_low >> k) & 1u) mark_dirty_xy(C, round, qx, qy);      // taken from a lower-priority region: it must re-verify
                    red_min64(&C.px[q].claim[0], mine);                          // fire and forget
                    push((unsigned)q);
                    if (overflow) break;
                    sumdx = f_add(sumdx, cxk);
                    sumdy = f_add(sumdy, cyk);
                    u2 = f_add(f_mul(sumdx, sumdx), f_mul(sumdy, sumdy));
                    dirty = true;
                }
            }
            if (overflow || done >= count) {
                // ---- s3_end
                SeedRec3 r; r.pad0 = r.pad1 = 0;
                if (overflow) { r.head = kNull; r.cnt = 0; r.bhead = kNull; r.bcnt = 0; r.x0 = r.y0 = r.x1 = r.y1 = 0; D.status[0] = OLF_ERR_CAPACITY; }
                else {
                    if (dirty) reg_angle = d_mul((double)olf::lsd::fast_atan2_deg(sumdy, sumdx), kDegToRads);
                    r.head = w_head; r.cnt = count; r.bhead = b_head; r.bcnt = bcnt;
                    r.x0 = (unsigned short)bx0; r.y0 = (unsigned short)by0; r.x1 = (unsigned short)bx1; r.y1 = (unsigned short)by1;
                    C.regang[i] = reg_angle;
                }
                C.srec[i] = r;
                D.tbox[i] = overflow ? kNull : tbox_pack(bx0, by0, bx1, by1);
                if (D.dbg) { atomicAdd(&D.dbg[round * TRACE_REC + 5], 1); atomicAdd(&D.dbg[round * TRACE_REC + 6], count); atomicMax(&D.dbg[round * TRACE_REC + 7], count); }
                active = false;
            } else if (steps >= budget) {
                // ---- out of budget in the middle of a region: park it
                GrowCont c;
                c.i = i; c.w_head = w_head; c.w_chunk = w_chunk; c.r_chunk = r_chunk; c.b_head = b_head; c.b_chunk = b_chunk;
                c.offs = (unsigned)w_off | ((unsigned)r_off << 8) | ((unsigned)b_off << 16) | ((unsigned)dirty << 24);
                c.count = count; c.done = done; c.bcnt = bcnt;
                c.bx0 = (unsigned short)bx0; c.by0 = (unsigned short)by0; c.bx1 = (unsigned short)bx1; c.by1 = (unsigned short)by1;
                c.sumdx = sumdx; c.sumdy = sumdy; c.reg_angle = reg_angle;
                c_out[atomicAdd(&st->cont_cnt[par ^ 1], 1u)] = c;
                active = false;
                break;
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
    int err = *(volatile int*)&D.status[0];
    st->ticket = 0; st->launches += 1;
    if (st->launches > 6000 && err == 0) { D.status[0] = OLF_ERR_INTERNAL; err = OLF_ERR_INTERNAL; }     // a WHILE graph must end whatever happens
    // The images of a batch advance through their waves TOGETHER: an image that has converged waits (mode 2) until every image
    // of the batch has, so the expensive first rounds of a wave -- whose critical path is the longest region -- coincide in
    // the same launches instead of adding up along the chain.
    const unsigned nb = (unsigned)B.n;
    if (mode == 0 && st->wl2_cnt > 0) {
        // parked regions or seeds nobody has taken yet: the grow pass of this round goes on in the next launch
        const unsigned par = st->cont_par;
        const unsigned parked = *(volatile unsigned*)&st->cont_cnt[par ^ 1];
        const bool seeds_left = *(volatile unsigned*)&st->wl2_pop < st->wl2_cnt;
        st->cont_cnt[par] = 0;
        if (parked > 0 || seeds_left) { st->cont_par = par ^ 1; st->growing = 1; __threadfence(); return; }
        st->growing = 0;
    }
    if (mode == 0) {
        if (D.dbg) D.dbg[round * TRACE_REC + 3] = (int)(gtime() & 0x7fffffff);
        const bool changed = *(volatile unsigned*)&st->changed != 0;
        if (!changed || round + 2 >= D.max_rounds || err != 0) {              // work list 1 is kept for the next passes
            const unsigned c = wv < LSD_MAX_WAVES ? atomicAdd(&B.conv[wv], 1u) + 1u : nb;
            st->mode = (c >= nb) ? 3 : 2;
        } else { st->round = round + 1; st->wl1_cnt = 0; }
        st->wl2_cnt = 0; st->wl2_pop = 0; st->changed = 0; st->force = 0; st->recheck = 0;
    } else if (mode == 2) {
        if (*(volatile unsigned*)&B.conv[wv] >= nb) st->mode = 3;
    } else if (mode == 3) {
        const bool pretend = D.test_recheck && st->pad0 != (unsigned)(wv + 1);      // once per wave
        if (pretend) st->pad0 = (unsigned)(wv + 1);
        if ((*(volatile unsigned*)&st->recheck != 0 || pretend) && err == 0 && round + 2 < D.max_rounds) {
            st->mode = 0; st->round = round + 1; st->wl1_cnt = 0; st->force = 1; st->recheck = 0;      // back to the rounds, everything dirty
        } else st->mode = 1;
    } else {
        st->mode = 0; st->round = round + 1; st->wave = wv + 1; st->wave_first_round = round + 1; st->wl1_cnt = 0; st->wl0_cnt = 0;
        *D.C.pool_ctr = 0;                                          // one bump pool per wave (lists live until the wave is final)
        if (wv + 1 >= D.plan->n_waves || err != 0) {
            st->done = 1; D.status[1] = (int)(round + 1); D.status[2] = D.plan->n_waves; D.status[3] = 1;
            for (int w = wv + 1; w < LSD_MAX_WAVES; ++w) atomicAdd(&B.conv[w], 1u);   // this image has no further waves
            image_done(B);
        }
    }
    __threadfence();
}

// ---- region2rect ---------------------------------------------------------------------------
struct RectRec { double x, y, theta; u64 prio; };
// Warp per region.  The reference's sums are sequential double additions in BFS order (not associative), so the additions
// stay a single chain -- but everything that feeds them (pixel fetch, gradient weight sqrt/div, products) is computed by
// the 32 lanes in parallel and staged in shared memory; lane 0 only walks the three add chains.
// per-image pointers of the rectangle passes and of the LBD passes (blockIdx.y = image)
struct LbdLine;
struct PostBatch {
    const LsdRegion* regs[LSD_PRE_MAX]; const unsigned* nreg[LSD_PRE_MAX]; const unsigned* final_pool[LSD_PRE_MAX]; const short2_t* dabc[LSD_PRE_MAX];
    RectRec* rect_host[LSD_PRE_MAX]; const double2* dir_host[LSD_PRE_MAX]; float4* seg_host[LSD_PRE_MAX];
    const int* status[LSD_PRE_MAX]; int* status_host[LSD_PRE_MAX]; unsigned* nreg_host[LSD_PRE_MAX];
    int nr[LSD_PRE_MAX];                                  // rect_b / LBD: regions (lines) of the image
    const uint8_t* lbd_blur[LSD_PRE_MAX]; short2_t* grad[LSD_PRE_MAX]; const LbdLine* lines[LSD_PRE_MAX]; float4* rowsum[LSD_PRE_MAX]; uint8_t* desc_host[LSD_PRE_MAX];
};
__global__ void __launch_bounds__(256) k_lsd_rect_a(const __grid_constant__ PostBatch PB, unsigned cap, int W, double prec) {
    const LsdRegion* __restrict__ regs = PB.regs[blockIdx.y]; const unsigned* __restrict__ nreg = PB.nreg[blockIdx.y];
    const unsigned* __restrict__ final_pool = PB.final_pool[blockIdx.y]; const short2_t* __restrict__ dabc = PB.dabc[blockIdx.y];
    RectRec* __restrict__ out_host = PB.rect_host[blockIdx.y];
    __shared__ double sh[8][3][32];
    const unsigned n = min(*nreg, cap);
    // what the host looks at after the chain, written straight to mapped host memory (no copies to enqueue per image)
    if (blockIdx.x == 0 && threadIdx.x < 4) { PB.status_host[blockIdx.y][threadIdx.x] = PB.status[blockIdx.y][threadIdx.x]; if (threadIdx.x == 0) *PB.nreg_host[blockIdx.y] = *nreg; }
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
__global__ void __launch_bounds__(256) k_lsd_rect_b(const __grid_constant__ PostBatch PB, int W, double scale) {
    const LsdRegion* __restrict__ regs = PB.regs[blockIdx.y]; const unsigned* __restrict__ final_pool = PB.final_pool[blockIdx.y];
    const RectRec* __restrict__ rect = PB.rect_host[blockIdx.y]; const double2* __restrict__ dir = PB.dir_host[blockIdx.y];
    float4* __restrict__ seg_host = PB.seg_host[blockIdx.y];
    const int n = PB.nr[blockIdx.y];
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
__global__ void __launch_bounds__(256) k_sobel3(const __grid_constant__ PostBatch PB, int w, int h, int pitch) {
    const uint8_t* __restrict__ img = PB.lbd_blur[blockIdx.z]; short2_t* __restrict__ out = PB.grad[blockIdx.z];
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
__global__ void __launch_bounds__(256) k_lbd_rows(const __grid_constant__ PostBatch PB, int imw, int imh) {
    const LbdLine* __restrict__ lines = PB.lines[blockIdx.y]; const short2_t* __restrict__ grad = PB.grad[blockIdx.y];
    float4* __restrict__ rowsum = PB.rowsum[blockIdx.y];
    const int n = PB.nr[blockIdx.y];
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
__global__ void __launch_bounds__(64) k_lbd_fold(const __grid_constant__ PostBatch PB) {
    const float4* __restrict__ rowsum = PB.rowsum[blockIdx.y]; uint8_t* __restrict__ desc_host = PB.desc_host[blockIdx.y];
    const int n = PB.nr[blockIdx.y];
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
    cudaStream_t stream = nullptr; bool owns_stream = true;
    // derived LSD constants
    double prec, rho; int n2_thresh; int blur_k; int blur_q[4];
    // size-dependent
    int img_w = 0, img_h = 0, ipitch = 0, W = 0, H = 0, wpitch = 0, S = 0, min_reg_size = 0;
    PinBuf<uint8_t> img_stage;
    DevBuf<uint8_t> img, blurred, scaled, lbd_blur;
    DevBuf<ExCoef> coef; size_t coef_y_off = 0;
    DevBuf<float> ang; DevBuf<short2_t> dabc;
    DevBuf<u64> seed_prio;
    DevBuf<int> seed_pix, n2max, status, wl0, wl1, wl2;
    DevBuf<unsigned> hist, bin_start, cursor, pool, ctrs, final_pool;
    DevBuf<SeedRec3> srec0;
    DevBuf<unsigned> conv, dirty;
    int tile_wpr = 0, dirty_words = 0;
    DevBuf<double> regang;
    DevBuf<LsdPlan> plan;
    DevBuf<LsdRegion> regs;
    DevBuf<float2_t> tab_seed, tab_acc, cs;
    DevBuf<PhaseState> phase;
    DevBuf<PxRec> px;
    DevBuf<GrowCont> cont;
    DevBuf<unsigned char> sstate; DevBuf<unsigned> tbox;
    TmaSet<1> tmaps[2]; bool has_tma = false;      // [0]: input image, 7x7 box (LSD pre-blur); [1]: input image, 5x5 box (LBD blur)
    cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr; unsigned long long graph_key = 0; int use_graph = 1;
    int grow_budget = 1 << 30;          // OLF_LSD_GROW_BUDGET: queue entries per thread and grow launch (default: no limit)
    int phase_batch = 52;
    bool trace = false;
    int first_wave = 4096, first_wave_latency = 262144, wave_growth = 16;
    int pipeline = -1;                                           // grow kernel: -1 = pipelined walk for single frames only, 0 / 1 forced
    DevBuf<int> dbg;
    unsigned pool_chunks = 0, reg_cap = 0, max_rounds = 4096;
    int scan_blocks = 0, verify_blocks = 0, scan_blocks_wide = 0, verify_blocks_wide = 0, grow_blocks_wide = 0, grow_blocks_narrow = 0;
    PinBuf<RectRec> rect_host; PinBuf<double2> dir_host; PinBuf<float4> seg_host; PinBuf<int> status_host; PinBuf<unsigned> nreg_host;
    // LBD
    DevBuf<short2_t> grad;
    PinBuf<LbdLine> lbd_lines; DevBuf<float4> rowsum; PinBuf<uint8_t> desc_host;
    int lbd_cap = 0;
    cudaEvent_t ev_grow0 = nullptr, ev_grow1 = nullptr;
    SyncEvent sync;                      // owned by the handle: callers may be short-lived threads
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

LineImpl* line_create(const olf_line_params* p, int device, cudaStream_t ext_stream, bool blocking_sync) {
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
    // ext_stream: a rig drives all its line extractors through one stream (line_extract_batch); streams are a scarce
    // resource (32 hardware work queues per device), so the borrowers do not create their own
    bool ok = true;
    if (ext_stream) { h->stream = ext_stream; h->owns_stream = false; }
    else ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev_grow0) == cudaSuccess && cudaEventCreate(&h->ev_grow1) == cudaSuccess && h->sync.create(blocking_sync) == cudaSuccess;
    ok = ok && h->tab_seed.ensure(ts.size()) == OLF_OK && h->tab_acc.ensure(ta.size()) == OLF_OK;
    ok = ok && cudaMemcpy(h->tab_seed.p, ts.data(), ts.size() * sizeof(float2_t), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->tab_acc.p, ta.data(), ta.size() * sizeof(float2_t), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpyToSymbol(c_gaussG, gG, sizeof(gG)) == cudaSuccess && cudaMemcpyToSymbol(c_gaussL, gL, sizeof(gL)) == cudaSuccess;
    ok = ok && cudaMemcpyToSymbol(c_lbd_comb, LBD_COMB, sizeof(LBD_COMB)) == cudaSuccess;
    int sms = 0;
    ok = ok && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    if (const char* e = getenv("OLF_LSD_FIRST_WAVE")) h->first_wave = h->first_wave_latency = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_WAVE_GROWTH")) h->wave_growth = std::max(2, atoi(e));
    h->trace = getenv("OLF_LSD_TRACE") != nullptr;                // per-round trace of the grow kernel (tools/lsd_trace.py)
    ok = ok && h->phase.ensure(1) == OLF_OK;
    if (const char* e = getenv("OLF_LSD_PHASE_BATCH")) h->phase_batch = std::max(4, atoi(e));
    if (!ok) { set_last_error(std::string("olf_line_create: ") + cudaGetErrorString(cudaGetLastError())); delete h; return nullptr; }
    // grids of the three region-growing passes (one thread per seed; see k_lsd_scan / k_lsd_verify / k_lsd_grow)
    // Grids per image and pass.  Most launches of the chain find little or nothing to do (tail rounds, images waiting for the
    // batch) and every pass is latency-bound, so the grids of a batch are small.  Two settings: a region's critical
    // path is sequential, so MORE threads only shorten a round while fewer threads keep the lanes of a warp busy (a lane
    // whose region is complete takes the next seed) -- measured 4.9 active lanes per warp instruction with one seed per
    // thread.  A single frame (<= 2 images) gets the wide grid (latency), a batch the narrow one (throughput).
    h->scan_blocks = 64; h->verify_blocks = 48; h->scan_blocks_wide = 2 * sms; h->verify_blocks_wide = sms;
    h->grow_blocks_wide = std::max(1, sms * 64 / GROW_THREADS); h->grow_blocks_narrow = std::max(1, 20 * 64 / GROW_THREADS);
    if (const char* e = getenv("OLF_LSD_SCAN_BLOCKS")) h->scan_blocks = h->scan_blocks_wide = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_VERIFY_BLOCKS")) h->verify_blocks = h->verify_blocks_wide = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_GROW_BLOCKS")) h->grow_blocks_wide = h->grow_blocks_narrow = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_GROW_BUDGET")) h->grow_budget = std::max(1, atoi(e));
    if (const char* e = getenv("OLF_LSD_GRAPH")) h->use_graph = atoi(e);
    if (const char* e = getenv("OLF_LSD_PIPELINE")) h->pipeline = atoi(e) != 0;      // 0 / 1 forces the plain / pipelined grow kernel (default: by batch size)
    ok = h->cont.ensure((size_t)2 * std::max(h->grow_blocks_wide, h->grow_blocks_narrow) * GROW_THREADS) == OLF_OK;
    if (!ok) { delete h; return nullptr; }
    return h;
}

void line_destroy(LineImpl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream && h->owns_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    if (h->ev_grow0) cudaEventDestroy(h->ev_grow0);
    if (h->ev_grow1) cudaEventDestroy(h->ev_grow1);
    h->sync.destroy();
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    h->img_stage.release(); h->img.release(); h->blurred.release(); h->scaled.release(); h->lbd_blur.release(); h->coef.release();
    h->ang.release(); h->dabc.release(); h->seed_prio.release(); h->seed_pix.release(); h->n2max.release(); h->status.release();
    h->wl0.release(); h->wl1.release(); h->wl2.release(); h->hist.release(); h->bin_start.release(); h->cursor.release(); h->pool.release(); h->ctrs.release();
    h->final_pool.release(); h->conv.release(); h->dirty.release(); h->srec0.release(); h->regang.release(); h->plan.release(); h->regs.release();
    h->tab_seed.release(); h->tab_acc.release(); h->cs.release(); h->phase.release(); h->px.release(); h->dbg.release(); h->cont.release(); h->sstate.release(); h->tbox.release();
    h->rect_host.release(); h->dir_host.release(); h->seg_host.release(); h->status_host.release(); h->nreg_host.release();
    h->grad.release(); h->lbd_lines.release(); h->rowsum.release(); h->desc_host.release();
    delete h;
}

static LevelTable single_level(int w, int hgt, int pitch) {
    LevelTable T; memset(&T, 0, sizeof(T));
    T.n = 1; T.w[0] = w; T.h[0] = hgt; T.pitch[0] = pitch; T.off[0] = 0;
    T.tiles_x[0] = (w + TILE_W - 1) / TILE_W; T.tile_start[0] = 0;
    T.tile_start[1] = T.tiles_x[0] * ((hgt + TILE_H - 1) / TILE_H);
    return T;
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
    h->tile_wpr = (((h->W + 31) >> 5) + 31) / 32; h->dirty_words = h->tile_wpr * ((h->H + 31) >> 5);
    if ((rc = h->dirty.ensure((size_t)2 * h->dirty_words))) return rc;
    h->pool_chunks = (unsigned)std::max<size_t>(S / 2, 1u << 16);       // 16 px of list space per image pixel per round
    h->reg_cap = (unsigned)(S / std::max(h->min_reg_size, 1) + 16);
    if ((rc = h->ang.ensure(S)) || (rc = h->dabc.ensure(S)) || (rc = h->seed_prio.ensure(S)) || (rc = h->seed_pix.ensure(S)) ||
        (rc = h->conv.ensure(LSD_MAX_WAVES + 1)) || (rc = h->srec0.ensure(S)) || (rc = h->wl0.ensure(S)) || (rc = h->wl1.ensure(S)) || (rc = h->wl2.ensure(S)) ||
        (rc = h->regang.ensure(S)) || (rc = h->final_pool.ensure(S)) || (rc = h->sstate.ensure(S)) || (rc = h->tbox.ensure(S)) ||
        (rc = h->n2max.ensure(1)) || (rc = h->status.ensure(4)) || (rc = h->hist.ensure(1024)) || (rc = h->bin_start.ensure(1024)) ||
        (rc = h->cursor.ensure(1024)) || (rc = h->ctrs.ensure(4)) || (rc = h->plan.ensure(1)) ||
        (rc = h->pool.ensure((size_t)h->pool_chunks * kChunk)) ||
        (rc = h->dbg.ensure((size_t)h->max_rounds * TRACE_REC)) || (rc = h->px.ensure((size_t)S + 2 * (h->W + 2))) ||
        (rc = h->regs.ensure(h->reg_cap)) || (rc = h->rect_host.ensure(h->reg_cap)) || (rc = h->dir_host.ensure(h->reg_cap)) ||
        (rc = h->seg_host.ensure(h->reg_cap)) || (rc = h->status_host.ensure(4)) || (rc = h->nreg_host.ensure(1))) return rc;
    {
        const LevelTable T1 = single_level(w, hgt, h->ipitch);
        h->has_tma = h->blur_k == 7 && build_level_maps(h->img.p, T1, 7, h->tmaps[0].m) && build_level_maps(h->img.p, T1, 5, h->tmaps[1].m);
    }
    // guard bands of the pixel records: zero = "final" claims, never a candidate (the grow pass may read, never use them)
    OLF_CUDA(cudaMemset(h->px.p, 0, ((size_t)S + 2 * (h->W + 2)) * sizeof(PxRec)));
    h->img_w = w; h->img_h = hgt;
    return OLF_OK;
}


static int line_upload(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, cudaStream_t s) {
    int rc = line_ensure_size(h, w, hgt);
    if (rc) return rc;
    if (on_device) OLF_CUDA(cudaMemcpy2DAsync(h->img.p, h->ipitch, img, stride, w, hgt, cudaMemcpyDeviceToDevice, s));
    else {
        cudaPointerAttributes a;
        const bool pinned = cudaPointerGetAttributes(&a, img) == cudaSuccess && a.type == cudaMemoryTypeHost;
        if (!pinned) cudaGetLastError();
        if (pinned) OLF_CUDA(cudaMemcpy2DAsync(h->img.p, h->ipitch, img, stride, w, hgt, cudaMemcpyHostToDevice, s));
        else {      // pageable caller memory goes through the handle's pinned staging buffer
            for (int y = 0; y < hgt; ++y) memcpy(h->img_stage.p + (size_t)y * w, img + (size_t)y * stride, w);
            OLF_CUDA(cudaMemcpy2DAsync(h->img.p, h->ipitch, h->img_stage.p, w, w, hgt, cudaMemcpyHostToDevice, s));
        }
    }
    return OLF_OK;
}

// (the per-image part that stays on the host: this image's descriptor for the batched passes follows in lsd_fill_dev)
// LSD on the uploaded images, stage 1: everything before region growing, ONE launch per stage for all images of the call
static int lsd_enqueue_pre_batch(LineImpl* const* hs, int n, cudaStream_t s) {
    LineImpl* h0 = hs[0];
    const int w = h0->img_w, hgt = h0->img_h, W = h0->W, H = h0->H, S = h0->S;
    for (int k = 1; k < n; ++k)
        if (hs[k]->img_w != w || hs[k]->img_h != hgt || hs[k]->blur_k != h0->blur_k || hs[k]->P.lsd_n_bins != h0->P.lsd_n_bins || hs[k]->n2_thresh != h0->n2_thresh ||
            hs[k]->P.lsd_scale != h0->P.lsd_scale || hs[k]->has_tma != h0->has_tma) { set_last_error("line batch: extractors of one batch must share size and parameters"); return OLF_ERR_ARG; }
    PreBatch PB; memset(&PB, 0, sizeof(PB)); PB.n = n;
    for (int k = 0; k < n; ++k) {
        LineImpl* h = hs[k];
        PB.work_src[k] = h->blur_k ? h->blurred.p : h->img.p; PB.scaled[k] = h->scaled.p;
        PB.grad_src[k] = h->blur_k ? h->scaled.p : h->img.p;
        PB.ang[k] = h->ang.p; PB.dabc[k] = h->dabc.p; PB.px[k] = h->px.p + h->W + 2; PB.n2max[k] = h->n2max.p; PB.hist[k] = h->hist.p;
        PB.bin_start[k] = h->bin_start.p; PB.cursor[k] = h->cursor.p; PB.seed_pix[k] = h->seed_pix.p; PB.seed_prio[k] = h->seed_prio.p; PB.plan[k] = h->plan.p;
        PB.ctrs[k] = h->ctrs.p; PB.status[k] = h->status.p; PB.phase[k] = h->phase.p;
        if (h->trace) OLF_CUDA(cudaMemsetAsync(h->dbg.p, 0, (size_t)h->max_rounds * TRACE_REC * sizeof(int), s));
    }
    k_lsd_reset<<<n, 256, 0, s>>>(PB);
    int wp = h0->ipitch;
    if (h0->blur_k) {
        const LevelTable T = single_level(w, hgt, h0->ipitch);
        const int nt = T.tile_start[1];
        BlurBatch bb; memset(&bb, 0, sizeof(bb));
        for (int k = 0; k < n; ++k) { bb.src[k] = hs[k]->img.p; bb.dst[k] = hs[k]->blurred.p; }
        if (h0->blur_k == 7 && h0->has_tma) {
            static thread_local TmaSet<LSD_PRE_MAX> M;
            for (int k = 0; k < n; ++k) M.m[k] = hs[k]->tmaps[0].m[0];
            k_blur_q8_tma<7, LSD_PRE_MAX><<<dim3(nt, n), 256, 0, s>>>(bb, M, T, h0->blur_q[0], h0->blur_q[1], h0->blur_q[2], h0->blur_q[3]);
        }
        else if (h0->blur_k == 7) k_blur_q8<7><<<dim3(nt, n), 256, 0, s>>>(bb, T, h0->blur_q[0], h0->blur_q[1], h0->blur_q[2], h0->blur_q[3]);
        else if (h0->blur_k == 5) k_blur_q8<5><<<dim3(nt, n), 256, 0, s>>>(bb, T, h0->blur_q[0], h0->blur_q[1], h0->blur_q[2], 0);
        else k_blur_q8<3><<<dim3(nt, n), 256, 0, s>>>(bb, T, h0->blur_q[0], h0->blur_q[1], 0, 0);
        dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8, n);
        k_resize_exact<<<g, b, 0, s>>>(PB, w, hgt, h0->ipitch, W, H, h0->wpitch, h0->coef.p, h0->coef.p + h0->coef_y_off);
        wp = h0->wpitch;
    }
    {
        dim3 g((W + 31) / 32, (H + 7) / 8, n);
        k_lsd_grad<<<g, 256, 0, s>>>(PB, W, H, wp, h0->n2_thresh, h0->tab_acc.p);
    }
    const int nb = h0->P.lsd_n_bins;
    k_lsd_hist<<<dim3(296, n), 256, nb * sizeof(unsigned), s>>>(PB, S, nb);
    // Wave plan: with persistent claims a round costs little, so a BIG first wave (few waves, few first rounds whose critical
    // path is the longest region) gives the lowest latency (5.7 instead of 10.5 ms per image) at the price of ~40 % more
    // speculative growth; a batch, which is throughput-bound, keeps the small first wave.
    k_lsd_plan<<<n, 1024, 0, s>>>(PB, nb, n <= 2 ? h0->first_wave_latency : h0->first_wave, h0->wave_growth);
    k_lsd_scatter<<<dim3(296, n), 256, 0, s>>>(PB, S, nb);
    count_launches((h0->blur_k ? 2 : 0) + 5);
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}
// this image's descriptor for the batched passes (host side only)
static int lsd_fill_dev(LineImpl* h, GrowDev& D) {
    const int W = h->W, H = h->H;
    D.C.W = W; D.C.H = H; D.C.px = h->px.p + W + 2; D.C.dabc = h->dabc.p; D.C.tab_seed = h->tab_seed.p;
    D.C.pool = h->pool.p; D.C.pool_ctr = h->ctrs.p; D.C.pool_chunks = h->pool_chunks;
    D.C.srec = h->srec0.p; D.C.regang = h->regang.p;
    D.C.dirty[0] = h->dirty.p; D.C.dirty[1] = h->dirty.p + h->dirty_words; D.C.tile_wpr = h->tile_wpr; D.C.stamp = 0; D.dirty_words = h->dirty_words;
    D.C.seed_pix = h->seed_pix.p; D.C.seed_prio = h->seed_prio.p; D.C.prec = h->prec;
    {
        const double margin = 0.1 * M_PI / 180.0;
        D.C.fast_align = (h->prec + margin < 80.0 * M_PI / 180.0) && !getenv("OLF_LSD_EXACT_ALIGN");
        D.C.c_hi2 = (float)(std::cos(h->prec - margin) * std::cos(h->prec - margin));
        D.C.c_lo2 = (float)(std::cos(h->prec + margin) * std::cos(h->prec + margin));
    }
    D.plan = h->plan.p; D.wl0 = h->wl0.p; D.wl1 = h->wl1.p; D.wl2 = h->wl2.p;
    D.F.final_pool = h->final_pool.p; D.F.final_ctr = h->ctrs.p + 2; D.F.regs = h->regs.p; D.F.nreg = h->ctrs.p + 3;
    D.F.reg_cap = h->reg_cap; D.F.min_reg_size = h->min_reg_size;
    D.status = h->status.p; D.max_rounds = h->max_rounds;
    D.defer = getenv("OLF_LSD_NO_DEFER") ? 0 : 1;
    D.test_recheck = getenv("OLF_LSD_TEST_RECHECK") ? 1 : 0;
    D.dbg = h->trace ? h->dbg.p : nullptr;
    D.cont[0] = h->cont.p; D.cont[1] = h->cont.p + h->cont.n / 2; D.budget = h->grow_budget;
    D.sstate = h->sstate.p; D.tbox = h->tbox.p; D.event_scan = getenv("OLF_LSD_FULL_SCAN") ? 0 : 1;
    return OLF_OK;
}
// stage 3 (after the passes): first half of the rectangle fit + the counters the host needs
static void post_fill(PostBatch& PB, LineImpl* const* hs, int n) {
    memset(&PB, 0, sizeof(PB));
    for (int k = 0; k < n; ++k) {
        LineImpl* h = hs[k];
        PB.regs[k] = h->regs.p; PB.nreg[k] = h->ctrs.p + 3; PB.final_pool[k] = h->final_pool.p; PB.dabc[k] = h->dabc.p;
        PB.rect_host[k] = h->rect_host.d; PB.dir_host[k] = h->dir_host.d; PB.seg_host[k] = h->seg_host.d;
        PB.status[k] = h->status.p; PB.status_host[k] = h->status_host.d; PB.nreg_host[k] = h->nreg_host.d;
        PB.lbd_blur[k] = h->lbd_blur.p; PB.grad[k] = h->grad.p; PB.lines[k] = h->lbd_lines.d; PB.rowsum[k] = h->rowsum.p; PB.desc_host[k] = h->desc_host.d;
    }
}
// stage 3 (after the passes): first half of the rectangle fit + the counters the host needs, all images of the batch in one launch
static int lsd_enqueue_rect_a(LineImpl* const* hs, int n, cudaStream_t s) {
    PostBatch PB; post_fill(PB, hs, n);
    k_lsd_rect_a<<<dim3(296, n), 256, 0, s>>>(PB, hs[0]->reg_cap, hs[0]->W, hs[0]->prec);
    count_launches(1);
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}

// LSD of a batch of uploaded images on ONE stream; segments of image k (seed order) -> segs[k]
static inline long long now_us() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int lsd_run_batch(LineImpl* const* hs, int n, cudaStream_t s, std::vector<float4>* segs) {
    const long long t_begin = now_us();
    if (n < 1 || n > LSD_MAX_BATCH) { set_last_error("LSD batch size out of range"); return OLF_ERR_ARG; }
    GrowBatch B; memset(&B, 0, sizeof(B));
    int rc;
    if ((rc = lsd_enqueue_pre_batch(hs, n, s))) return rc;
    for (int k = 0; k < n; ++k) { if ((rc = lsd_fill_dev(hs[k], B.d[k]))) return rc; B.st[k] = hs[k]->phase.p; }
    LineImpl* h0 = hs[0];
    B.conv = h0->conv.p; B.n = n;
    OLF_CUDA(cudaMemsetAsync(h0->conv.p, 0, (LSD_MAX_WAVES + 1) * sizeof(unsigned), s));
    OLF_CUDA(cudaEventRecord(h0->ev_grow0, s));
    const bool pipelined = h0->pipeline < 0 ? n <= 2 : h0->pipeline != 0;           // single frame: latency; batch: throughput (see k_lsd_grow)
    auto enqueue_phases = [&](int count) {
        for (int k = 0; k < count; ++k) {
            const dim3 gs(n <= 2 ? h0->scan_blocks_wide : h0->scan_blocks, n), gv(n <= 2 ? h0->verify_blocks_wide : h0->verify_blocks, n);
            const dim3 gg(n <= 2 ? h0->grow_blocks_wide : h0->grow_blocks_narrow, n);
#if OLF_PDL
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t cfg = {}; cfg.stream = s; cfg.attrs = at; cfg.numAttrs = 1; cfg.dynamicSmemBytes = 0;
            cfg.gridDim = gs; cfg.blockDim = dim3(256); cudaLaunchKernelEx(&cfg, k_lsd_scan, B);
            cfg.gridDim = gv; cfg.blockDim = dim3(128); cudaLaunchKernelEx(&cfg, k_lsd_verify, B);
            cfg.gridDim = gg; cfg.blockDim = dim3(GROW_THREADS);
            if (pipelined) cudaLaunchKernelEx(&cfg, k_lsd_grow<true>, B); else cudaLaunchKernelEx(&cfg, k_lsd_grow<false>, B);
#else
            k_lsd_scan<<<gs, 256, 0, s>>>(B);
            k_lsd_verify<<<gv, 128, 0, s>>>(B);
            if (pipelined) k_lsd_grow<true><<<gg, GROW_THREADS, 0, s>>>(B); else k_lsd_grow<false><<<gg, GROW_THREADS, 0, s>>>(B);
#endif
        }
        count_launches(3 * count);
    };
    // The wave / round state machines run on the device.  Default: the three passes are the body of a CUDA-graph WHILE node
    // that the last finishing image ends (cudaGraphSetConditional): one graph launch per batch, no launch is enqueued that
    // has nothing to do.  OLF_LSD_GRAPH=0: a fixed batch of plain launches that normally covers all rounds of all images
    // (launches after an image is done return at once), otherwise keep going (rare).
    const dim3 g_scan(n <= 2 ? h0->scan_blocks_wide : h0->scan_blocks, n), g_verify(n <= 2 ? h0->verify_blocks_wide : h0->verify_blocks, n),
               g_grow(n <= 2 ? h0->grow_blocks_wide : h0->grow_blocks_narrow, n);
    bool graph_ok = false;
    if (h0->use_graph && !OLF_PDL) {
        unsigned long long key = 1469598103934665603ull;
        { const unsigned char* b = (const unsigned char*)&B; for (size_t i = 0; i < offsetof(GrowBatch, cond); ++i) { key ^= b[i]; key *= 1099511628211ull; } }
        if (!h0->graph_exec || h0->graph_key != key) {
            if (h0->graph_exec) { cudaGraphExecDestroy(h0->graph_exec); h0->graph_exec = nullptr; }
            if (h0->graph) { cudaGraphDestroy(h0->graph); h0->graph = nullptr; }
            cudaGraph_t g = nullptr; cudaGraphConditionalHandle cond = 0;
            bool ok = cudaGraphCreate(&g, 0) == cudaSuccess && cudaGraphConditionalHandleCreate(&cond, g, 1, cudaGraphCondAssignDefault) == cudaSuccess;
            cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
            np.type = cudaGraphNodeTypeConditional; np.conditional.handle = cond; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
            cudaGraphNode_t wnode = nullptr;
            ok = ok && cudaGraphAddNode(&wnode, g, nullptr, 0, &np) == cudaSuccess;
            if (ok) {
                cudaGraph_t body = np.conditional.phGraph_out[0];
                GrowBatch Bg = B; Bg.cond = cond;
                void* args[1] = {&Bg};
                cudaKernelNodeParams kp = {};
                cudaGraphNode_t n1 = nullptr, n2 = nullptr, n3 = nullptr;
                kp.kernelParams = args; kp.sharedMemBytes = 0;
                kp.func = (void*)k_lsd_scan; kp.gridDim = g_scan; kp.blockDim = dim3(256);
                ok = ok && cudaGraphAddKernelNode(&n1, body, nullptr, 0, &kp) == cudaSuccess;
                kp.func = (void*)k_lsd_verify; kp.gridDim = g_verify; kp.blockDim = dim3(128);
                ok = ok && cudaGraphAddKernelNode(&n2, body, &n1, 1, &kp) == cudaSuccess;
                kp.func = pipelined ? (void*)k_lsd_grow<true> : (void*)k_lsd_grow<false>; kp.gridDim = g_grow; kp.blockDim = dim3(GROW_THREADS);
                ok = ok && cudaGraphAddKernelNode(&n3, body, &n2, 1, &kp) == cudaSuccess;
                ok = ok && cudaGraphInstantiate(&h0->graph_exec, g, 0) == cudaSuccess;
            }
            if (!ok) { cudaGetLastError(); if (g) cudaGraphDestroy(g); h0->graph_exec = nullptr; h0->use_graph = 0; }      // driver without conditional nodes: plain launches
            else { h0->graph = g; h0->graph_key = key; }
        }
        graph_ok = h0->graph_exec != nullptr;
    }
    for (int guard = 0; guard < 400; ++guard) {
        if (graph_ok) { OLF_CUDA(cudaGraphLaunch(h0->graph_exec, s)); count_launches(3 * 30); }      // ~30 rounds x 3 passes (the exact count stays on the device)
        else enqueue_phases(h0->phase_batch);
        OLF_CUDA(cudaEventRecord(h0->ev_grow1, s));
        if ((rc = lsd_enqueue_rect_a(hs, n, s))) return rc;
        OLF_CUDA(h0->sync.sync(s));
        bool all = true;
        for (int k = 0; k < n; ++k) all = all && (hs[k]->status_host.p[3] || hs[k]->status_host.p[0]);
        if (all) break;
    }
    const long long t_chain = now_us();
    float grow_ms = 0; cudaEventElapsedTime(&grow_ms, h0->ev_grow0, h0->ev_grow1);
    for (int k = 0; k < n; ++k) {
        LineImpl* h = hs[k];
        h->last_stats[0] = h->status_host.p[1]; h->last_stats[1] = h->status_host.p[2];
        h->last_stats[3] = (int)(grow_ms * 1000.f); h->last_stats[4] = n;
        if (h->status_host.p[0] != 0) { set_last_error("LSD region growing: internal pool overflow"); return OLF_ERR_CAPACITY; }
        if (!h->status_host.p[3] || h->status_host.p[1] + 2 >= (int)h->max_rounds) { set_last_error("LSD region growing did not converge"); return OLF_ERR_INTERNAL; }
        const int nr = (int)*h->nreg_host.p;
        h->last_stats[2] = nr;
        segs[k].clear();
        if (nr == 0) continue;
        // libm cos/sin of the O(#regions) rectangle angles; the projection pass of all images follows in one launch
        for (int i = 0; i < nr; ++i) { const double t = h->rect_host.p[i].theta; h->dir_host.p[i] = make_double2(std::cos(t), std::sin(t)); }
    }
    {
        PostBatch PB; post_fill(PB, hs, n);
        int max_nr = 0;
        for (int k = 0; k < n; ++k) { PB.nr[k] = hs[k]->last_stats[2]; max_nr = std::max(max_nr, PB.nr[k]); }
        if (max_nr > 0) { k_lsd_rect_b<<<dim3((max_nr + 7) / 8, n), 256, 0, s>>>(PB, h0->W, h0->P.lsd_scale); count_launches(1); }
    }
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(h0->sync.sync(s));
    h0->last_stats[5] = (int)(t_chain - t_begin); h0->last_stats[6] = (int)(now_us() - t_chain);      // host view: enqueue + chain wait / trig + rect_b wait (us)
    for (int k = 0; k < n; ++k) {
        LineImpl* h = hs[k];
        const int nr = h->last_stats[2];
        if (nr == 0) continue;
        // seed order = ascending priority key
        std::vector<int> order(nr);
        std::iota(order.begin(), order.end(), 0);
        const RectRec* rr = h->rect_host.p;
        std::sort(order.begin(), order.end(), [rr](int a, int b) { return rr[a].prio < rr[b].prio; });
        segs[k].resize(nr);
        for (int i = 0; i < nr; ++i) segs[k][i] = h->seg_host.p[order[i]];
    }
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
        // the reference's unqualified atan2(float, float) is atan2f: its precompiled header includes <math.h> (LSDDetector_custom.cpp:298)
        kl.angle = atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);
        kl.class_id = ++class_counter;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, hgt);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2; kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        out.push_back(kl);
    }
}

// LBD of `n` keylines on the uploaded image: blur 5x5 sigma 1 + Sobel, row sums, fold + binarise (enqueue only; the
// descriptors land in h->desc_host once the stream has been waited for)
// LBD, image half (blur 5x5 + Sobel; binary_descriptor_custom.cpp:350-412): independent of the keylines, so a batch runs it with the LSD pre-phase,
// off the critical path after the region growing; one launch per stage for all images
static int lbd_prepare_batch(LineImpl* const* hs, int n, cudaStream_t s) {
    LineImpl* h0 = hs[0];
    const int w = h0->img_w, hgt = h0->img_h;
    const LevelTable T = single_level(w, hgt, h0->ipitch);
    BlurBatch bb; memset(&bb, 0, sizeof(bb));
    bool tma = true;
    for (int k = 0; k < n; ++k) { bb.src[k] = hs[k]->img.p; bb.dst[k] = hs[k]->lbd_blur.p; tma = tma && hs[k]->has_tma; }
    if (tma) {
        static thread_local TmaSet<LSD_PRE_MAX> M;
        for (int k = 0; k < n; ++k) M.m[k] = hs[k]->tmaps[1].m[0];
        k_blur_q8_tma<5, LSD_PRE_MAX><<<dim3(T.tile_start[1], n), 256, 0, s>>>(bb, M, T, 14, 62, 104, 0);       // 5x5 sigma 1 (:358)
    } else k_blur_q8<5><<<dim3(T.tile_start[1], n), 256, 0, s>>>(bb, T, 14, 62, 104, 0);
    PostBatch PB; post_fill(PB, hs, n);
    dim3 g((w + 31) / 32, (hgt + 7) / 8, n);
    k_sobel3<<<g, 256, 0, s>>>(PB, w, hgt, h0->ipitch);
    count_launches(2);
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}
// LBD, line half (computeLBD, binary_descriptor_custom.cpp:1026-1372): the keylines of every image of the batch in two launches
static int lbd_describe_batch(LineImpl* const* hs, int n, const std::vector<olf_keyline>* kls, cudaStream_t s) {
    int rc, max_n = 0;
    for (int k = 0; k < n; ++k) {
        LineImpl* h = hs[k];
        const int nl = (int)kls[k].size();
        max_n = std::max(max_n, nl);
        if (nl > h->lbd_cap) {
            const int cap = std::max(2 * nl, 1024);
            if ((rc = h->lbd_lines.ensure(cap)) || (rc = h->rowsum.ensure((size_t)cap * 63)) || (rc = h->desc_host.ensure((size_t)cap * 32))) return rc;
            h->lbd_cap = cap;
        }
        for (int i = 0; i < nl; ++i) {
            const olf_keyline& kk = kls[k][i];
            LbdLine L;
            L.sx = kk.sPointInOctaveX; L.sy = kk.sPointInOctaveY; L.ex = kk.ePointInOctaveX; L.ey = kk.ePointInOctaveY;
            L.dL0 = cosf(kk.angle); L.dL1 = sinf(kk.angle);       // cos( float ), sin( float ) = cosf, sinf in the reference's build (:1130-1131)
            L.num_px = kk.numOfPixels;
            h->lbd_lines.p[i] = L;
        }
    }
    if (max_n == 0) return OLF_OK;
    PostBatch PB; post_fill(PB, hs, n);                          // after the ensure() calls: the buffers may have moved
    for (int k = 0; k < n; ++k) PB.nr[k] = (int)kls[k].size();
    k_lbd_rows<<<dim3((max_n * 63 + 255) / 256, n), 256, 0, s>>>(PB, hs[0]->img_w, hs[0]->img_h);
    k_lbd_fold<<<dim3((max_n + 63) / 64, n), 64, 0, s>>>(PB);
    count_launches(2);
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}
static int lbd_enqueue(LineImpl* h, const olf_keyline* kls, int n, cudaStream_t s) {
    if (n == 0) return OLF_OK;
    int rc;
    if ((rc = lbd_prepare_batch(&h, 1, s))) return rc;
    std::vector<olf_keyline> v(kls, kls + n);
    return lbd_describe_batch(&h, 1, &v, s);
}

int line_lsd_detect(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, float* segs, int cap, int* n) {
    if (!h || !img || !n || w <= 0 || hgt <= 0 || stride < w) { set_last_error("olf_lsd_detect: bad arguments"); return OLF_ERR_ARG; }
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = line_upload(h, img, w, hgt, stride, on_device, h->stream);
    if (rc) return rc;
    std::vector<float4> sg;
    if ((rc = lsd_run_batch(&h, 1, h->stream, &sg))) return rc;
    *n = (int)sg.size();
    if (*n > cap) { set_last_error("olf_lsd_detect: segment capacity too small"); return OLF_ERR_CAPACITY; }
    if (*n) memcpy(segs, sg.data(), sg.size() * sizeof(float4));
    return OLF_OK;
}

int line_lbd_compute(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, uint8_t* desc) {
    if (!h || !img || w <= 0 || hgt <= 0 || stride < w || n < 0) { set_last_error("olf_lbd_compute: bad arguments"); return OLF_ERR_ARG; }
    if (n == 0) return OLF_OK;                           // "keypoint list is empty": descriptors untouched (:556-560)
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = line_upload(h, img, w, hgt, stride, false, h->stream);
    if (rc) return rc;
    if ((rc = lbd_enqueue(h, kls, n, h->stream))) return rc;
    OLF_CUDA(h->sync.sync(h->stream));
    memcpy(desc, h->desc_host.p, (size_t)n * 32);
    return OLF_OK;
}

// Lineextractor::operator() (src/LineExtractor.cc:31-67) for a batch of images (one per handle) on the first handle's stream
int line_extract_batch(LineImpl* const* hs, int nimg, const uint8_t* const* imgs, int w, int hgt, int stride, bool on_device,
                       olf_keyline* const* kls, uint8_t* const* desc, int cap, int* n) {
    if (!hs || nimg < 1 || nimg > LSD_MAX_BATCH || !imgs || !n || w <= 0 || hgt <= 0 || stride < w) { set_last_error("olf_line_extract: bad arguments"); return OLF_ERR_ARG; }
    for (int k = 0; k < nimg; ++k) { if (!hs[k] || !imgs[k]) { set_last_error("olf_line_extract: bad arguments"); return OLF_ERR_ARG; } n[k] = 0; }
    OLF_CUDA(cudaSetDevice(hs[0]->device));
    cudaStream_t s = hs[0]->stream;
    int rc;
    for (int k = 0; k < nimg; ++k) if ((rc = line_upload(hs[k], imgs[k], w, hgt, stride, on_device, s))) return rc;
    std::vector<float4> sg[LSD_MAX_BATCH];
    if ((rc = lbd_prepare_batch(hs, nimg, s))) return rc;           // ahead of the LSD chain: needs the uploaded images only
    if ((rc = lsd_run_batch(hs, nimg, s, sg))) return rc;
    const long long t_lsd = now_us();
    std::vector<olf_keyline> kl[LSD_MAX_BATCH];
    for (int k = 0; k < nimg; ++k) {
        LineImpl* h = hs[k];
        std::vector<olf_keyline>& v = kl[k];
        make_keylines(sg[k], w, hgt, h->P.min_line_length * std::min(w, hgt), v);
        const int nf = h->P.lsd_nfeatures;
        if ((int)v.size() > nf && nf != 0) {
            // canonical: stable sort by response
            std::stable_sort(v.begin(), v.end(), [](const olf_keyline& a, const olf_keyline& b) { return a.response > b.response; });
            v.resize(nf);
            for (int i = 0; i < nf; i++) v[i].class_id = i;
        }
        if ((int)v.size() > cap) { set_last_error("olf_line_extract: keyline capacity too small"); return OLF_ERR_CAPACITY; }
        n[k] = (int)v.size();
        if (v.empty()) continue;
        memcpy(kls[k], v.data(), v.size() * sizeof(olf_keyline));
    }
    if ((rc = lbd_describe_batch(hs, nimg, kl, s))) return rc;
    OLF_CUDA(hs[0]->sync.sync(s));
    hs[0]->last_stats[7] = (int)(now_us() - t_lsd);                      // keylines + LBD (us)
    for (int k = 0; k < nimg; ++k) if (n[k]) memcpy(desc[k], hs[k]->desc_host.p, (size_t)n[k] * 32);
    return OLF_OK;
}
int line_extract(LineImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keyline* kls, uint8_t* desc, int cap, int* n) {
    if (!h || !n) { set_last_error("olf_line_extract: bad arguments"); return OLF_ERR_ARG; }
    return line_extract_batch(&h, 1, &img, w, hgt, stride, on_device, &kls, &desc, cap, n);
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
