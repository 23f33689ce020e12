// Matching half of the front end on sm_100a.
//   olf_knn2_hamming / olf_match_nnr / olf_match_lines : cv::BFMatcher(NORM_HAMMING).knnMatch + matchNNR + match
//                                                        (reference src/LineMatcher.cpp:42-132)
//   olf_stereo_points : Frame::ComputeStereoMatches     (src/Frame.cc:702-876)
//   olf_stereo_lines  : Frame::ComputeStereoMatches_Lines + matchGrid(lines) + GridStructure + LineIterator
//                       (src/Frame.cc:878-1048, src/LineMatcher.cpp:220-299, src/gridStructure.cpp, src/LineIterator.cpp)
//   olf_search_by_projection_last / _map : ORBmatcher::SearchByProjection (src/ORBmatcher.cc:47-139, 1474-1618, 1749-1790)
//                       with Frame::PosInGrid / GetFeaturesInArea (src/Frame.cc:517-582)
// All Hamming distances are warp-parallel POPC work; the order-dependent parts of the reference (running
// `distances[i2]` cross-check, "already holds a MapPoint" blocking) are reproduced exactly: either by walking the
// dependent index sequentially inside one thread per independent chain, or by iterating the assignment operator to
// its (unique, well-founded) fixed point.
#include "common.cuh"
#include "orb.h"
#include "match.h"
#include <algorithm>
#include <climits>
#include <cmath>

namespace olf {

// ---- per-thread scratch: one stream + growing device / pinned arenas per device -------------------------------
struct Arena {
    DevBuf<uint8_t> dev; PinBuf<uint8_t> pin;
    size_t dev_off = 0, pin_off = 0;
    void reset() { dev_off = pin_off = 0; }
};
struct MatchCtx {
    int device = -1;
    cudaStream_t stream = nullptr;     // the calling thread's own stream
    cudaStream_t cur = nullptr;        // stream of this call: the thread's own, or the one a rig lent (match_use_stream)
    Arena a;
    // a thread that ends gives its stream and arenas back (callers may be short-lived threads)
    ~MatchCtx() {
        if (device < 0) return;
        if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        a.dev.release(); a.pin.release();
        cudaGetLastError();
    }
};
static thread_local MatchCtx g_ctx[16];
static thread_local cudaStream_t g_lent_stream = nullptr;
static thread_local MatchCtx* g_lent_ctx = nullptr;
// A rig owns its matcher scratch (arenas) instead of taking the calling thread's: callers may be short-lived threads, and
// creating a context costs device allocations, which synchronise the whole device.
MatchCtx* match_ctx_create(int device) { MatchCtx* c = new MatchCtx(); c->device = device; return c; }
void match_ctx_destroy(MatchCtx* c) { delete c; }
void match_use_ctx(MatchCtx* c) { g_lent_ctx = c; }
// A rig runs its stereo matchers on one of its own streams instead of one more stream per calling thread: every stream
// beyond the 32 hardware work queues aliases another one and queues behind its (long) chains.
void match_use_stream(cudaStream_t s) { g_lent_stream = s; }

static int get_ctx(int device, MatchCtx** out) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || device >= 16) {
        cudaGetLastError();
        set_last_error("no such CUDA device (this library has no CPU path)");
        return OLF_ERR_NO_DEVICE;
    }
    OLF_CUDA(cudaSetDevice(device));
    MatchCtx& c = (g_lent_ctx && g_lent_ctx->device == device) ? *g_lent_ctx : g_ctx[device];
    if (g_lent_stream) c.cur = g_lent_stream;
    else {
        if (!c.stream) OLF_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.cur = c.stream;
    }
    c.device = device;
    c.a.reset();
    *out = &c;
    return OLF_OK;
}
// Two-pass use: first call plan(bytes) for everything, then ensure(), then take().
struct Planner {
    size_t dev = 0, pin = 0;
    size_t d(size_t bytes) { size_t o = dev; dev += align_up_sz(bytes, 256); return o; }
    size_t p(size_t bytes) { size_t o = pin; pin += align_up_sz(bytes, 256); return o; }
};
static int arena_ensure(MatchCtx* c, const Planner& pl) {
    int rc;
    // generous first allocation: the arena sizes follow the per-frame feature counts, which wobble from frame to frame
    if (c->a.dev.n < pl.dev + 256 && (rc = c->a.dev.ensure(std::max<size_t>(2 * (pl.dev + 256), (size_t)4 << 20)))) return rc;
    if (c->a.pin.n < pl.pin + 256 && (rc = c->a.pin.ensure(std::max<size_t>(2 * (pl.pin + 256), (size_t)1 << 20)))) return rc;
    return OLF_OK;
}
template <typename T> static T* dptr(MatchCtx* c, size_t off) { return (T*)(c->a.dev.p + off); }
template <typename T> static T* hptr(MatchCtx* c, size_t off) { return (T*)(c->a.pin.p + off); }
template <typename T> static T* hdptr(MatchCtx* c, size_t off) { return (T*)(c->a.pin.d + off); }     // device alias of pinned

// ======================================================================================================
// brute-force kNN(2), ties -> lowest train index
// ======================================================================================================
#define KNN_Q 128      // queries per block (one per thread)
#define KNN_T 128      // train descriptors per smem tile
struct Knn2 { int d0, i0, d1, i1; };
__device__ __forceinline__ void knn_update(Knn2& k, int d, int j) {
    if (d < k.d0) { k.d1 = k.d0; k.i1 = k.i0; k.d0 = d; k.i0 = j; }
    else if (d < k.d1) { k.d1 = d; k.i1 = j; }
}
__global__ void __launch_bounds__(KNN_Q) k_knn2_partial(const uint4* __restrict__ q, int n1, const uint4* __restrict__ t, int n2,
                                                        int per_split, Knn2* __restrict__ part) {
    __shared__ uint4 tile[KNN_T * 2];
    const int qi = blockIdx.x * KNN_Q + threadIdx.x;
    const int t0 = blockIdx.y * per_split, t1 = min(t0 + per_split, n2);
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (qi < n1) { a = q[2 * qi]; b = q[2 * qi + 1]; }
    Knn2 k; k.d0 = INT_MAX; k.d1 = INT_MAX; k.i0 = -1; k.i1 = -1;
    for (int base = t0; base < t1; base += KNN_T) {
        const int cnt = min(KNN_T, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += KNN_Q) tile[i] = t[2 * (size_t)base + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const uint4 c = tile[2 * j], e = tile[2 * j + 1];
            const int d = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) +
                          __popc(b.x ^ e.x) + __popc(b.y ^ e.y) + __popc(b.z ^ e.z) + __popc(b.w ^ e.w);
            knn_update(k, d, base + j);
        }
    }
    if (qi < n1) part[(size_t)blockIdx.y * n1 + qi] = k;
}
__global__ void k_knn2_merge(const Knn2* __restrict__ part, int n1, int splits, int* __restrict__ i0, int* __restrict__ d0,
                             int* __restrict__ i1, int* __restrict__ d1) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n1) return;
    Knn2 k = part[qi];
    for (int s = 1; s < splits; ++s) {
        const Knn2 p = part[(size_t)s * n1 + qi];
        if (p.i0 >= 0) knn_update(k, p.d0, p.i0);
        if (p.i1 >= 0) knn_update(k, p.d1, p.i1);
    }
    i0[qi] = k.i0; d0[qi] = k.d0; i1[qi] = k.i1; d1[qi] = k.d1;
}
// matchNNR acceptance (src/LineMatcher.cpp:54-59) + optional mutual check (src/LineMatcher.cpp:120-127)
__global__ void k_nnr_accept(const int* __restrict__ i0, const int* __restrict__ d0, const int* __restrict__ d1, int n1, int n2, float nnr, int* __restrict__ m12) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    int m = -1;
    if (n2 >= 2 && (float)d0[i] < fmul((float)d1[i], nnr)) m = i0[i];
    m12[i] = m;
}
__global__ void k_mutual(int* __restrict__ m12, const int* __restrict__ m21, int n1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int j = m12[i];
    if (j >= 0 && m21[j] != i) m12[i] = -1;
}

static void launch_knn2(cudaStream_t s, const uint4* dq, int n1, const uint4* dt, int n2, Knn2* part, int* i0, int* d0, int* i1, int* d1, int splits, int per_split) {
    dim3 g((n1 + KNN_Q - 1) / KNN_Q, splits);
    k_knn2_partial<<<g, KNN_Q, 0, s>>>(dq, n1, dt, n2, per_split, part);
    k_knn2_merge<<<(n1 + 127) / 128, 128, 0, s>>>(part, n1, splits, i0, d0, i1, d1);
    count_launches(2);
}
static void knn_splits(int n1, int n2, int* splits, int* per_split) {
    const int qb = std::max((n1 + KNN_Q - 1) / KNN_Q, 1);
    int sp = std::max(1, std::min((592 + qb - 1) / qb, (n2 + 255) / 256));
    int per = (std::max(n2, 1) + sp - 1) / sp;
    per = (per + KNN_T - 1) / KNN_T * KNN_T;
    sp = (std::max(n2, 1) + per - 1) / per;
    *splits = sp; *per_split = per;
}

int knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx0, int* dist0, int* idx1, int* dist1, int device) {
    if (n1 < 0 || n2 < 0 || (n1 && !d1) || (n2 && !d2)) { set_last_error("olf_knn2_hamming: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    if (n1 == 0) return OLF_OK;
    int splits, per; knn_splits(n1, n2, &splits, &per);
    Planner pl;
    const size_t o_q = pl.d((size_t)n1 * 32), o_t = pl.d((size_t)std::max(n2, 1) * 32), o_part = pl.d((size_t)splits * n1 * sizeof(Knn2)), o_out = pl.d((size_t)4 * n1 * 4);
    const size_t p_q = pl.p((size_t)n1 * 32), p_t = pl.p((size_t)std::max(n2, 1) * 32), p_out = pl.p((size_t)4 * n1 * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    memcpy(hptr<uint8_t>(c, p_q), d1, (size_t)n1 * 32);
    if (n2) memcpy(hptr<uint8_t>(c, p_t), d2, (size_t)n2 * 32);
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_q), hptr<uint8_t>(c, p_q), (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    if (n2) OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_t), hptr<uint8_t>(c, p_t), (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    int* o = dptr<int>(c, o_out);
    launch_knn2(s, dptr<uint4>(c, o_q), n1, dptr<uint4>(c, o_t), n2, dptr<Knn2>(c, o_part), o, o + n1, o + 2 * n1, o + 3 * n1, splits, per);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_out), o, (size_t)4 * n1 * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    const int* ho = hptr<int>(c, p_out);
    memcpy(idx0, ho, n1 * 4); memcpy(dist0, ho + n1, n1 * 4); memcpy(idx1, ho + 2 * n1, n1 * 4); memcpy(dist1, ho + 3 * n1, n1 * 4);
    return OLF_OK;
}

// ---- C5 micro-benchmark support (BASELINE.json configs[4]) -----------------------------------------------------------
// Integer-pipe ceiling for the all-pairs Hamming kernel: independent xor + popc + add chains, the instruction mix of one
// 32-bit word of a descriptor pair.  popc_ops = word-pairs per second the chip sustains when nothing else is in the way.
__global__ void __launch_bounds__(256) k_popc_peak(unsigned* __restrict__ out, int iters) {
    unsigned a0 = threadIdx.x * 2654435761u + blockIdx.x, a1 = a0 ^ 0x9e3779b9u, a2 = a0 + 0x7f4a7c15u, a3 = ~a0;
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < iters; ++i) {
        const unsigned k = (unsigned)i * 0x01000193u;
        s0 += __popc(a0 ^ k); s1 += __popc(a1 ^ k); s2 += __popc(a2 ^ k); s3 += __popc(a3 ^ k);
        s4 += __popc(a0 ^ ~k); s5 += __popc(a1 ^ ~k); s6 += __popc(a2 ^ ~k); s7 += __popc(a3 ^ ~k);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}
// kernel-only time of the knn2 pair (partial + merge) on device-resident random descriptors, and the measured popc ceiling
int knn2_bench(int n1, int n2, int iters, int device, double* kernel_ms, double* popc_word_pairs_per_s) {
    if (n1 < 1 || n2 < 2 || iters < 1 || !kernel_ms || !popc_word_pairs_per_s) return OLF_ERR_ARG;
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    int splits, per; knn_splits(n1, n2, &splits, &per);
    Planner pl;
    const size_t o_q = pl.d((size_t)n1 * 32), o_t = pl.d((size_t)n2 * 32), o_part = pl.d((size_t)splits * n1 * sizeof(Knn2)), o_out = pl.d((size_t)4 * n1 * 4);
    const size_t o_pk = pl.d((size_t)148 * 8 * 256 * 4);
    const size_t p_q = pl.p((size_t)std::max(n1, n2) * 32);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    uint32_t x = 12345u;
    uint32_t* h = hptr<uint32_t>(c, p_q);
    for (size_t i = 0; i < (size_t)std::max(n1, n2) * 8; ++i) { x = x * 1664525u + 1013904223u; h[i] = x ^ (x >> 13); }
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_q), h, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_t), h, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    cudaEvent_t e0, e1;
    OLF_CUDA(cudaEventCreate(&e0)); OLF_CUDA(cudaEventCreate(&e1));
    int* o = dptr<int>(c, o_out);
    for (int w = 0; w < 3; ++w) launch_knn2(s, dptr<uint4>(c, o_q), n1, dptr<uint4>(c, o_t), n2, dptr<Knn2>(c, o_part), o, o + n1, o + 2 * n1, o + 3 * n1, splits, per);
    OLF_CUDA(cudaEventRecord(e0, s));
    for (int w = 0; w < iters; ++w) launch_knn2(s, dptr<uint4>(c, o_q), n1, dptr<uint4>(c, o_t), n2, dptr<Knn2>(c, o_part), o, o + n1, o + 2 * n1, o + 3 * n1, splits, per);
    OLF_CUDA(cudaEventRecord(e1, s));
    OLF_CUDA(cudaEventSynchronize(e1));
    float ms = 0; OLF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *kernel_ms = ms / iters;
    const int pk_iters = 20000, pk_blocks = 148 * 8;
    k_popc_peak<<<pk_blocks, 256, 0, s>>>(dptr<unsigned>(c, o_pk), 100);
    OLF_CUDA(cudaEventRecord(e0, s));
    k_popc_peak<<<pk_blocks, 256, 0, s>>>(dptr<unsigned>(c, o_pk), pk_iters);
    OLF_CUDA(cudaEventRecord(e1, s));
    OLF_CUDA(cudaEventSynchronize(e1));
    OLF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *popc_word_pairs_per_s = (double)pk_blocks * 256 * 8.0 * pk_iters / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}

int match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int mutual, int* m12, int* nmatches, int device) {
    if (n1 < 0 || n2 < 0 || (n1 && !d1) || (n2 && !d2) || !nmatches) { set_last_error("olf_match: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    *nmatches = 0;
    if (n1 == 0) return OLF_OK;
    if (n2 == 0) { for (int i = 0; i < n1; ++i) m12[i] = -1; return OLF_OK; }
    int s12, p12, s21, p21;
    knn_splits(n1, n2, &s12, &p12); knn_splits(n2, n1, &s21, &p21);
    Planner pl;
    const size_t o_q = pl.d((size_t)n1 * 32), o_t = pl.d((size_t)n2 * 32);
    const size_t o_part = pl.d(std::max((size_t)s12 * n1, (size_t)s21 * n2) * sizeof(Knn2));
    const size_t o_a = pl.d((size_t)4 * n1 * 4), o_b = pl.d((size_t)4 * n2 * 4), o_m12 = pl.d((size_t)n1 * 4), o_m21 = pl.d((size_t)n2 * 4);
    const size_t p_q = pl.p((size_t)n1 * 32), p_t = pl.p((size_t)n2 * 32), p_m = pl.p((size_t)n1 * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    memcpy(hptr<uint8_t>(c, p_q), d1, (size_t)n1 * 32); memcpy(hptr<uint8_t>(c, p_t), d2, (size_t)n2 * 32);
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_q), hptr<uint8_t>(c, p_q), (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_t), hptr<uint8_t>(c, p_t), (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    int* a = dptr<int>(c, o_a); int* b = dptr<int>(c, o_b);
    launch_knn2(s, dptr<uint4>(c, o_q), n1, dptr<uint4>(c, o_t), n2, dptr<Knn2>(c, o_part), a, a + n1, a + 2 * n1, a + 3 * n1, s12, p12);
    k_nnr_accept<<<(n1 + 127) / 128, 128, 0, s>>>(a, a + n1, a + 3 * n1, n1, n2, nnr, dptr<int>(c, o_m12));
    count_launches(mutual ? 3 : 1);
    if (mutual) {
        launch_knn2(s, dptr<uint4>(c, o_t), n2, dptr<uint4>(c, o_q), n1, dptr<Knn2>(c, o_part), b, b + n2, b + 2 * n2, b + 3 * n2, s21, p21);
        k_nnr_accept<<<(n2 + 127) / 128, 128, 0, s>>>(b, b + n2, b + 3 * n2, n2, n1, nnr, dptr<int>(c, o_m21));
        k_mutual<<<(n1 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o_m12), dptr<int>(c, o_m21), n1);
    }
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_m), dptr<int>(c, o_m12), (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    const int* hm = hptr<int>(c, p_m);
    int cnt = 0;
    for (int i = 0; i < n1; ++i) { m12[i] = hm[i]; cnt += hm[i] >= 0; }
    *nmatches = cnt;
    return OLF_OK;
}

// ======================================================================================================
// Frame::ComputeStereoMatches (src/Frame.cc:702-876): warp per left keypoint
// ======================================================================================================
struct PyrView { const uint8_t* pyr; int w[OLF_MAX_LEVELS], h[OLF_MAX_LEVELS], pitch[OLF_MAX_LEVELS]; unsigned off[OLF_MAX_LEVELS]; float scale[OLF_MAX_LEVELS], inv_scale[OLF_MAX_LEVELS]; };
__device__ __forceinline__ int px_reflect(const PyrView& P, int l, int x, int y) {
    return P.pyr[P.off[l] + (size_t)reflect101(y, P.h[l]) * P.pitch[l] + reflect101(x, P.w[l])];
}
// Np / Nrp: when not null the keypoint counts are read from the device (the extractor's result is still in flight when
// this kernel is enqueued); N / Nr are then the capacities.
__global__ void __launch_bounds__(256) k_stereo_points(const olf_keypoint* __restrict__ kl, const uint32_t* __restrict__ dl, int N, const int* __restrict__ Np,
                                                       const olf_keypoint* __restrict__ kr, const uint32_t* __restrict__ dr, int Nr, const int* __restrict__ Nrp,
                                                       const __grid_constant__ PyrView PL, const __grid_constant__ PyrView PR,
                                                       float mbf, float fx, float* __restrict__ uRight, float* __restrict__ depth, int* __restrict__ sad) {
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (Np) N = min(N, *Np);
    if (Nrp) Nr = min(Nr, *Nrp);
    if (iL >= N) return;
    const olf_keypoint kpL = kl[iL];
    const int levelL = kpL.octave;
    const float uL = kpL.x, vL = kpL.y;
    const int rowi = (int)vL;                                   // vRowIndices[vL]: float -> size_t truncation
    const float mb = fdiv(mbf, fx);
    const float maxD = fdiv(mbf, mb);
    const float minU = fsub(uL, maxD), maxU = uL;               // minD = 0
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = dl[(size_t)iL * 8 + k];
    // best right keypoint on this row: lexicographic min of (dist, iR), dist < TH_HIGH
    unsigned best = ((unsigned)OLF_TH_HIGH << 20);              // key = dist << 20 | iR  (iR < 2^20)
    if (!(maxU < 0) && rowi >= 0 && rowi < PL.h[0]) {
        for (int iR = lane; iR < Nr; iR += 32) {
            const olf_keypoint k = kr[iR];
            const float r = fmul(2.0f, PL.scale[k.octave]);
            const int maxr = (int)ceilf(fadd(k.y, r)), minr = (int)floorf(fsub(k.y, r));
            if (rowi < minr || rowi > maxr) continue;
            if (k.octave < levelL - 1 || k.octave > levelL + 1) continue;
            if (!(k.x >= minU && k.x <= maxU)) continue;
            const int d = hamming256(a, dr + (size_t)iR * 8);
            const unsigned key = ((unsigned)d << 20) | (unsigned)iR;
            best = min(best, key);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    const int bestDist = (int)(best >> 20);
    float out_u = -1.f, out_d = -1.f; int out_s = -1;
    if (bestDist < (OLF_TH_HIGH + OLF_TH_LOW) / 2) {
        const int bestIdxR = (int)(best & 0xFFFFF);
        const float uR0 = kr[bestIdxR].x;
        const float sf = PL.inv_scale[levelL];
        const float scaleduL = roundf(fmul(kpL.x, sf)), scaledvL = roundf(fmul(kpL.y, sf)), scaleduR0 = roundf(fmul(uR0, sf));
        const int cxL = (int)scaleduL, cyL = (int)scaledvL, cxR = (int)scaleduR0;
        const float iniu = fsub(fadd(scaleduR0, 5.f), 5.f);
        const float endu = fadd(fadd(fadd(scaleduR0, 5.f), 5.f), 1.f);
        if (!(iniu < 0 || endu >= (float)PR.w[levelL])) {
            // 11 SAD windows of 11x11 (minus centre), integer sums (exact in any order)
            int acc[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) acc[k] = 0;
            const int cL = px_reflect(PL, levelL, cxL, cyL);
            int cR[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) cR[k] = px_reflect(PR, levelL, cxR + k - 5, cyL);
            for (int p = lane; p < 121; p += 32) {
                const int dy = p / 11 - 5, dx = p % 11 - 5;
                const int va = px_reflect(PL, levelL, cxL + dx, cyL + dy) - cL;
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    const int vb = px_reflect(PR, levelL, cxR + k - 5 + dx, cyL + dy) - cR[k];
                    acc[k] += abs(va - vb);
                }
            }
#pragma unroll
            for (int k = 0; k < 11; ++k)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            int bestS = INT_MAX, bestinc = 0;
#pragma unroll
            for (int k = 0; k < 11; ++k) if (acc[k] < bestS) { bestS = acc[k]; bestinc = k - 5; }
            if (!(bestinc == -5 || bestinc == 5)) {
                float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
                for (int k = 1; k < 10; ++k) if (k - 5 == bestinc) { d1 = (float)acc[k - 1]; d2 = (float)acc[k]; d3 = (float)acc[k + 1]; }
                const float deltaR = fdiv(fsub(d1, d3), fmul(2.0f, fsub(fadd(d1, d3), fmul(2.0f, d2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = fmul(PL.scale[levelL], fadd(fadd(scaleduR0, (float)bestinc), deltaR));
                    float disparity = fsub(uL, bestuR);
                    if (disparity >= 0 && disparity < maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        out_d = fdiv(mbf, disparity);
                        out_u = bestuR;
                        out_s = bestS;
                    }
                }
            }
        }
    }
    if (lane == 0) { uRight[iL] = out_u; depth[iL] = out_d; sad[iL] = out_s; }
}
// median SAD outlier rejection (src/Frame.cc:861-875): one block; the size/2-th order statistic of the SAD values
// (all < 2^16: 121 px x 510) by a two-level radix select on shared-memory histograms
__global__ void __launch_bounds__(1024) k_stereo_median(float* __restrict__ uRight, float* __restrict__ depth, const int* __restrict__ sad, int N, const int* __restrict__ Np) {
    __shared__ unsigned hist[256];
    if (Np) N = min(N, *Np);
    __shared__ int s_cnt, s_hi, s_rank, s_median;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_cnt = 0; s_median = -1; }
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < N; i += 1024) { const int v = sad[i]; if (v >= 0) { ++local; atomicAdd(&hist[min(v, 65535) >> 8], 1u); } }
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) return;
    if (threadIdx.x == 0) {
        int k = cnt / 2, b = 0;
        while (k >= (int)hist[b]) { k -= (int)hist[b]; ++b; }
        s_hi = b; s_rank = k;
    }
    __syncthreads();
    const int hi = s_hi;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += 1024) { const int v = sad[i]; if (v >= 0 && (min(v, 65535) >> 8) == hi) atomicAdd(&hist[min(v, 65535) & 255], 1u); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = s_rank, b = 0;
        while (k >= (int)hist[b]) { k -= (int)hist[b]; ++b; }
        s_median = (hi << 8) | b;
    }
    __syncthreads();
    const float thDist = fmul(0x1.0cccccp+1f /* 1.5f*1.4f */, (float)s_median);
    for (int i = threadIdx.x; i < N; i += 1024) {
        const int v = sad[i];
        if (v >= 0 && !((float)v < thDist)) { uRight[i] = -1.f; depth[i] = -1.f; }
    }
}

static PyrView make_view(const OrbDeviceView& v) {
    PyrView p; memset(&p, 0, sizeof(p));
    p.pyr = v.pyr;
    for (int l = 0; l < v.nlevels; ++l) { p.w[l] = v.w[l]; p.h[l] = v.h[l]; p.pitch[l] = v.pitch[l]; p.off[l] = v.off[l]; p.scale[l] = v.scale[l]; p.inv_scale[l] = v.inv_scale[l]; }
    return p;
}

int stereo_points(OrbImpl* left, OrbImpl* right, const olf_keypoint* kl, const uint8_t* dl, int N, const olf_keypoint* kr, const uint8_t* dr, int Nr,
                  float bf, float fx, float* uRight, float* depth) {
    if (!left || !right || N < 0 || Nr < 0 || (N && (!kl || !dl || !uRight || !depth)) || (Nr && (!kr || !dr))) { set_last_error("olf_stereo_points: bad arguments"); return OLF_ERR_ARG; }
    const OrbDeviceView vl = orb_device_view(left), vr = orb_device_view(right);
    if (vl.device != vr.device || !vl.pyr || !vr.pyr || vl.nlevels != vr.nlevels) { set_last_error("olf_stereo_points: extractors must have run on the same device"); return OLF_ERR_ARG; }
    if (Nr >= (1 << 20)) { set_last_error("olf_stereo_points: too many right keypoints"); return OLF_ERR_CAPACITY; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(vl.device, &c))) return rc;
    if (N == 0) return OLF_OK;
    Planner pl;
    const size_t o_kl = pl.d((size_t)N * sizeof(olf_keypoint)), o_dl = pl.d((size_t)N * 32), o_kr = pl.d((size_t)std::max(Nr, 1) * sizeof(olf_keypoint)), o_dr = pl.d((size_t)std::max(Nr, 1) * 32);
    const size_t o_u = pl.d((size_t)N * 4), o_d = pl.d((size_t)N * 4), o_s = pl.d((size_t)N * 4);
    const size_t p_in = pl.p((size_t)(N + Nr) * (sizeof(olf_keypoint) + 32)), p_out = pl.p((size_t)2 * N * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    uint8_t* hp = hptr<uint8_t>(c, p_in);
    uint8_t* h_kl = hp; uint8_t* h_dl = h_kl + (size_t)N * sizeof(olf_keypoint); uint8_t* h_kr = h_dl + (size_t)N * 32; uint8_t* h_dr = h_kr + (size_t)Nr * sizeof(olf_keypoint);
    memcpy(h_kl, kl, (size_t)N * sizeof(olf_keypoint)); memcpy(h_dl, dl, (size_t)N * 32);
    if (Nr) { memcpy(h_kr, kr, (size_t)Nr * sizeof(olf_keypoint)); memcpy(h_dr, dr, (size_t)Nr * 32); }
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_kl), h_kl, (size_t)N * sizeof(olf_keypoint), cudaMemcpyHostToDevice, s));
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_dl), h_dl, (size_t)N * 32, cudaMemcpyHostToDevice, s));
    if (Nr) {
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_kr), h_kr, (size_t)Nr * sizeof(olf_keypoint), cudaMemcpyHostToDevice, s));
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_dr), h_dr, (size_t)Nr * 32, cudaMemcpyHostToDevice, s));
    }
    // the extractors' entry points synchronise their own streams before returning, so the pyramids are complete
    k_stereo_points<<<(N + 7) / 8, 256, 0, s>>>(dptr<olf_keypoint>(c, o_kl), dptr<uint32_t>(c, o_dl), N, nullptr, dptr<olf_keypoint>(c, o_kr), dptr<uint32_t>(c, o_dr), Nr, nullptr,
                                                 make_view(vl), make_view(vr), bf, fx, dptr<float>(c, o_u), dptr<float>(c, o_d), dptr<int>(c, o_s));
    k_stereo_median<<<1, 1024, 0, s>>>(dptr<float>(c, o_u), dptr<float>(c, o_d), dptr<int>(c, o_s), N, nullptr);
    count_launches(2);
    float* ho = hptr<float>(c, p_out);
    OLF_CUDA(cudaMemcpyAsync(ho, dptr<float>(c, o_u), (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaMemcpyAsync(ho + N, dptr<float>(c, o_d), (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(uRight, ho, (size_t)N * 4); memcpy(depth, ho + N, (size_t)N * 4);
    return OLF_OK;
}

// Frame::ComputeStereoMatches on the extractors' DEVICE-RESIDENT results, enqueued behind them on stream s (nothing waits):
// uRight / depth of up to `cap` left keypoints land in ws->out (pinned) as [uRight[cap] | depth[cap]].
int stereo_ws_ensure(StereoWs* ws, int cap) {
    int rc;
    if (cap <= ws->cap) return OLF_OK;
    if ((rc = ws->u.ensure(cap)) || (rc = ws->d.ensure(cap)) || (rc = ws->sad.ensure(cap)) || (rc = ws->out.ensure((size_t)2 * cap))) return rc;
    ws->cap = cap;
    return OLF_OK;
}
void stereo_ws_release(StereoWs* ws) { ws->u.release(); ws->d.release(); ws->sad.release(); ws->out.release(); ws->cap = 0; }
int stereo_points_enqueue(StereoWs* ws, OrbImpl* left, OrbImpl* right, float bf, float fx, int cap, cudaStream_t s) {
    const OrbDeviceView vl = orb_device_view(left), vr = orb_device_view(right);
    if (vl.device != vr.device || !vl.pyr || !vr.pyr || vl.nlevels != vr.nlevels) { set_last_error("olf_stereo_points: extractors must have run on the same device"); return OLF_ERR_ARG; }
    if (vr.cap >= (1 << 20)) { set_last_error("olf_stereo_points: too many right keypoints"); return OLF_ERR_CAPACITY; }
    int rc;
    const int N = std::min(cap, vl.cap);
    if (N <= 0) return OLF_OK;
    if ((rc = stereo_ws_ensure(ws, N))) return rc;
    k_stereo_points<<<(N + 7) / 8, 256, 0, s>>>(vl.kps, (const uint32_t*)vl.desc, N, vl.n, vr.kps, (const uint32_t*)vr.desc, vr.cap, vr.n,
                                                 make_view(vl), make_view(vr), bf, fx, ws->u.p, ws->d.p, ws->sad.p);
    k_stereo_median<<<1, 1024, 0, s>>>(ws->u.p, ws->d.p, ws->sad.p, N, vl.n);
    count_launches(2);
    OLF_CUDA(cudaMemcpyAsync(ws->out.p, ws->u.p, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaMemcpyAsync(ws->out.p + ws->cap, ws->d.p, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}

// ======================================================================================================
// Frame::ComputeStereoMatches_Lines (src/Frame.cc:878-1000) + matchGrid(lines) (src/LineMatcher.cpp:220-299)
// ======================================================================================================
#define LCELLS 128
// thread per right line: normalised direction + Bresenham raster (src/LineIterator.cpp:34-77) into grid cells
__global__ void k_lines_raster(const olf_keyline* __restrict__ kr, int n2, double inv_w, double inv_h,
                               double2* __restrict__ dir2, short2* __restrict__ cells, int* __restrict__ ncells) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const olf_keyline k = kr[i];
    const double vx = __dmul_rn((double)fsub(k.endPointX, k.startPointX), inv_w), vy = __dmul_rn((double)fsub(k.endPointY, k.startPointY), inv_h);
    const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    dir2[i] = make_double2(__ddiv_rn(vx, mag), __ddiv_rn(vy, mag));
    double x1 = __dmul_rn((double)k.startPointX, inv_w), y1 = __dmul_rn((double)k.startPointY, inv_h);
    double x2 = __dmul_rn((double)k.endPointX, inv_w), y2 = __dmul_rn((double)k.endPointY, inv_h);
    const bool steep = fabs(y2 - y1) > fabs(x2 - x1);
    if (steep) { double t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
    if (x1 > x2) { double t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const double dx = __dsub_rn(x2, x1), dy = fabs(__dsub_rn(y2, y1));
    double error = __ddiv_rn(dx, 2.0);
    const int ystep = (y1 < y2) ? 1 : -1;
    int x = (int)x1, y = (int)y1;
    const int maxX = (int)x2;
    int n = 0;
    while (x <= maxX) {
        const int px = steep ? y : x, py = steep ? x : y;
        if (px >= 0 && px < OLF_GRID_COLS && py >= 0 && py < OLF_GRID_ROWS && n < LCELLS) cells[(size_t)i * LCELLS + n++] = make_short2((short)px, (short)py);
        error = __dsub_rn(error, dy);
        if (error < 0) { y += ystep; error = __dadd_rn(error, dx); }
        x++;
    }
    ncells[i] = n;
}
// thread per (i1, i2): candidate test (grid windows around both end points) + direction test + Hamming; -1 = not a candidate
__global__ void __launch_bounds__(256) k_lines_cand(const olf_keyline* __restrict__ kl, const uint32_t* __restrict__ dl, int n1,
                                                    const uint32_t* __restrict__ dr, int n2, double inv_w, double inv_h,
                                                    const double2* __restrict__ dir2, const short2* __restrict__ cells, const int* __restrict__ ncells,
                                                    int ws, double sim_th, int* __restrict__ cand) {
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y;
    if (i2 >= n2) return;
    const olf_keyline k = kl[i1];
    const int spx = (int)__dmul_rn((double)k.startPointX, inv_w), spy = (int)__dmul_rn((double)k.startPointY, inv_h);
    const int epx = (int)__dmul_rn((double)k.endPointX, inv_w), epy = (int)__dmul_rn((double)k.endPointY, inv_h);
    // GridStructure::get windows: x in [max(0,x-ws), min(cols, x+1)), y in [max(0,y), min(rows, y+1))
    const int ax0 = max(0, spx - ws), ax1 = min(OLF_GRID_COLS, spx + 1), ay0 = max(0, spy), ay1 = min(OLF_GRID_ROWS, spy + 1);
    const int bx0 = max(0, epx - ws), bx1 = min(OLF_GRID_COLS, epx + 1), by0 = max(0, epy), by1 = min(OLF_GRID_ROWS, epy + 1);
    bool in = false;
    const int nc = ncells[i2];
    for (int c = 0; c < nc && !in; ++c) {
        const short2 p = cells[(size_t)i2 * LCELLS + c];
        in = (p.x >= ax0 && p.x < ax1 && p.y >= ay0 && p.y < ay1) || (p.x >= bx0 && p.x < bx1 && p.y >= by0 && p.y < by1);
    }
    int d = -1;
    if (in) {
        double vx = (double)(epx - spx), vy = (double)(epy - spy);
        const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
        vx = __ddiv_rn(vx, mag); vy = __ddiv_rn(vy, mag);
        const double2 o = dir2[i2];
        const double dot = __dadd_rn(__dmul_rn(vx, o.x), __dmul_rn(vy, o.y));
        if (!(fabs(dot) < sim_th)) d = hamming256(dl + (size_t)i1 * 8, dr + (size_t)i2 * 8);
    }
    cand[(size_t)i1 * n2 + i2] = d;
}
// thread per right line: walk i1 ascending, keep the running best distance (distances[i2], matches_21[i2]); a pair
// "passes" only if it improves it (src/LineMatcher.cpp:267-272).  Non-passing candidates are erased (-1).
__global__ void k_lines_pass(int* __restrict__ cand, int n1, int n2, int* __restrict__ m21) {
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i2 >= n2) return;
    int best = INT_MAX, who = -1;
    for (int i1 = 0; i1 < n1; ++i1) {
        const int d = cand[(size_t)i1 * n2 + i2];
        if (d < 0) continue;
        if (d < best) { best = d; who = i1; } else cand[(size_t)i1 * n2 + i2] = -1;
    }
    m21[i2] = who;
}
// thread per left line: best / second best over passing candidates in ascending i2, ratio test (:274-286)
__global__ void k_lines_best(const int* __restrict__ cand, int n1, int n2, double min_ratio, int* __restrict__ m12) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i1 >= n1) return;
    int best_d = INT_MAX, best_d2 = INT_MAX, best_idx = -1;
    for (int i2 = 0; i2 < n2; ++i2) {
        const int d = cand[(size_t)i1 * n2 + i2];
        if (d < 0) continue;
        if (d < best_d) { best_d2 = best_d; best_d = d; best_idx = i2; }
        else if (d < best_d2) best_d2 = d;
    }
    m12[i1] = ((double)best_d < __dmul_rn((double)best_d2, min_ratio)) ? best_idx : -1;
}
// thread per left line: mutual check (:288-296) + geometric filters (src/Frame.cc:934-958, 1002-1048), all double
__global__ void k_lines_geom(const olf_keyline* __restrict__ kl, const olf_keyline* __restrict__ kr, int n1, int* __restrict__ m12, const int* __restrict__ m21,
                             int mutual, olf_line_match_params P, float* __restrict__ disp, double* __restrict__ le) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i1 >= n1) return;
    disp[2 * i1] = -1.f; disp[2 * i1 + 1] = -1.f; le[3 * i1] = 0; le[3 * i1 + 1] = 0; le[3 * i1 + 2] = 0;
    int i2 = m12[i1];
    if (mutual && i2 >= 0 && m21[i2] != i1) { i2 = -1; m12[i1] = -1; }
    if (i2 < 0) return;
    const double spl0 = kl[i1].startPointX, spl1 = kl[i1].startPointY, epl0 = kl[i1].endPointX, epl1 = kl[i1].endPointY;
    // le_l = sp_l x ep_l (third components 1.0), normalised by the norm of its first two entries
    double l0 = __dsub_rn(__dmul_rn(spl1, 1.0), __dmul_rn(1.0, epl1));
    double l1 = __dsub_rn(__dmul_rn(1.0, epl0), __dmul_rn(spl0, 1.0));
    double l2 = __dsub_rn(__dmul_rn(spl0, epl1), __dmul_rn(spl1, epl0));
    const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)));
    l0 = __ddiv_rn(l0, nrm); l1 = __ddiv_rn(l1, nrm); l2 = __ddiv_rn(l2, nrm);
    double spr0 = kr[i2].startPointX, spr1 = kr[i2].startPointY, epr0 = kr[i2].endPointX, epr1 = kr[i2].endPointY;
    double overlap = 1.0;
    if (fabs(__dsub_rn(epl1, spl1)) > P.line_horiz_th) {
        const double sln = fmin(spl1, epl1), eln = fmax(spl1, epl1), spn = fmin(spr1, epr1), epn = fmax(spr1, epr1);
        const double length = __dsub_rn(eln, spn);
        if ((epn < sln) || (spn > eln)) overlap = 0.0;
        else if ((epn > eln) && (spn < sln)) overlap = __dsub_rn(eln, sln);
        else overlap = __dsub_rn(fmin(eln, epn), fmax(sln, spn));
        if (length > (double)0.01f) overlap = __ddiv_rn(overlap, length); else overlap = 0.0;
        if (overlap > 1.0) overlap = 1.0;
    }
    const double spr1_old = spr1;
    const double nx = __ddiv_rn(__dadd_rn(__dmul_rn(spr0, __dsub_rn(spl1, epr1)), __dmul_rn(epr0, __dsub_rn(spr1, spl1))), __dsub_rn(spr1, epr1));
    spr0 = nx; spr1 = spl1;
    const double mx = __ddiv_rn(__dadd_rn(__dmul_rn(spr0, __dsub_rn(epl1, epr1)), __dmul_rn(epr0, __dsub_rn(spr1, epl1))), __dsub_rn(spr1, epr1));
    const double epr1_old = epr1;
    epr0 = mx; epr1 = epl1;
    (void)spr1_old; (void)epr1_old;
    double disp_s = __dsub_rn(spl0, spr0), disp_e = __dsub_rn(epl0, epr0);
    if (__ddiv_rn(fmin(disp_s, disp_e), fmax(disp_s, disp_e)) < P.ls_min_disp_ratio) { disp_s = -1.0; disp_e = -1.0; }
    if (disp_s >= P.min_disp && disp_e >= P.min_disp && fabs(__dsub_rn(spl1, epl1)) > P.line_horiz_th &&
        fabs(__dsub_rn(spr1, epr1)) > P.line_horiz_th && overlap > P.stereo_overlap_th) {
        disp[2 * i1] = (float)disp_s; disp[2 * i1 + 1] = (float)disp_e;
        le[3 * i1] = l0; le[3 * i1 + 1] = l1; le[3 * i1 + 2] = l2;
    }
}

// Frame::ComputeStereoMatches_Lines for `nf` independent stereo frames: the five kernels of every frame are enqueued back to back on the context's
// stream and waited for ONCE (a rig call has four frames; a wait costs a host wake-up, and under load that is the expensive part of this phase)
int stereo_lines_batch(int nf, const olf_keyline* const* kl, const uint8_t* const* dl, const int* n1s, const olf_keyline* const* kr, const uint8_t* const* dr, const int* n2s,
                       int img_w, int img_h, const olf_line_match_params* P, int* const* matches12, float* const* disp, double* const* le, int device) {
    if (!P || nf < 1 || nf > 16 || img_w <= 0 || img_h <= 0) { set_last_error("olf_stereo_lines: bad arguments"); return OLF_ERR_ARG; }
    for (int f = 0; f < nf; ++f) {
        const int n1 = n1s[f], n2 = n2s[f];
        if (n1 < 0 || n2 < 0 || (n1 && (!kl[f] || !dl[f] || !matches12[f] || !disp[f] || !le[f])) || (n2 && (!kr[f] || !dr[f]))) { set_last_error("olf_stereo_lines: bad arguments"); return OLF_ERR_ARG; }
    }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    const double inv_w = OLF_GRID_COLS / (double)img_w, inv_h = OLF_GRID_ROWS / (double)img_h;
    struct Off { size_t kl, dl, kr, dr, dir, cells, nc, cand, m12, m21, disp, le, p_in, p_m, p_disp, p_le; bool live; } off[16];
    Planner pl;
    bool any = false;
    for (int f = 0; f < nf; ++f) {
        const int n1 = n1s[f], n2 = n2s[f];
        for (int i = 0; i < n1; ++i) { matches12[f][i] = -1; disp[f][2 * i] = disp[f][2 * i + 1] = -1.f; le[f][3 * i] = le[f][3 * i + 1] = le[f][3 * i + 2] = 0; }
        Off& o = off[f];
        o.live = n1 > 0 && n2 > 0;
        if (!o.live) continue;
        any = true;
        o.kl = pl.d((size_t)n1 * sizeof(olf_keyline)); o.dl = pl.d((size_t)n1 * 32); o.kr = pl.d((size_t)n2 * sizeof(olf_keyline)); o.dr = pl.d((size_t)n2 * 32);
        o.dir = pl.d((size_t)n2 * sizeof(double2)); o.cells = pl.d((size_t)n2 * LCELLS * sizeof(short2)); o.nc = pl.d((size_t)n2 * 4);
        o.cand = pl.d((size_t)n1 * n2 * 4); o.m12 = pl.d((size_t)n1 * 4); o.m21 = pl.d((size_t)n2 * 4); o.disp = pl.d((size_t)n1 * 8); o.le = pl.d((size_t)n1 * 24);
        o.p_in = pl.p((size_t)(n1 + n2) * (sizeof(olf_keyline) + 32)); o.p_m = pl.p((size_t)n1 * 4); o.p_disp = pl.p((size_t)n1 * 8); o.p_le = pl.p((size_t)n1 * 24);
    }
    if (!any) return OLF_OK;
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    for (int f = 0; f < nf; ++f) {
        const Off& o = off[f];
        if (!o.live) continue;
        const int n1 = n1s[f], n2 = n2s[f];
        uint8_t* h_kl = hptr<uint8_t>(c, o.p_in); uint8_t* h_dl = h_kl + (size_t)n1 * sizeof(olf_keyline); uint8_t* h_kr = h_dl + (size_t)n1 * 32; uint8_t* h_dr = h_kr + (size_t)n2 * sizeof(olf_keyline);
        memcpy(h_kl, kl[f], (size_t)n1 * sizeof(olf_keyline)); memcpy(h_dl, dl[f], (size_t)n1 * 32); memcpy(h_kr, kr[f], (size_t)n2 * sizeof(olf_keyline)); memcpy(h_dr, dr[f], (size_t)n2 * 32);
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o.kl), h_kl, (size_t)n1 * sizeof(olf_keyline), cudaMemcpyHostToDevice, s));
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o.dl), h_dl, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o.kr), h_kr, (size_t)n2 * sizeof(olf_keyline), cudaMemcpyHostToDevice, s));
        OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o.dr), h_dr, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
        k_lines_raster<<<(n2 + 127) / 128, 128, 0, s>>>(dptr<olf_keyline>(c, o.kr), n2, inv_w, inv_h, dptr<double2>(c, o.dir), dptr<short2>(c, o.cells), dptr<int>(c, o.nc));
        k_lines_cand<<<dim3((n2 + 255) / 256, n1), 256, 0, s>>>(dptr<olf_keyline>(c, o.kl), dptr<uint32_t>(c, o.dl), n1, dptr<uint32_t>(c, o.dr), n2, inv_w, inv_h,
                                                                dptr<double2>(c, o.dir), dptr<short2>(c, o.cells), dptr<int>(c, o.nc), P->matching_s_ws, P->line_sim_th, dptr<int>(c, o.cand));
        count_launches(P->best_lr_matches ? 5 : 4);
        if (P->best_lr_matches) k_lines_pass<<<(n2 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o.cand), n1, n2, dptr<int>(c, o.m21));
        k_lines_best<<<(n1 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o.cand), n1, n2, P->min_ratio_12_l, dptr<int>(c, o.m12));
        k_lines_geom<<<(n1 + 127) / 128, 128, 0, s>>>(dptr<olf_keyline>(c, o.kl), dptr<olf_keyline>(c, o.kr), n1, dptr<int>(c, o.m12), dptr<int>(c, o.m21), P->best_lr_matches, *P,
                                                      dptr<float>(c, o.disp), dptr<double>(c, o.le));
        OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, o.p_m), dptr<int>(c, o.m12), (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaMemcpyAsync(hptr<float>(c, o.p_disp), dptr<float>(c, o.disp), (size_t)n1 * 8, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaMemcpyAsync(hptr<double>(c, o.p_le), dptr<double>(c, o.le), (size_t)n1 * 24, cudaMemcpyDeviceToHost, s));
    }
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    for (int f = 0; f < nf; ++f) {
        const Off& o = off[f];
        if (!o.live) continue;
        const int n1 = n1s[f];
        memcpy(matches12[f], hptr<int>(c, o.p_m), (size_t)n1 * 4); memcpy(disp[f], hptr<float>(c, o.p_disp), (size_t)n1 * 8); memcpy(le[f], hptr<double>(c, o.p_le), (size_t)n1 * 24);
    }
    return OLF_OK;
}
int stereo_lines(const olf_keyline* kl, const uint8_t* dl, int n1, const olf_keyline* kr, const uint8_t* dr, int n2, int img_w, int img_h,
                 const olf_line_match_params* P, int* matches12, float* disp, double* le, int device) {
    return stereo_lines_batch(1, &kl, &dl, &n1, &kr, &dr, &n2, img_w, img_h, P, &matches12, &disp, &le, device);
}

// ---- matchGrid(lines) as a stand-alone entry (src/LineMatcher.cpp:220-299) for callers that build the GridStructure
// themselves (the reference's Frame::ComputeStereoMatches_Lines, src/Frame.cc:896-927): grid cells -> per right line cell list,
// the candidate test with a general GridWindow, then the same passes as olf_stereo_lines.
__global__ void __launch_bounds__(256) k_lines_cand_grid(const int4* __restrict__ l1, const uint32_t* __restrict__ dl, int n1,
                                                         const uint32_t* __restrict__ dr, int n2, const double2* __restrict__ dir2,
                                                         const short2* __restrict__ cells, const int* __restrict__ ncells, int rows, int cols,
                                                         int4 win /* width.first, width.second, height.first, height.second */, double sim_th, int* __restrict__ cand) {
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y;
    if (i2 >= n2) return;
    const int4 L = l1[i1];
    // GridStructure::get (src/gridStructure.cpp:68-79): x in [max(0,x-w.first), min(cols, x+w.second+1)), same for y
    const int ax0 = max(0, L.x - win.x), ax1 = min(cols, L.x + win.y + 1), ay0 = max(0, L.y - win.z), ay1 = min(rows, L.y + win.w + 1);
    const int bx0 = max(0, L.z - win.x), bx1 = min(cols, L.z + win.y + 1), by0 = max(0, L.w - win.z), by1 = min(rows, L.w + win.w + 1);
    bool in = false;
    const int nc = ncells[i2];
    for (int c = 0; c < nc && !in; ++c) {
        const short2 p = cells[(size_t)i2 * LCELLS + c];
        in = (p.x >= ax0 && p.x < ax1 && p.y >= ay0 && p.y < ay1) || (p.x >= bx0 && p.x < bx1 && p.y >= by0 && p.y < by1);
    }
    int d = -1;
    if (in) {
        double vx = (double)(L.z - L.x), vy = (double)(L.w - L.y);
        const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
        vx = __ddiv_rn(vx, mag); vy = __ddiv_rn(vy, mag);
        const double2 o = dir2[i2];
        const double dot = __dadd_rn(__dmul_rn(vx, o.x), __dmul_rn(vy, o.y));
        if (!(fabs(dot) < sim_th)) d = hamming256(dl + (size_t)i1 * 8, dr + (size_t)i2 * 8);
    }
    cand[(size_t)i1 * n2 + i2] = d;
}
__global__ void k_lines_mutual(int* __restrict__ m12, const int* __restrict__ m21, int n1, int mutual) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i1 >= n1) return;
    const int i2 = m12[i1];
    if (mutual && i2 >= 0 && m21[i2] != i1) m12[i1] = -1;
}
int match_grid_lines(const int* lines1, const uint8_t* desc1, int n1, const olf_grid_csr* grid, const uint8_t* desc2, int n2, const double* dir2,
                     const int* window, const olf_line_match_params* P, int* matches12, int* nmatches, int device) {
    if (!P || !grid || !window || !nmatches || n1 < 0 || n2 < 0 || grid->rows <= 0 || grid->cols <= 0 || !grid->cell_begin ||
        (n1 && (!lines1 || !desc1 || !matches12)) || (n2 && (!desc2 || !dir2))) { set_last_error("olf_match_grid_lines: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return OLF_OK;
    Planner pl;
    const size_t o_l1 = pl.d((size_t)n1 * 16), o_dl = pl.d((size_t)n1 * 32), o_dr = pl.d((size_t)n2 * 32), o_dir = pl.d((size_t)n2 * 16);
    const size_t o_cells = pl.d((size_t)n2 * LCELLS * sizeof(short2)), o_nc = pl.d((size_t)n2 * 4), o_cand = pl.d((size_t)n1 * n2 * 4), o_m12 = pl.d((size_t)n1 * 4), o_m21 = pl.d((size_t)n2 * 4);
    const size_t p_l1 = pl.p((size_t)n1 * 16), p_dl = pl.p((size_t)n1 * 32), p_dr = pl.p((size_t)n2 * 32), p_dir = pl.p((size_t)n2 * 16);
    const size_t p_cells = pl.p((size_t)n2 * LCELLS * sizeof(short2)), p_nc = pl.p((size_t)n2 * 4), p_m = pl.p((size_t)n1 * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    memcpy(hptr<uint8_t>(c, p_l1), lines1, (size_t)n1 * 16); memcpy(hptr<uint8_t>(c, p_dl), desc1, (size_t)n1 * 32);
    memcpy(hptr<uint8_t>(c, p_dr), desc2, (size_t)n2 * 32); memcpy(hptr<uint8_t>(c, p_dir), dir2, (size_t)n2 * 16);
    // cell lists per right line from the grid (an index outside [0, n2) is never a candidate, :256)
    short2* hc = hptr<short2>(c, p_cells); int* hn = hptr<int>(c, p_nc);
    for (int i = 0; i < n2; ++i) hn[i] = 0;
    for (int x = 0; x < grid->cols; ++x)
        for (int y = 0; y < grid->rows; ++y) {
            const int cell = x * grid->rows + y;
            for (int k = grid->cell_begin[cell]; k < grid->cell_begin[cell + 1]; ++k) {
                const int i2 = grid->items[k];
                if (i2 < 0 || i2 >= n2) continue;
                if (hn[i2] >= LCELLS) { set_last_error("olf_match_grid_lines: a line covers more than 128 grid cells"); return OLF_ERR_CAPACITY; }
                hc[(size_t)i2 * LCELLS + hn[i2]++] = make_short2((short)x, (short)y);
            }
        }
    auto up = [&](size_t od, size_t op, size_t bytes) { return cudaMemcpyAsync(dptr<uint8_t>(c, od), hptr<uint8_t>(c, op), bytes, cudaMemcpyHostToDevice, s); };
    OLF_CUDA(up(o_l1, p_l1, (size_t)n1 * 16)); OLF_CUDA(up(o_dl, p_dl, (size_t)n1 * 32)); OLF_CUDA(up(o_dr, p_dr, (size_t)n2 * 32)); OLF_CUDA(up(o_dir, p_dir, (size_t)n2 * 16));
    OLF_CUDA(up(o_cells, p_cells, (size_t)n2 * LCELLS * sizeof(short2))); OLF_CUDA(up(o_nc, p_nc, (size_t)n2 * 4));
    k_lines_cand_grid<<<dim3((n2 + 255) / 256, n1), 256, 0, s>>>(dptr<int4>(c, o_l1), dptr<uint32_t>(c, o_dl), n1, dptr<uint32_t>(c, o_dr), n2, dptr<double2>(c, o_dir),
                                                                 dptr<short2>(c, o_cells), dptr<int>(c, o_nc), grid->rows, grid->cols,
                                                                 make_int4(window[0], window[1], window[2], window[3]), P->line_sim_th, dptr<int>(c, o_cand));
    if (P->best_lr_matches) k_lines_pass<<<(n2 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o_cand), n1, n2, dptr<int>(c, o_m21));
    k_lines_best<<<(n1 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o_cand), n1, n2, P->min_ratio_12_l, dptr<int>(c, o_m12));
    k_lines_mutual<<<(n1 + 127) / 128, 128, 0, s>>>(dptr<int>(c, o_m12), dptr<int>(c, o_m21), n1, P->best_lr_matches);
    count_launches(P->best_lr_matches ? 4 : 3);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_m), dptr<int>(c, o_m12), (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    int cnt = 0;
    for (int i = 0; i < n1; ++i) { matches12[i] = hptr<int>(c, p_m)[i]; cnt += matches12[i] >= 0; }
    *nmatches = cnt;
    return OLF_OK;
}

// ======================================================================================================
// MapPoint / MapLine::ComputeDistinctiveDescriptors (src/MapPoint.cc:254-322, src/MapLine.cc:257-322), batched over landmarks:
// the observation whose MEDIAN Hamming distance to all observations of the landmark (itself included, distance 0) is smallest,
// first one on ties.  One warp per landmark: a row of the distance matrix at a time, the (int)(0.5*(N-1))-th smallest of the
// row from a 257-bin histogram in shared memory (distances are integers in [0, 256]).
// ======================================================================================================
__global__ void __launch_bounds__(256) k_distinctive(const uint32_t* __restrict__ desc, const int* __restrict__ begin, int n_groups, int* __restrict__ best) {
    __shared__ unsigned hist[8][264];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + warp;
    if (g >= n_groups) return;
    const int b = begin[g], N = begin[g + 1] - b;
    if (N <= 0) { if (lane == 0) best[g] = -1; return; }
    const int kth = (int)(0.5 * (double)(N - 1));
    int best_median = INT_MAX, best_idx = 0;
    for (int i = 0; i < N; ++i) {
        for (int k = lane; k < 264; k += 32) hist[warp][k] = 0;
        __syncwarp();
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = desc[(size_t)(b + i) * 8 + k];
        for (int j = lane; j < N; j += 32) atomicAdd(&hist[warp][hamming256(a, desc + (size_t)(b + j) * 8)], 1u);
        __syncwarp();
        // smallest d with cumulative count > kth: per-lane partial sums over 9 consecutive bins, warp scan
        unsigned part = 0;
        for (int k = 0; k < 9; ++k) { const int d = lane * 9 + k; if (d <= 256) part += hist[warp][d]; }
        unsigned incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const unsigned excl = incl - part;
        int median = INT_MAX;
        if (excl <= (unsigned)kth && (unsigned)kth < incl) {
            unsigned run = excl;
            for (int k = 0; k < 9; ++k) { const int d = lane * 9 + k; if (d > 256) break; run += hist[warp][d]; if ((unsigned)kth < run) { median = d; break; } }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) median = min(median, __shfl_xor_sync(0xffffffffu, median, o));
        if (median < best_median) { best_median = median; best_idx = i; }
        __syncwarp();
    }
    if (lane == 0) best[g] = best_idx;
}
int distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best, int device) {
    if (n_groups < 0 || (n_groups && (!group_begin || !best))) { set_last_error("olf_distinctive_descriptors: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    if (n_groups == 0) return OLF_OK;
    const int total = group_begin[n_groups];
    if (total < 0 || (total && !desc)) { set_last_error("olf_distinctive_descriptors: bad arguments"); return OLF_ERR_ARG; }
    Planner pl;
    const size_t o_d = pl.d((size_t)std::max(total, 1) * 32), o_b = pl.d((size_t)(n_groups + 1) * 4), o_o = pl.d((size_t)n_groups * 4);
    const size_t p_d = pl.p((size_t)std::max(total, 1) * 32), p_b = pl.p((size_t)(n_groups + 1) * 4), p_o = pl.p((size_t)n_groups * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    if (total) memcpy(hptr<uint8_t>(c, p_d), desc, (size_t)total * 32);
    memcpy(hptr<int>(c, p_b), group_begin, (size_t)(n_groups + 1) * 4);
    if (total) OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_d), hptr<uint8_t>(c, p_d), (size_t)total * 32, cudaMemcpyHostToDevice, s));
    OLF_CUDA(cudaMemcpyAsync(dptr<int>(c, o_b), hptr<int>(c, p_b), (size_t)(n_groups + 1) * 4, cudaMemcpyHostToDevice, s));
    k_distinctive<<<(n_groups + 7) / 8, 256, 0, s>>>(dptr<uint32_t>(c, o_d), dptr<int>(c, o_b), n_groups, dptr<int>(c, o_o));
    count_launches(1);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_o), dptr<int>(c, o_o), (size_t)n_groups * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(best, hptr<int>(c, p_o), (size_t)n_groups * 4);
    return OLF_OK;
}

// ======================================================================================================
// ORBmatcher::SearchByProjection: candidate lists (warp per query) + fixed-point resolution of the blocking rule
// ======================================================================================================
#define SBP_K 128
struct SbpQuery { float u, v, radius, ur; int min_level, max_level; int valid; };
struct GridParams { float min_x, min_y, inv_w, inv_h; };
typedef unsigned long long u64;

// Candidates of query i = Frame::GetFeaturesInArea(u, v, radius, min_level, max_level) (src/Frame.cc:517-570) minus
// those failing the stereo check |ur - uRight[j]| > radius.  The reference enumerates cells ix-outer / iy-inner and
// keypoints in ascending index inside a cell, and keeps the FIRST minimum, so the preference order is the
// lexicographic key (dist, ix, iy, j).  Lists are written sorted by that key.
__global__ void __launch_bounds__(256) k_sbp_candidates(const SbpQuery* __restrict__ q, const uint32_t* __restrict__ qdesc, int nq,
                                                        const olf_keypoint* __restrict__ kps, const uint32_t* __restrict__ desc, const float* __restrict__ uRight, int n_cur,
                                                        GridParams G, int max_dist, int gate, const float* __restrict__ inv_sigma2,
                                                        u64* __restrict__ lists, int* __restrict__ counts,
                                                        const unsigned* __restrict__ off = nullptr, u64* __restrict__ stage_g = nullptr) {
    // gate 1: |ur - uRight[j]| <= radius for stereo keypoints (SearchByProjection last frame / local map, :1556-1561, :93-98)
    // gate 2: chi-square reprojection gate of Fuse(KF, MPs, th) (:916-940); gate 0: none (KeyFrame window searches)
    // Lists: SBP_K slots per query, staged in shared memory (off == nullptr); a window that holds more makes the host run the
    // pass again with exact per-query capacities (CSR offsets `off` from the counts of the first pass, staging in global memory)
    __shared__ u64 sh[8][SBP_K];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + w;
    if (i >= nq) return;
    const SbpQuery Q = q[i];
    u64* const stage = off ? stage_g + off[i] : sh[w];
    u64* const out = off ? lists + off[i] : lists + (size_t)i * SBP_K;
    const int cap = off ? (int)(off[i + 1] - off[i]) : SBP_K;
    int n = 0;
    if (Q.valid) {
        const int nMinCellX = max(0, (int)floorf(fmul(fsub(fsub(Q.u, G.min_x), Q.radius), G.inv_w)));
        const int nMaxCellX = min(OLF_GRID_COLS - 1, (int)ceilf(fmul(fadd(fsub(Q.u, G.min_x), Q.radius), G.inv_w)));
        const int nMinCellY = max(0, (int)floorf(fmul(fsub(fsub(Q.v, G.min_y), Q.radius), G.inv_h)));
        const int nMaxCellY = min(OLF_GRID_ROWS - 1, (int)ceilf(fmul(fadd(fsub(Q.v, G.min_y), Q.radius), G.inv_h)));
        const bool ok = !(nMinCellX >= OLF_GRID_COLS) && !(nMaxCellX < 0) && !(nMinCellY >= OLF_GRID_ROWS) && !(nMaxCellY < 0);
        const bool check_levels = (Q.min_level > 0) || (Q.max_level >= 0);
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = qdesc[(size_t)i * 8 + k];
        for (int base = 0; base < n_cur && ok; base += 32) {
            const int j = base + lane;
            bool take = false; u64 key = 0;
            if (j < n_cur) {
                const olf_keypoint kp = kps[j];
                const int posX = (int)roundf(fmul(fsub(kp.x, G.min_x), G.inv_w)), posY = (int)roundf(fmul(fsub(kp.y, G.min_y), G.inv_h));   // PosInGrid
                take = !(posX < 0 || posX >= OLF_GRID_COLS || posY < 0 || posY >= OLF_GRID_ROWS) &&
                       posX >= nMinCellX && posX <= nMaxCellX && posY >= nMinCellY && posY <= nMaxCellY;
                if (take && check_levels) { if (kp.octave < Q.min_level) take = false; if (Q.max_level >= 0 && kp.octave > Q.max_level) take = false; }
                if (take) take = fabsf(fsub(kp.x, Q.u)) < Q.radius && fabsf(fsub(kp.y, Q.v)) < Q.radius;
                if (take && gate == 1) { const float ur = uRight[j]; if (ur > 0 && fabsf(fsub(Q.ur, ur)) > Q.radius) take = false; }
                if (take && gate == 2) {
                    const float ex = fsub(Q.u, kp.x), ey = fsub(Q.v, kp.y), kpr = uRight[j];
                    float e2 = fadd(fmul(ex, ex), fmul(ey, ey));
                    double lim = 5.99;
                    if (kpr >= 0) { const float er = fsub(Q.ur, kpr); e2 = fadd(e2, fmul(er, er)); lim = 7.8; }
                    if ((double)fmul(e2, inv_sigma2[kp.octave]) > lim) take = false;
                }
                if (take) {
                    const int d = hamming256(a, desc + (size_t)j * 8);
                    if (d > max_dist) take = false;
                    key = ((u64)d << 40) | ((u64)posX << 32) | ((u64)posY << 24) | (u64)j;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) { const int p = n + __popc(m & ((1u << lane) - 1)); if (p < cap) stage[p] = key; }
            n += __popc(m);
        }
    }
    __syncwarp();
    const int stored = min(n, cap);
    // rank sort (keys are distinct: they embed j)
    for (int e = lane; e < stored; e += 32) {
        const u64 key = stage[e];
        int rank = 0;
        for (int f = 0; f < stored; ++f) rank += stage[f] < key;
        out[rank] = key;
    }
    if (lane == 0) counts[i] = n;          // n > SBP_K in the first pass: the host runs the exact-capacity pass
}

// One block.  Iterates  assign[i] = first candidate of i not owned by an earlier observed point  to its fixed point.
// mode 0 (last frame, src/ORBmatcher.cc:1536-1593): best only, accept dist <= TH_HIGH (lists are pre-filtered).
// mode 1 (local map, src/ORBmatcher.cc:79-127): best + second best with level / ratio test.
__global__ void __launch_bounds__(1024) k_sbp_resolve(const u64* __restrict__ lists, const int* __restrict__ counts, int nq, int n_cur,
                                                      const uint8_t* __restrict__ observed, const uint8_t* __restrict__ occupied,
                                                      const olf_keypoint* __restrict__ kps, int mode, float nn_ratio,
                                                      int* __restrict__ owner_a, int* __restrict__ owner_b, int* __restrict__ assign, int* __restrict__ rounds_out,
                                                      int* __restrict__ adist = nullptr, const unsigned* __restrict__ off = nullptr) {
    __shared__ int s_changed;
    int* own_prev = owner_a; int* own_new = owner_b;
    for (int j = threadIdx.x; j < n_cur; j += 1024) { own_prev[j] = (occupied && occupied[j]) ? -1 : INT_MAX; }
    for (int i = threadIdx.x; i < nq; i += 1024) assign[i] = -1;
    __syncthreads();
    int rounds = 0;
    for (;;) {
        if (threadIdx.x == 0) s_changed = 0;
        for (int j = threadIdx.x; j < n_cur; j += 1024) own_new[j] = (occupied && occupied[j]) ? -1 : INT_MAX;
        __syncthreads();
        for (int i = threadIdx.x; i < nq; i += 1024) {
            const int cnt = off ? counts[i] : min(counts[i], SBP_K);
            const u64* const L = off ? lists + off[i] : lists + (size_t)i * SBP_K;
            int sel = -1, seld = 256;
            if (mode == 0) {
                for (int e = 0; e < cnt; ++e) { const u64 key = L[e]; const int j = (int)(key & 0xFFFFFF); if (!(own_prev[j] < i)) { sel = j; seld = (int)(key >> 40); break; } }
            } else {
                int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1, found = 0;
                for (int e = 0; e < cnt && found < 2; ++e) {
                    const u64 key = L[e];
                    const int j = (int)(key & 0xFFFFFF), d = (int)(key >> 40);
                    if (own_prev[j] < i) continue;
                    if (found == 0) { if (d < bestDist) { bestDist = d; bestLevel = kps[j].octave; bestIdx = j; } }
                    else if (d < bestDist2) { bestDist2 = d; bestLevel2 = kps[j].octave; }
                    ++found;
                }
                if (bestIdx >= 0 && bestDist <= OLF_TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > fmul(nn_ratio, (float)bestDist2))) sel = bestIdx;
            }
            if (sel != assign[i]) { assign[i] = sel; s_changed = 1; }
            if (adist) adist[i] = seld;                    // mode 0 only: the Hamming distance of the chosen candidate
            if (sel >= 0 && observed[i]) atomicMin(&own_new[sel], i);
        }
        __syncthreads();
        ++rounds;
        const int changed = s_changed;
        int* t = own_prev; own_prev = own_new; own_new = t;
        __syncthreads();
        if (!changed || rounds > 4 * nq + 8) break;
    }
    if (threadIdx.x == 0) *rounds_out = rounds;
}

// A window that holds more than SBP_K candidates (dense texture, SearchForInitialization's 100-pixel window) sends the call through the pass
// again with exact per-window capacities: CSR offsets from the counts of the first pass.  The lists of all windows together are bounded
// (16 bytes per candidate on the device); beyond that the call fails loudly.
#define SBP_CSR_MAX ((size_t)32 << 20)
static int csr_from_counts(const int* cnt, int nq, std::vector<unsigned>& off, const char* who) {
    off.assign((size_t)nq + 1, 0u);
    size_t total = 0;
    for (int i = 0; i < nq; ++i) {
        off[i] = (unsigned)total; total += (size_t)std::max(cnt[i], 0);
        if (total > SBP_CSR_MAX) { set_last_error(std::string(who) + ": the search windows hold more than 2^25 candidates together"); return OLF_ERR_CAPACITY; }
    }
    off[nq] = (unsigned)total;
    return OLF_OK;
}

static int sbp_common(MatchCtx* c, const std::vector<SbpQuery>& q, const uint8_t* qdesc, const uint8_t* observed, const uint8_t* occupied,
                      const olf_keypoint* cur_kps, const uint8_t* cur_desc, const float* cur_u_right, int n_cur, const olf_camera& cam,
                      int mode, int max_dist, float nn_ratio, int* assign_out, int gate = 1, const float* inv_sigma2 = nullptr, int nlevels = 0, int* dist_out = nullptr) {
    const int nq = (int)q.size();
    int rc;
    if (n_cur >= (1 << 24)) { set_last_error("olf_search_by_projection: too many keypoints"); return OLF_ERR_CAPACITY; }
    std::vector<unsigned> csr;                                  // empty: first pass, SBP_K slots per window
    for (int pass = 0; pass < 2; ++pass) {
        const bool wide = !csr.empty();
        const size_t slots = wide ? (size_t)csr.back() : (size_t)nq * SBP_K;
        Planner pl;
        const size_t o_q = pl.d((size_t)nq * sizeof(SbpQuery)), o_qd = pl.d((size_t)nq * 32), o_obs = pl.d(nq), o_occ = pl.d(std::max(n_cur, 1));
        const size_t o_k = pl.d((size_t)n_cur * sizeof(olf_keypoint)), o_d = pl.d((size_t)n_cur * 32), o_u = pl.d((size_t)n_cur * 4);
        const size_t o_lists = pl.d(slots * 8), o_cnt = pl.d((size_t)nq * 4), o_oa = pl.d((size_t)n_cur * 4), o_ob = pl.d((size_t)n_cur * 4), o_as = pl.d((size_t)nq * 4), o_r = pl.d(4);
        const size_t o_sig = pl.d((size_t)OLF_MAX_LEVELS * 4), o_ad = pl.d((size_t)nq * 4);
        const size_t o_stage = pl.d(wide ? slots * 8 : 0), o_off = pl.d(wide ? ((size_t)nq + 1) * 4 : 0);
        const size_t p_in = pl.p((size_t)nq * (sizeof(SbpQuery) + 33) + (size_t)n_cur * (sizeof(olf_keypoint) + 37) + 64 + OLF_MAX_LEVELS * 4 + (wide ? ((size_t)nq + 1) * 4 : 0)),
                     p_out = pl.p((size_t)nq * 12 + 16);
        if ((rc = arena_ensure(c, pl))) return rc;
        cudaStream_t s = c->cur;
        uint8_t* hp = hptr<uint8_t>(c, p_in);
        size_t off = 0;
        auto up = [&](size_t dev_off, const void* src, size_t bytes) -> int {
            if (!bytes) return OLF_OK;
            memcpy(hp + off, src, bytes);
            cudaError_t e = cudaMemcpyAsync(c->a.dev.p + dev_off, hp + off, bytes, cudaMemcpyHostToDevice, s);
            off += bytes;
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__);
            return OLF_OK;
        };
        if (wide && (rc = up(o_off, csr.data(), csr.size() * 4))) return rc;
        if ((rc = up(o_q, q.data(), (size_t)nq * sizeof(SbpQuery))) || (rc = up(o_qd, qdesc, (size_t)nq * 32)) || (rc = up(o_obs, observed, nq)) ||
            (rc = up(o_k, cur_kps, (size_t)n_cur * sizeof(olf_keypoint))) || (rc = up(o_d, cur_desc, (size_t)n_cur * 32))) return rc;
        if (cur_u_right && (rc = up(o_u, cur_u_right, (size_t)n_cur * 4))) return rc;
        if (gate == 2 && (rc = up(o_sig, inv_sigma2, (size_t)std::min(nlevels, OLF_MAX_LEVELS) * 4))) return rc;
        if (occupied && (rc = up(o_occ, occupied, n_cur))) return rc;
        GridParams G; G.min_x = cam.min_x; G.min_y = cam.min_y;
        G.inv_w = (float)OLF_GRID_COLS / (cam.max_x - cam.min_x); G.inv_h = (float)OLF_GRID_ROWS / (cam.max_y - cam.min_y);       // src/Frame.cc:185-186
        const unsigned* d_off = wide ? dptr<unsigned>(c, o_off) : nullptr;
        k_sbp_candidates<<<(nq + 7) / 8, 256, 0, s>>>(dptr<SbpQuery>(c, o_q), dptr<uint32_t>(c, o_qd), nq, dptr<olf_keypoint>(c, o_k), dptr<uint32_t>(c, o_d), dptr<float>(c, o_u), n_cur,
                                                      G, max_dist, gate, dptr<float>(c, o_sig), dptr<u64>(c, o_lists), dptr<int>(c, o_cnt), d_off, wide ? dptr<u64>(c, o_stage) : nullptr);
        k_sbp_resolve<<<1, 1024, 0, s>>>(dptr<u64>(c, o_lists), dptr<int>(c, o_cnt), nq, n_cur, dptr<uint8_t>(c, o_obs), occupied ? dptr<uint8_t>(c, o_occ) : nullptr,
                                         dptr<olf_keypoint>(c, o_k), mode, nn_ratio, dptr<int>(c, o_oa), dptr<int>(c, o_ob), dptr<int>(c, o_as), dptr<int>(c, o_r),
                                         dist_out ? dptr<int>(c, o_ad) : nullptr, d_off);
        count_launches(2);
        int* ho = hptr<int>(c, p_out);
        OLF_CUDA(cudaMemcpyAsync(ho, dptr<int>(c, o_as), (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaMemcpyAsync(ho + nq, dptr<int>(c, o_cnt), (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
        if (dist_out) OLF_CUDA(cudaMemcpyAsync(ho + 2 * nq, dptr<int>(c, o_ad), (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaGetLastError());
        OLF_CUDA(stream_sync(s));
        bool overflow = false;
        for (int i = 0; i < nq && !wide; ++i) overflow |= ho[nq + i] > SBP_K;
        if (overflow) {
            if ((rc = csr_from_counts(ho + nq, nq, csr, "olf_search_by_projection"))) return rc;
            continue;
        }
        for (int i = 0; i < nq; ++i) {
            if (wide && ho[nq + i] != (int)(csr[i + 1] - csr[i])) { set_last_error("olf_search_by_projection: candidate counts changed between the passes"); return OLF_ERR_INTERNAL; }
            assign_out[i] = ho[i];
            if (dist_out) dist_out[i] = ho[i] >= 0 ? ho[2 * nq + i] : 256;
        }
        return OLF_OK;
    }
    set_last_error("olf_search_by_projection: internal error"); return OLF_ERR_INTERNAL;
}

static void mat3_mul_vec(const float* R, const float* v, float* o) { for (int r = 0; r < 3; ++r) o[r] = R[3 * r] * v[0] + R[3 * r + 1] * v[1] + R[3 * r + 2] * v[2]; }

int search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches_out, int device) {
    if (!a || !assigned_cur || !cur_point || !nmatches_out || a->n_cur < 0 || a->n_last < 0 || !a->scale_factors) { set_last_error("olf_search_by_projection_last: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    for (int j = 0; j < a->n_cur; ++j) cur_point[j] = -1;
    for (int i = 0; i < a->n_last; ++i) assigned_cur[i] = -1;
    *nmatches_out = 0;
    if (a->n_cur == 0 || a->n_last == 0) return OLF_OK;
    // host part of the reference function: motion direction and per-point projection (:1485-1531), float ops in order
    float twc[3], tlc[3];
    for (int r = 0; r < 3; ++r) twc[r] = -(a->Rcw[r] * a->tcw[0] + a->Rcw[3 + r] * a->tcw[1] + a->Rcw[6 + r] * a->tcw[2]);
    mat3_mul_vec(a->Rlw, twc, tlc);
    for (int r = 0; r < 3; ++r) tlc[r] = tlc[r] + a->tlw[r];
    const float mb = a->cam.bf / a->cam.fx;
    const bool bForward = tlc[2] > mb && !a->mono, bBackward = -tlc[2] > mb && !a->mono;
    std::vector<SbpQuery> q(a->n_last);
    for (int i = 0; i < a->n_last; ++i) {
        SbpQuery& Q = q[i];
        memset(&Q, 0, sizeof(Q));
        if (!a->last_has_point[i]) continue;
        float x3Dc[3];
        mat3_mul_vec(a->Rcw, a->last_world_pos + 3 * i, x3Dc);
        for (int r = 0; r < 3; ++r) x3Dc[r] = x3Dc[r] + a->tcw[r];
        const float invzc = (float)(1.0 / x3Dc[2]);
        if (invzc < 0) continue;
        const float u = a->cam.fx * x3Dc[0] * invzc + a->cam.cx, v = a->cam.fy * x3Dc[1] * invzc + a->cam.cy;
        if (u < a->cam.min_x || u > a->cam.max_x) continue;
        if (v < a->cam.min_y || v > a->cam.max_y) continue;
        const int oct = a->last_kps[i].octave;
        Q.u = u; Q.v = v; Q.radius = a->th * a->scale_factors[oct]; Q.ur = u - a->cam.bf * invzc;
        if (bForward) { Q.min_level = oct; Q.max_level = -1; }
        else if (bBackward) { Q.min_level = 0; Q.max_level = oct; }
        else { Q.min_level = oct - 1; Q.max_level = oct + 1; }
        Q.valid = 1;
    }
    if ((rc = sbp_common(c, q, a->last_point_desc, a->last_point_observed, a->cur_occupied, a->cur_kps, a->cur_desc, a->cur_u_right, a->n_cur, a->cam, 0, OLF_TH_HIGH, 0.f, assigned_cur))) return rc;
    // bookkeeping of the reference on the resolved assignments (:1564-1615): last writer wins, rotation histogram pruning
    int nmatches = 0;
    std::vector<int> rot[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    for (int i = 0; i < a->n_last; ++i) {
        const int j = assigned_cur[i];
        if (j < 0) continue;
        cur_point[j] = i; nmatches++;
        if (a->check_orientation) {
            float r = a->last_kps[i].angle - a->cur_kps[j].angle;
            if (r < 0.0) r += 360.0f;
            int bin = (int)roundf(r * factor);
            if (bin == OLF_HISTO_LENGTH) bin = 0;
            rot[bin].push_back(j);
        }
    }
    if (a->check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;          // ComputeThreeMaxima (:1749-1790)
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            const int sz = (int)rot[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < OLF_HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rot[i]) { cur_point[j] = -1; nmatches--; }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}

int search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur, int* nmatches_out, int device) {
    if (!a || !assigned_cur || !nmatches_out || a->n_cur < 0 || a->n_points < 0 || !a->scale_factors) { set_last_error("olf_search_by_projection_map: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    for (int i = 0; i < a->n_points; ++i) assigned_cur[i] = -1;
    *nmatches_out = 0;
    if (a->n_cur == 0 || a->n_points == 0) return OLF_OK;
    const bool bFactor = a->th != 1.0;
    std::vector<SbpQuery> q(a->n_points);
    for (int i = 0; i < a->n_points; ++i) {
        SbpQuery& Q = q[i];
        const int lvl = a->pred_level[i];
        float r = (a->view_cos[i] > 0.998) ? 2.5f : 4.0f;          // RadiusByViewingCos (:133-139)
        if (bFactor) r *= a->th;
        Q.u = a->proj_x[i]; Q.v = a->proj_y[i]; Q.radius = r * a->scale_factors[lvl]; Q.ur = a->proj_xr[i];
        Q.min_level = lvl - 1; Q.max_level = lvl; Q.valid = 1;
    }
    if ((rc = sbp_common(c, q, a->point_desc, a->point_observed, a->cur_occupied, a->cur_kps, a->cur_desc, a->cur_u_right, a->n_cur, a->cam, 1, 256, a->nn_ratio, assigned_cur))) return rc;
    int n = 0;
    for (int i = 0; i < a->n_points; ++i) n += assigned_cur[i] >= 0;
    *nmatches_out = n;
    return OLF_OK;
}

// ======================================================================================================
// grid-window search of Fuse / SearchByProjection(KF, Scw) / SearchBySim3 / SearchByProjection(Frame, KF)
// (src/ORBmatcher.cc:292-405, 827-1328, 1620-1747) on the candidate / resolve kernels above
// ======================================================================================================
int window_search(const olf_window_search_args* a, int* best_idx, int* best_dist, int device) {
    if (!a || !best_idx || !best_dist || a->n < 0 || a->n_queries < 0 || (a->n_queries > 0 && (!a->u || !a->v || !a->radius || !a->min_level || !a->max_level || !a->qdesc)) ||
        (a->chi2_check && (!a->u_right || !a->inv_level_sigma2 || !a->ur || a->nlevels < 1 || a->nlevels > OLF_MAX_LEVELS)) || a->max_dist < 0) {
        set_last_error("olf_window_search: bad arguments"); return OLF_ERR_ARG;
    }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    for (int i = 0; i < a->n_queries; ++i) { best_idx[i] = -1; best_dist[i] = 256; }
    if (a->n == 0 || a->n_queries == 0) return OLF_OK;
    if (!a->kps || !a->desc) { set_last_error("olf_window_search: bad arguments"); return OLF_ERR_ARG; }
    std::vector<SbpQuery> q(a->n_queries);
    std::vector<uint8_t> observed(a->n_queries, a->sequential_blocking ? 1 : 0);
    for (int i = 0; i < a->n_queries; ++i) {
        SbpQuery& Q = q[i];
        Q.u = a->u[i]; Q.v = a->v[i]; Q.radius = a->radius[i]; Q.ur = a->ur ? a->ur[i] : 0.f;
        // an octave window [min, max]; max < 0 would switch the level test off in the kernel (Frame::GetFeaturesInArea semantics), so an
        // empty window is expressed as an invalid query
        Q.min_level = a->min_level[i]; Q.max_level = a->max_level[i]; Q.valid = a->max_level[i] >= 0 && a->max_level[i] >= a->min_level[i];
    }
    return sbp_common(c, q, a->qdesc, observed.data(), a->blocked, a->kps, a->desc, a->u_right, a->n, a->cam, 0, std::min(a->max_dist, 256), 0.f, best_idx,
                      a->chi2_check ? 2 : 0, a->inv_level_sigma2, a->nlevels, best_dist);
}

static void rotation_filter(const std::vector<std::pair<int, float>>& rots, std::vector<int>& reject);
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:407-522): the device builds, per level-0 keypoint of F1, the list of level-0 keypoints of
// F2 in its window sorted by (distance, cell, index) -- the reference's preference order; the take-over rule (vMatchedDistance / vnMatches21) depends
// on the order of the keypoints of F1 and runs on the host over those lists
int search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2,
                              const olf_camera* cam, float* prev_matched, int window_size, float nn_ratio, int check_orientation,
                              int* m12, int* nmatches_out, int device) {
    if (!cam || !m12 || !nmatches_out || n1 < 0 || n2 < 0 || (n1 && (!kps1 || !desc1 || !prev_matched)) || (n2 && (!kps2 || !desc2)) || window_size < 0) {
        set_last_error("olf_search_for_initialization: bad arguments"); return OLF_ERR_ARG;
    }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    *nmatches_out = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    if (n1 == 0 || n2 == 0) return OLF_OK;
    if (n2 >= (1 << 24)) { set_last_error("olf_search_for_initialization: too many keypoints"); return OLF_ERR_CAPACITY; }
    std::vector<SbpQuery> q(n1);
    for (int i = 0; i < n1; ++i) {
        SbpQuery& Q = q[i];
        Q.u = prev_matched[2 * i]; Q.v = prev_matched[2 * i + 1]; Q.radius = (float)window_size; Q.ur = 0.f;
        Q.min_level = 0; Q.max_level = 0; Q.valid = kps1[i].octave <= 0;                     // `if (level1 > 0) continue`, GetFeaturesInArea(.., level1, level1)
    }
    std::vector<unsigned> csr;                                  // empty: first pass, SBP_K slots per window (see sbp_common)
    const u64* lists = nullptr; const int* cnt = nullptr;
    for (int pass = 0; pass < 2 && !lists; ++pass) {
        const bool wide = !csr.empty();
        const size_t slots = wide ? (size_t)csr.back() : (size_t)n1 * SBP_K;
        Planner pl;
        const size_t o_q = pl.d((size_t)n1 * sizeof(SbpQuery)), o_qd = pl.d((size_t)n1 * 32), o_k = pl.d((size_t)n2 * sizeof(olf_keypoint)), o_d = pl.d((size_t)n2 * 32),
                     o_lists = pl.d(slots * 8), o_cnt = pl.d((size_t)n1 * 4), o_stage = pl.d(wide ? slots * 8 : 0), o_off = pl.d(wide ? ((size_t)n1 + 1) * 4 : 0);
        const size_t p_in = pl.p((size_t)n1 * (sizeof(SbpQuery) + 32) + (size_t)n2 * (sizeof(olf_keypoint) + 32) + 96 + (wide ? ((size_t)n1 + 1) * 4 : 0)),
                     p_l = pl.p(slots * 8), p_c = pl.p((size_t)n1 * 4);
        if ((rc = arena_ensure(c, pl))) return rc;
        cudaStream_t s = c->cur;
        uint8_t* hp = hptr<uint8_t>(c, p_in);
        size_t off = 0;
        auto up = [&](size_t dev_off, const void* src, size_t bytes) -> int {
            memcpy(hp + off, src, bytes);
            cudaError_t e = cudaMemcpyAsync(c->a.dev.p + dev_off, hp + off, bytes, cudaMemcpyHostToDevice, s);
            off += (bytes + 15) & ~(size_t)15;
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__);
            return OLF_OK;
        };
        if ((rc = up(o_q, q.data(), (size_t)n1 * sizeof(SbpQuery))) || (rc = up(o_qd, desc1, (size_t)n1 * 32)) || (rc = up(o_k, kps2, (size_t)n2 * sizeof(olf_keypoint))) ||
            (rc = up(o_d, desc2, (size_t)n2 * 32)) || (wide && (rc = up(o_off, csr.data(), csr.size() * 4)))) return rc;
        GridParams G; G.min_x = cam->min_x; G.min_y = cam->min_y;
        G.inv_w = (float)OLF_GRID_COLS / (cam->max_x - cam->min_x); G.inv_h = (float)OLF_GRID_ROWS / (cam->max_y - cam->min_y);
        k_sbp_candidates<<<(n1 + 7) / 8, 256, 0, s>>>(dptr<SbpQuery>(c, o_q), dptr<uint32_t>(c, o_qd), n1, dptr<olf_keypoint>(c, o_k), dptr<uint32_t>(c, o_d), nullptr, n2,
                                                      G, 256, 0, nullptr, dptr<u64>(c, o_lists), dptr<int>(c, o_cnt),
                                                      wide ? dptr<unsigned>(c, o_off) : nullptr, wide ? dptr<u64>(c, o_stage) : nullptr);
        count_launches(1);
        if (slots) OLF_CUDA(cudaMemcpyAsync(hptr<u64>(c, p_l), dptr<u64>(c, o_lists), slots * 8, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_c), dptr<int>(c, o_cnt), (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
        OLF_CUDA(cudaGetLastError());
        OLF_CUDA(stream_sync(s));
        const int* hc = hptr<int>(c, p_c);
        bool overflow = false;
        for (int i1 = 0; i1 < n1 && !wide; ++i1) overflow |= hc[i1] > SBP_K;
        if (overflow) {
            if ((rc = csr_from_counts(hc, n1, csr, "olf_search_for_initialization"))) return rc;
            continue;
        }
        if (!wide) { csr.resize((size_t)n1 + 1); for (int i1 = 0; i1 <= n1; ++i1) csr[i1] = (unsigned)i1 * SBP_K; }
        lists = hptr<u64>(c, p_l); cnt = hc;
    }
    if (!lists) { set_last_error("olf_search_for_initialization: internal error"); return OLF_ERR_INTERNAL; }
    std::vector<int> vMatchedDistance(n2, INT_MAX), vnMatches21(n2, -1);
    std::vector<std::pair<int, float>> rots;
    int nmatches = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        // first and second admissible entry of the sorted list = bestDist / bestIdx2 and bestDist2 of the reference's loop (:443-458)
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int e = 0; e < cnt[i1]; ++e) {
            const u64 key = lists[(size_t)csr[i1] + e];
            const int i2 = (int)(key & 0xFFFFFF), dist = (int)(key >> 40);
            if (vMatchedDistance[i2] <= dist) continue;
            if (bestIdx2 < 0) { bestDist = dist; bestIdx2 = i2; } else { bestDist2 = dist; break; }
        }
        if (bestIdx2 >= 0 && bestDist <= OLF_TH_LOW && (float)bestDist < (float)bestDist2 * nn_ratio) {
            if (vnMatches21[bestIdx2] >= 0) { m12[vnMatches21[bestIdx2]] = -1; nmatches--; }
            m12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist;
            nmatches++;
            rots.push_back({i1, kps1[i1].angle - kps2[bestIdx2].angle});
        }
    }
    if (check_orientation) {
        std::vector<int> reject;
        rotation_filter(rots, reject);
        for (int i1 : reject) if (m12[i1] >= 0) { m12[i1] = -1; nmatches--; }
    }
    for (int i1 = 0; i1 < n1; ++i1) if (m12[i1] >= 0) { prev_matched[2 * i1] = kps2[m12[i1]].x; prev_matched[2 * i1 + 1] = kps2[m12[i1]].y; }
    *nmatches_out = nmatches;
    return OLF_OK;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:659-825): a warp per feature of KF1 (the features are independent: the
// reference never sets vbMatched2), the features of KF2 in the same vocabulary node over the lanes; minimum distance, the LAST
// position among equals (`dist > bestDist` rejects, so an equal later candidate replaces)
struct TriItem { int idx1, b2, e2; };
__global__ void __launch_bounds__(128) k_triangulate(const TriItem* __restrict__ items, int n_items, const olf_keypoint* __restrict__ kps1, const uint4* __restrict__ desc1,
                                                     const float* __restrict__ ur1, const olf_keypoint* __restrict__ kps2, const uint4* __restrict__ desc2,
                                                     const uint8_t* __restrict__ skip2, const float* __restrict__ ur2, const int* __restrict__ idx2s,
                                                     const float* __restrict__ sf2, const float* __restrict__ sig2, const float* __restrict__ F, float ex, float ey,
                                                     int only_stereo, int* __restrict__ m12) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_items) return;
    const TriItem it = items[w];
    const olf_keypoint kp1 = kps1[it.idx1];
    const bool stereo1 = ur1[it.idx1] >= 0;
    const uint4 a = desc1[2 * it.idx1], b = desc1[2 * it.idx1 + 1];
    // epipolar line of kp1 in the second image (:145-147)
    const float la = fadd(fadd(fmul(kp1.x, F[0]), fmul(kp1.y, F[3])), F[6]);
    const float lb = fadd(fadd(fmul(kp1.x, F[1]), fmul(kp1.y, F[4])), F[7]);
    const float lc = fadd(fadd(fmul(kp1.x, F[2]), fmul(kp1.y, F[5])), F[8]);
    const float den = fadd(fmul(la, la), fmul(lb, lb));
    int best = INT_MAX, pos = -1, bi = -1;
    for (int q = it.b2 + lane; q < it.e2; q += 32) {
        const int j = idx2s[q];
        if (skip2[j]) continue;
        const bool stereo2 = ur2[j] >= 0;
        if (only_stereo && !stereo2) continue;
        const uint4 x = desc2[2 * j], y = desc2[2 * j + 1];
        const int d = __popc(a.x ^ x.x) + __popc(a.y ^ x.y) + __popc(a.z ^ x.z) + __popc(a.w ^ x.w) +
                      __popc(b.x ^ y.x) + __popc(b.y ^ y.y) + __popc(b.z ^ y.z) + __popc(b.w ^ y.w);
        if (d > OLF_TH_LOW) continue;
        const olf_keypoint kp2 = kps2[j];
        if (!stereo1 && !stereo2) {
            const float dx = fsub(ex, kp2.x), dy = fsub(ey, kp2.y);
            if (fadd(fmul(dx, dx), fmul(dy, dy)) < fmul(100.f, sf2[kp2.octave])) continue;
        }
        if (den == 0) continue;
        const float num = fadd(fadd(fmul(la, kp2.x), fmul(lb, kp2.y)), lc);
        const float dsqr = fdiv(fmul(num, num), den);
        if (!((double)dsqr < __dmul_rn(3.84, (double)sig2[kp2.octave]))) continue;
        if (d < best || (d == best && q > pos)) { best = d; pos = q; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int eb = __shfl_xor_sync(0xffffffffu, best, o), ep = __shfl_xor_sync(0xffffffffu, pos, o), ei = __shfl_xor_sync(0xffffffffu, bi, o);
        if (eb < best || (eb == best && ep > pos)) { best = eb; pos = ep; bi = ei; }
    }
    if (lane == 0) m12[it.idx1] = bi;
}

static void rotation_filter(const std::vector<std::pair<int, float>>& rots, std::vector<int>& reject) {     // (entry, angle difference) -> entries of the pruned bins
    std::vector<int> rot[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    for (const auto& e : rots) {
        float r = e.second;
        if (r < 0.0) r += 360.0f;
        int bin = (int)roundf(r * factor);
        if (bin == OLF_HISTO_LENGTH) bin = 0;
        rot[bin].push_back(e.first);
    }
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;          // ComputeThreeMaxima (:1749-1790)
    for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
        const int sz = (int)rot[i].size();
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
        else if (sz > max3) { max3 = sz; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
    for (int i = 0; i < OLF_HISTO_LENGTH; i++)
        if (i != ind1 && i != ind2 && i != ind3) for (int e : rot[i]) reject.push_back(e);
}

int search_for_triangulation(const olf_triangulation_args* a, int* matches12, int* nmatches_out, int device) {
    if (!a || !matches12 || !nmatches_out || a->n1 < 0 || a->n2 < 0 || a->fv1_n_nodes < 0 || a->fv2_n_nodes < 0 || a->nlevels < 1 || a->nlevels > OLF_MAX_LEVELS) {
        set_last_error("olf_search_for_triangulation: bad arguments"); return OLF_ERR_ARG;
    }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    *nmatches_out = 0;
    for (int i = 0; i < a->n1; ++i) matches12[i] = -1;
    if (a->n1 == 0 || a->n2 == 0 || a->fv1_n_nodes == 0 || a->fv2_n_nodes == 0) return OLF_OK;
    std::vector<TriItem> items;
    for (int i = 0, j = 0; i < a->fv1_n_nodes && j < a->fv2_n_nodes;) {
        if (a->fv1_node[i] == a->fv2_node[j]) {
            for (int p = a->fv1_begin[i]; p < a->fv1_begin[i + 1]; ++p) {
                const int idx1 = a->fv1_index[p];
                if (a->skip1[idx1] || (a->only_stereo && !(a->u_right1[idx1] >= 0))) continue;
                items.push_back({idx1, a->fv2_begin[j], a->fv2_begin[j + 1]});
            }
            ++i; ++j;
        } else if (a->fv1_node[i] < a->fv2_node[j]) ++i;
        else ++j;
    }
    if (items.empty()) return OLF_OK;
    const int ni = (int)items.size(), n2i = a->fv2_begin[a->fv2_n_nodes];
    Planner pl;
    const size_t o_it = pl.d((size_t)ni * sizeof(TriItem)), o_k1 = pl.d((size_t)a->n1 * sizeof(olf_keypoint)), o_d1 = pl.d((size_t)a->n1 * 32), o_u1 = pl.d((size_t)a->n1 * 4),
                 o_k2 = pl.d((size_t)a->n2 * sizeof(olf_keypoint)), o_d2 = pl.d((size_t)a->n2 * 32), o_s2 = pl.d((size_t)a->n2), o_u2 = pl.d((size_t)a->n2 * 4),
                 o_i2 = pl.d((size_t)n2i * 4), o_sf = pl.d(OLF_MAX_LEVELS * 4), o_sg = pl.d(OLF_MAX_LEVELS * 4), o_F = pl.d(64), o_m = pl.d((size_t)a->n1 * 4);
    const size_t in_bytes = (size_t)ni * sizeof(TriItem) + (size_t)a->n1 * (sizeof(olf_keypoint) + 36) + (size_t)a->n2 * (sizeof(olf_keypoint) + 37) + (size_t)n2i * 4 + 2 * OLF_MAX_LEVELS * 4 + 64;
    const size_t p_in = pl.p(in_bytes + 256), p_m = pl.p((size_t)a->n1 * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    uint8_t* hp = hptr<uint8_t>(c, p_in);
    size_t off = 0;
    auto up = [&](size_t dev_off, const void* src, size_t bytes) -> int {
        if (!bytes) return OLF_OK;
        memcpy(hp + off, src, bytes);
        cudaError_t e = cudaMemcpyAsync(c->a.dev.p + dev_off, hp + off, bytes, cudaMemcpyHostToDevice, s);
        off += (bytes + 15) & ~(size_t)15;
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__);
        return OLF_OK;
    };
    if ((rc = up(o_it, items.data(), (size_t)ni * sizeof(TriItem))) || (rc = up(o_k1, a->kps1, (size_t)a->n1 * sizeof(olf_keypoint))) || (rc = up(o_d1, a->desc1, (size_t)a->n1 * 32)) ||
        (rc = up(o_u1, a->u_right1, (size_t)a->n1 * 4)) || (rc = up(o_k2, a->kps2, (size_t)a->n2 * sizeof(olf_keypoint))) || (rc = up(o_d2, a->desc2, (size_t)a->n2 * 32)) ||
        (rc = up(o_s2, a->skip2, (size_t)a->n2)) || (rc = up(o_u2, a->u_right2, (size_t)a->n2 * 4)) || (rc = up(o_i2, a->fv2_index, (size_t)n2i * 4)) ||
        (rc = up(o_sf, a->scale_factors2, (size_t)a->nlevels * 4)) || (rc = up(o_sg, a->level_sigma2_2, (size_t)a->nlevels * 4)) || (rc = up(o_F, a->F12, 36))) return rc;
    OLF_CUDA(cudaMemsetAsync(dptr<int>(c, o_m), 0xFF, (size_t)a->n1 * 4, s));
    k_triangulate<<<(ni * 32 + 127) / 128, 128, 0, s>>>(dptr<TriItem>(c, o_it), ni, dptr<olf_keypoint>(c, o_k1), dptr<uint4>(c, o_d1), dptr<float>(c, o_u1),
                                                        dptr<olf_keypoint>(c, o_k2), dptr<uint4>(c, o_d2), dptr<uint8_t>(c, o_s2), dptr<float>(c, o_u2), dptr<int>(c, o_i2),
                                                        dptr<float>(c, o_sf), dptr<float>(c, o_sg), dptr<float>(c, o_F), a->ex, a->ey, a->only_stereo, dptr<int>(c, o_m));
    count_launches(1);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_m), dptr<int>(c, o_m), (size_t)a->n1 * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(matches12, hptr<int>(c, p_m), (size_t)a->n1 * 4);
    int nm = 0;
    std::vector<std::pair<int, float>> rots;
    for (int i = 0; i < a->n1; ++i) if (matches12[i] >= 0) { ++nm; rots.push_back({i, a->kps1[i].angle - a->kps2[matches12[i]].angle}); }
    if (a->check_orientation) {
        std::vector<int> reject;
        rotation_filter(rots, reject);
        for (int i : reject) { matches12[i] = -1; --nm; }
    }
    *nmatches_out = nm;
    return OLF_OK;
}

// ======================================================================================================
// Bag of words: Frame::ComputeBoW (src/Frame.cc:585-597) -> TemplatedVocabulary::transform
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1260), and ORBmatcher::SearchByBoW (src/ORBmatcher.cc:161-290)
// ======================================================================================================
struct VocabImpl {
    int device = 0, k = 0, L = 0, n = 0;
    DevBuf<uint4> desc; DevBuf<int> child_begin, child_count, children, word_id; DevBuf<double> weight;
};
VocabImpl* vocab_create(const olf_vocab_desc* v, int device) {
    if (!v || v->n_nodes < 1 || !v->node_desc || !v->child_begin || !v->child_count || !v->children || !v->word_id || !v->weight || v->child_count[0] < 1) {
        set_last_error("olf_vocab_create: bad arguments"); return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError(); set_last_error("olf_vocab_create: no such CUDA device (this library has no CPU path)"); return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return nullptr; }
    int nch = 0;
    for (int i = 0; i < v->n_nodes; ++i) {
        if (v->child_count[i] < 0 || v->child_begin[i] < 0) { set_last_error("olf_vocab_create: bad tree"); return nullptr; }
        nch = std::max(nch, v->child_begin[i] + v->child_count[i]);
    }
    for (int i = 0; i < nch; ++i) if (v->children[i] <= 0 || v->children[i] >= v->n_nodes) { set_last_error("olf_vocab_create: bad tree"); return nullptr; }
    VocabImpl* h = new VocabImpl();
    h->device = device; h->k = v->k; h->L = v->L; h->n = v->n_nodes;
    const size_t n = (size_t)v->n_nodes;
    bool ok = h->desc.ensure(2 * n) == OLF_OK && h->child_begin.ensure(n) == OLF_OK && h->child_count.ensure(n) == OLF_OK &&
              h->children.ensure(std::max(nch, 1)) == OLF_OK && h->word_id.ensure(n) == OLF_OK && h->weight.ensure(n) == OLF_OK;
    ok = ok && cudaMemcpy(h->desc.p, v->node_desc, n * 32, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->child_begin.p, v->child_begin, n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->child_count.p, v->child_count, n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->children.p, v->children, (size_t)nch * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->word_id.p, v->word_id, n * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->weight.p, v->weight, n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { set_last_error(std::string("olf_vocab_create: ") + cudaGetErrorString(cudaGetLastError())); vocab_destroy(h); return nullptr; }
    return h;
}
void vocab_destroy(VocabImpl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->desc.release(); h->child_begin.release(); h->child_count.release(); h->children.release(); h->word_id.release(); h->weight.release();
    delete h;
}
// thread per descriptor: descend the tree, at every node the first child at minimal Hamming distance (strict <)
__global__ void __launch_bounds__(128) k_bow_transform(const uint4* __restrict__ feat, int n, const uint4* __restrict__ ndesc,
                                                       const int* __restrict__ child_begin, const int* __restrict__ child_count,
                                                       const int* __restrict__ children, const int* __restrict__ word_id,
                                                       const double* __restrict__ weight, int nid_level,
                                                       int* __restrict__ out_word, double* __restrict__ out_weight, int* __restrict__ out_node) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const uint4 a = feat[2 * f], b = feat[2 * f + 1];
    int nid = 0, final_id = 0, level = 0;
    do {
        ++level;
        const int cb = child_begin[final_id], cc = child_count[final_id];
        int best = INT_MAX, best_id = 0;
        for (int c = 0; c < cc; ++c) {
            const int id = children[cb + c];
            const uint4 x = ndesc[2 * (size_t)id], y = ndesc[2 * (size_t)id + 1];
            const int d = __popc(a.x ^ x.x) + __popc(a.y ^ x.y) + __popc(a.z ^ x.z) + __popc(a.w ^ x.w) +
                          __popc(b.x ^ y.x) + __popc(b.y ^ y.y) + __popc(b.z ^ y.z) + __popc(b.w ^ y.w);
            if (d < best) { best = d; best_id = id; }
        }
        final_id = best_id;
        if (level == nid_level) nid = final_id;
    } while (child_count[final_id] != 0);
    out_word[f] = word_id[final_id]; out_weight[f] = weight[final_id]; out_node[f] = nid;
}
int bow_transform(VocabImpl* V, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id) {
    if (!V || n < 0 || (n && (!desc || !word_id || !weight || !node_id))) { set_last_error("olf_bow_transform: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(V->device, &c))) return rc;
    if (n == 0) return OLF_OK;
    Planner pl;
    const size_t o_f = pl.d((size_t)n * 32), o_w = pl.d((size_t)n * 4), o_v = pl.d((size_t)n * 8), o_n = pl.d((size_t)n * 4);
    const size_t p_f = pl.p((size_t)n * 32), p_w = pl.p((size_t)n * 4), p_v = pl.p((size_t)n * 8), p_n = pl.p((size_t)n * 4);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    memcpy(hptr<uint8_t>(c, p_f), desc, (size_t)n * 32);
    OLF_CUDA(cudaMemcpyAsync(dptr<uint8_t>(c, o_f), hptr<uint8_t>(c, p_f), (size_t)n * 32, cudaMemcpyHostToDevice, s));
    k_bow_transform<<<(n + 127) / 128, 128, 0, s>>>(dptr<uint4>(c, o_f), n, V->desc.p, V->child_begin.p, V->child_count.p, V->children.p, V->word_id.p,
                                                    V->weight.p, V->L - levelsup, dptr<int>(c, o_w), dptr<double>(c, o_v), dptr<int>(c, o_n));
    count_launches(1);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_w), dptr<int>(c, o_w), (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaMemcpyAsync(hptr<double>(c, p_v), dptr<double>(c, o_v), (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_n), dptr<int>(c, o_n), (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(word_id, hptr<int>(c, p_w), (size_t)n * 4); memcpy(weight, hptr<double>(c, p_v), (size_t)n * 8); memcpy(node_id, hptr<int>(c, p_n), (size_t)n * 4);
    return OLF_OK;
}
// transform(features, v, fv, levelsup) with TF_IDF + L1_NORM (TemplatedVocabulary.h:1131-1194): the O(n) std::map bookkeeping,
// host side by design (double sums in feature order, words / nodes ascending)
int bow_assemble(const int* word_id, const double* weight, const int* node_id, int n, int* bow_word, double* bow_value, int* n_words,
                 int* fv_node, int* fv_begin, int* fv_index, int* n_nodes) {
    if (n < 0 || !n_words || !n_nodes || (n && (!word_id || !weight || !node_id || !bow_word || !bow_value || !fv_node || !fv_begin || !fv_index))) {
        set_last_error("olf_bow_assemble: bad arguments"); return OLF_ERR_ARG;
    }
    std::vector<int> ord;                                          // features that are not stopped, in feature order
    for (int i = 0; i < n; ++i) if (weight[i] > 0) ord.push_back(i);
    std::vector<int> by_word(ord), by_node(ord);
    std::stable_sort(by_word.begin(), by_word.end(), [&](int a, int b) { return (unsigned)word_id[a] < (unsigned)word_id[b]; });
    std::stable_sort(by_node.begin(), by_node.end(), [&](int a, int b) { return (unsigned)node_id[a] < (unsigned)node_id[b]; });
    int k = 0;
    for (size_t i = 0; i < by_word.size();) {                      // BowVector::addWeight: v[word] += w in feature order
        const int w = word_id[by_word[i]];
        double acc = weight[by_word[i]];
        size_t j = i + 1;
        for (; j < by_word.size() && word_id[by_word[j]] == w; ++j) acc += weight[by_word[j]];
        bow_word[k] = w; bow_value[k] = acc; ++k; i = j;
    }
    double norm = 0.0;                                             // BowVector::normalize(L1)
    for (int i = 0; i < k; ++i) norm += std::fabs(bow_value[i]);
    if (norm > 0.0) for (int i = 0; i < k; ++i) bow_value[i] /= norm;
    *n_words = k;
    int m = 0, pos = 0;
    for (size_t i = 0; i < by_node.size();) {                      // FeatureVector::addFeature
        const int nd = node_id[by_node[i]];
        fv_node[m] = nd; fv_begin[m] = pos;
        for (; i < by_node.size() && node_id[by_node[i]] == nd; ++i) fv_index[pos++] = by_node[i];
        ++m;
    }
    if (fv_begin) fv_begin[m] = pos;
    *n_nodes = m;
    return OLF_OK;
}

// warp per vocabulary node shared by the key frame and the frame: the key-frame features of the node in order (the
// "already matched" rule is loop-carried inside a node only: a frame feature belongs to one node), the frame features of
// the node over the lanes, (best, first index, second best) by shuffle
struct BowPair { int kf_b, kf_e, f_b, f_e; };
__global__ void __launch_bounds__(128) k_bow_match(const BowPair* __restrict__ pairs, int n_pairs, const uint4* __restrict__ kf_desc,
                                                   const uint8_t* __restrict__ kf_has, const int* __restrict__ kf_idx,
                                                   const uint4* __restrict__ f_desc, const int* __restrict__ f_idx, const uint8_t* __restrict__ f_has, int accept_below,
                                                   float nn_ratio, int* match_f) {
    // f_has / accept_below: SearchByBoW(KF, KF) needs a good map point on the second side too and accepts best < TH_LOW
    // (src/ORBmatcher.cc:580-585, 606); SearchByBoW(KF, Frame) accepts best <= TH_LOW (:234): accept_below = TH_LOW + 1
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const BowPair P = pairs[w];
    for (int p = P.kf_b; p < P.kf_e; ++p) {
        const int ikf = kf_idx[p];
        if (!kf_has[ikf]) continue;
        const uint4 a = kf_desc[2 * ikf], b = kf_desc[2 * ikf + 1];
        int d1 = 256, pos1 = INT_MAX, i1 = -1, d2 = 256;
        for (int q = P.f_b + lane; q < P.f_e; q += 32) {
            const int jf = f_idx[q];
            if (*(volatile int*)&match_f[jf] >= 0) continue;
            if (f_has && !f_has[jf]) continue;
            const uint4 x = f_desc[2 * jf], y = f_desc[2 * jf + 1];
            const int d = __popc(a.x ^ x.x) + __popc(a.y ^ x.y) + __popc(a.z ^ x.z) + __popc(a.w ^ x.w) +
                          __popc(b.x ^ y.x) + __popc(b.y ^ y.y) + __popc(b.z ^ y.z) + __popc(b.w ^ y.w);
            if (d < d1) { d2 = d1; d1 = d; pos1 = q; i1 = jf; }
            else if (d < d2) d2 = d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int e1 = __shfl_xor_sync(0xffffffffu, d1, o), ep = __shfl_xor_sync(0xffffffffu, pos1, o);
            const int ei = __shfl_xor_sync(0xffffffffu, i1, o), e2 = __shfl_xor_sync(0xffffffffu, d2, o);
            if (e1 < d1 || (e1 == d1 && ep < pos1)) { d2 = min(e2, d1); d1 = e1; pos1 = ep; i1 = ei; }
            else d2 = min(d2, e1);
        }
        if (lane == 0 && i1 >= 0 && d1 < accept_below && (float)d1 < fmul(nn_ratio, (float)d2)) match_f[i1] = ikf;
        __syncwarp();
    }
}
static int bow_match_common(const olf_bow_match_args* a, const uint8_t* f_has, int accept_below, int* match_f, int* nmatches_out, int device) {
    if (!a || !match_f || !nmatches_out || a->n_kf < 0 || a->n_f < 0) { set_last_error("olf_search_by_bow: bad arguments"); return OLF_ERR_ARG; }
    MatchCtx* c; int rc;
    if ((rc = get_ctx(device, &c))) return rc;
    *nmatches_out = 0;
    for (int i = 0; i < a->n_f; ++i) match_f[i] = -1;
    // the nodes both feature vectors contain (merge of two ascending lists = the lower_bound walk of :185-279)
    std::vector<BowPair> pairs;
    for (int ik = 0, jf = 0; ik < a->kf_n_nodes && jf < a->f_n_nodes;) {
        if (a->kf_fv_node[ik] == a->f_fv_node[jf]) { pairs.push_back({a->kf_fv_begin[ik], a->kf_fv_begin[ik + 1], a->f_fv_begin[jf], a->f_fv_begin[jf + 1]}); ++ik; ++jf; }
        else if (a->kf_fv_node[ik] < a->f_fv_node[jf]) ++ik;
        else ++jf;
    }
    if (pairs.empty() || a->n_f == 0 || a->n_kf == 0) return OLF_OK;
    const int nkf_idx = a->kf_fv_begin[a->kf_n_nodes], nf_idx = a->f_fv_begin[a->f_n_nodes], np = (int)pairs.size();
    Planner pl;
    const size_t o_kd = pl.d((size_t)a->n_kf * 32), o_kh = pl.d((size_t)a->n_kf), o_ki = pl.d((size_t)nkf_idx * 4), o_fd = pl.d((size_t)a->n_f * 32),
                 o_fi = pl.d((size_t)nf_idx * 4), o_p = pl.d((size_t)np * sizeof(BowPair)), o_m = pl.d((size_t)a->n_f * 4), o_fh = pl.d((size_t)a->n_f);
    const size_t p_kd = pl.p((size_t)a->n_kf * 32), p_kh = pl.p((size_t)a->n_kf), p_ki = pl.p((size_t)nkf_idx * 4), p_fd = pl.p((size_t)a->n_f * 32),
                 p_fi = pl.p((size_t)nf_idx * 4), p_p = pl.p((size_t)np * sizeof(BowPair)), p_m = pl.p((size_t)a->n_f * 4), p_fh = pl.p((size_t)a->n_f);
    if ((rc = arena_ensure(c, pl))) return rc;
    cudaStream_t s = c->cur;
    auto up = [&](size_t po, size_t dof, const void* src, size_t bytes) -> cudaError_t {
        memcpy(hptr<uint8_t>(c, po), src, bytes);
        return cudaMemcpyAsync(dptr<uint8_t>(c, dof), hptr<uint8_t>(c, po), bytes, cudaMemcpyHostToDevice, s);
    };
    OLF_CUDA(up(p_kd, o_kd, a->kf_desc, (size_t)a->n_kf * 32)); OLF_CUDA(up(p_kh, o_kh, a->kf_has_point, (size_t)a->n_kf));
    OLF_CUDA(up(p_ki, o_ki, a->kf_fv_index, (size_t)nkf_idx * 4)); OLF_CUDA(up(p_fd, o_fd, a->f_desc, (size_t)a->n_f * 32));
    OLF_CUDA(up(p_fi, o_fi, a->f_fv_index, (size_t)nf_idx * 4)); OLF_CUDA(up(p_p, o_p, pairs.data(), (size_t)np * sizeof(BowPair)));
    if (f_has) OLF_CUDA(up(p_fh, o_fh, f_has, (size_t)a->n_f));
    OLF_CUDA(cudaMemsetAsync(dptr<int>(c, o_m), 0xFF, (size_t)a->n_f * 4, s));
    k_bow_match<<<(np * 32 + 127) / 128, 128, 0, s>>>(dptr<BowPair>(c, o_p), np, dptr<uint4>(c, o_kd), dptr<uint8_t>(c, o_kh), dptr<int>(c, o_ki),
                                                      dptr<uint4>(c, o_fd), dptr<int>(c, o_fi), f_has ? dptr<uint8_t>(c, o_fh) : nullptr, accept_below,
                                                      a->nn_ratio, dptr<int>(c, o_m));
    count_launches(1);
    OLF_CUDA(cudaMemcpyAsync(hptr<int>(c, p_m), dptr<int>(c, o_m), (size_t)a->n_f * 4, cudaMemcpyDeviceToHost, s));
    OLF_CUDA(cudaGetLastError());
    OLF_CUDA(stream_sync(s));
    memcpy(match_f, hptr<int>(c, p_m), (size_t)a->n_f * 4);
    // rotation histogram (:248-258, 281-302): O(#matches) bookkeeping on the host, like the projection matchers
    int nm = 0;
    std::vector<int> rot[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    for (int j = 0; j < a->n_f; ++j) {
        if (match_f[j] < 0) continue;
        ++nm;
        if (a->check_orientation) {
            float r = a->kf_kps_un[match_f[j]].angle - a->f_kps[j].angle;
            if (r < 0.0) r += 360.0f;
            int bin = (int)roundf(r * factor);
            if (bin == OLF_HISTO_LENGTH) bin = 0;
            rot[bin].push_back(j);
        }
    }
    if (a->check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;          // ComputeThreeMaxima (:1749-1790)
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            const int sz = (int)rot[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rot[i]) { match_f[j] = -1; --nm; }
        }
    }
    *nmatches_out = nm;
    return OLF_OK;
}
int search_by_bow(const olf_bow_match_args* a, int* match_f, int* nmatches_out, int device) { return bow_match_common(a, nullptr, OLF_TH_LOW + 1, match_f, nmatches_out, device); }
// SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:524-657): a feature of KF2 is matched at most once, so the
// per-KF2 result inverts into the reference's per-KF1 vector
int search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2, int* matches12, int* nmatches_out, int device) {
    if (!a || !has_point2 || !matches12 || !nmatches_out || a->n_kf < 0 || a->n_f < 0) { set_last_error("olf_search_by_bow_kf: bad arguments"); return OLF_ERR_ARG; }
    std::vector<int> match_f(std::max(a->n_f, 1), -1);
    for (int i = 0; i < a->n_kf; ++i) matches12[i] = -1;
    const int rc = bow_match_common(a, has_point2, OLF_TH_LOW, match_f.data(), nmatches_out, device);
    if (rc) return rc;
    for (int j = 0; j < a->n_f; ++j) if (match_f[j] >= 0) matches12[match_f[j]] = j;
    return OLF_OK;
}


}  // namespace olf
