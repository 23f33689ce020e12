// ORB half of the front end on sm_100a: image pyramid, FAST-9/16 score map, per-cell threshold fallback + NMS +
// ordered compaction, quadtree distribution, intensity-centroid orientation, 7x7 fixed-point blur, 256-bit rBRIEF.
// Replaces ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc:412-472, 483-765, 767-855, 1045-1134).
// Arithmetic follows cv2 4.13 semantics pinned in SURVEY.md Appendix A (the reference's OpenCV is un-vendored).
//
// The whole extractor is ONE asynchronous chain of launches per call, for a BATCH of images (blockIdx.y = image): nothing
// returns to the host between the upload and the finished keypoints + descriptors.  The quadtree (DistributeOctTree,
// :541-765) is order-defining list surgery in the reference; here one CTA per (image, level) replays it with prefix sums:
// the list after a pass is [children of the divided nodes, newest first] ++ [undivided nodes in their old order], which is
// exactly what the reference's push_front / erase sequence produces.
#include "common.cuh"
#include "orb.h"
#include "img_kernels.cuh"
#include "sincosf_exact.h"
#define ORB_TMA_MAPS 64            // tensor maps per blur launch (8 images x 8 levels), 8 KB of kernel parameters
#include "../../include/olf_brief_pattern.h"
#include <algorithm>
#include <cmath>

namespace olf {

typedef unsigned long long u64;

// ---- per-image device pointers, a batch of them by value ---------------------------------------------------------
struct OrbDev {
    uint8_t *pyr, *score, *blur;
    int *cell_counts, *cell_thr, *cell_off, *total; unsigned* ticket;
    int4* cand; int cand_cap;
    int* cnode;                 // quadtree: node position of every candidate
    int4* nodes;                // two node lists (A, B) of qt_total entries each: {x0|y0<<16, x1|y1<<16, count, seq}
    int* qi;                    // 12 ints per node slot: child counts[4], child positions[4], rof, moved, proc, E
    u64* qk;                    // 3 u64 per node slot: expand keys (two lists), best candidate per node
    int* lvl_count;             // kept keypoints per level
    int* keep;                  // kept candidate index (into cand) per level, list order
    olf_keypoint* kps; uint8_t* desc; int* out;    // out[0] = n, out[1] = error flag
    int out_cap;
};
struct OrbBatch { int n; OrbDev d[IMG_MAX_BATCH]; };
struct QtTable { int N[OLF_MAX_LEVELS], node_off[OLF_MAX_LEVELS], node_cap[OLF_MAX_LEVELS]; int total_cap; float scale[OLF_MAX_LEVELS]; };

// ---- pyramid: cv::resize INTER_LINEAR 8UC1, one launch per level (level l is made from level l-1) ------
// coefficient tables (x: dst_w entries, y: dst_h entries) of {src index, c0, c1} are built on the host once per size.
struct LinCoef { int s; short c0, c1; };

__global__ void k_resize_linear(const __grid_constant__ OrbBatch B, unsigned soff, int sw, int sh, int spitch,
                                unsigned doff, int dw, int dh, int dpitch,
                                const LinCoef* __restrict__ cx, const LinCoef* __restrict__ cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const uint8_t* __restrict__ src = B.d[blockIdx.z].pyr + soff;
    uint8_t* __restrict__ dst = B.d[blockIdx.z].pyr + doff;
    const LinCoef a = cx[x], b = cy[y];
    const int x1 = min(a.s + 1, sw - 1), y1 = min(b.s + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)b.s * spitch;
    const uint8_t* r1 = src + (size_t)y1 * spitch;
    const int t0 = r0[a.s] * a.c0 + r0[x1] * a.c1;
    const int t1 = r1[a.s] * a.c0 + r1[x1] * a.c1;
    dst[(size_t)y * dpitch + x] = (uint8_t)((((b.c0 * (t0 >> 4)) >> 16) + ((b.c1 * (t1 >> 4)) >> 16) + 2) >> 2);
}

// ---- FAST-9/16 threshold-free score map, all levels in one launch --------------------------
// score = max over the 16 cyclic 9-arcs of min(d) / min(-d), minus 1; stored 0 when < min_th (never consulted then).
__global__ void __launch_bounds__(256) k_fast_score(const __grid_constant__ OrbBatch B, const __grid_constant__ LevelTable T, int min_th) {
    __shared__ uint8_t tile[TILE_H + 6][TILE_W + 8];
    int level, tx, ty;
    locate_tile(T, blockIdx.x, level, tx, ty);
    const int w = T.w[level], h = T.h[level], pitch = T.pitch[level];
    const uint8_t* img = B.d[blockIdx.y].pyr + T.off[level];
    uint8_t* out = B.d[blockIdx.y].score + T.off[level];
    const int x0 = tx * TILE_W, y0 = ty * TILE_H;
    for (int i = threadIdx.x; i < (TILE_H + 6) * (TILE_W + 6); i += 256) {
        const int r = i / (TILE_W + 6), c = i % (TILE_W + 6);
        const int gx = min(max(x0 + c - 3, 0), w - 1), gy = min(max(y0 + r - 3, 0), h - 1);
        tile[r][c] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    const int lx = threadIdx.x % TILE_W;
#pragma unroll 1
    for (int ly = threadIdx.x / TILE_W; ly < TILE_H; ly += 256 / TILE_W) {
        const int x = x0 + lx, y = y0 + ly;
        if (x >= w || y >= h) continue;
        int s = 0;
        if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
            const int cx = lx + 3, cy = ly + 3;
            const int v = tile[cy][cx];
            int d[16];
            d[0] = v - tile[cy + 3][cx];      d[1] = v - tile[cy + 3][cx + 1];  d[2] = v - tile[cy + 2][cx + 2];  d[3] = v - tile[cy + 1][cx + 3];
            d[4] = v - tile[cy][cx + 3];      d[5] = v - tile[cy - 1][cx + 3];  d[6] = v - tile[cy - 2][cx + 2];  d[7] = v - tile[cy - 3][cx + 1];
            d[8] = v - tile[cy - 3][cx];      d[9] = v - tile[cy - 3][cx - 1];  d[10] = v - tile[cy - 2][cx - 2]; d[11] = v - tile[cy - 1][cx - 3];
            d[12] = v - tile[cy][cx - 3];     d[13] = v - tile[cy + 1][cx - 3]; d[14] = v - tile[cy + 2][cx - 2]; d[15] = v - tile[cy + 3][cx - 1];
            unsigned mpos = 0, mneg = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) { mpos |= (d[k] > min_th) << k; mneg |= (d[k] < -min_th) << k; }
            if (has_run9(mpos) || has_run9(mneg)) {
                // exact score: sliding min/max over 9 via doubling (2,4,8,+1)
                int mn2[16], mx2[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
                int mn4[16], mx4[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
                int A = -255, Bm = -255;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
                    const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
                    A = max(A, mn9);
                    Bm = max(Bm, -mx9);
                }
                s = max(A, Bm) - 1;
                if (s < min_th) s = 0;
            }
        }
        out[(size_t)y * pitch + x] = (uint8_t)s;
    }
}

// ---- per-cell threshold fallback + NMS + ordered compaction (src/ORBextractor.cc:791-831) --------
// One warp per 30-px cell.  A pixel survives cv::FAST(cell window, th, nms) iff score >= th and score > every
// 8-neighbour inside the window's evaluated interior (threshold-independent), so one score map
// serves both thresholds; the cell uses ini_th unless that leaves it empty, then min_th.
struct CellGeom { int x_lo, x_hi, y_lo, y_hi; bool valid; };   // evaluated interior [x_lo,x_hi) x [y_lo,y_hi)
__device__ __forceinline__ CellGeom cell_geom(const LevelTable& T, int level, int ci, int cj) {
    CellGeom g;
    const int maxBX = T.w[level] - 16, maxBY = T.h[level] - 16;
    const int iniY = 16 + ci * T.hcell[level], iniX = 16 + cj * T.wcell[level];
    int maxY = iniY + T.hcell[level] + 6, maxX = iniX + T.wcell[level] + 6;
    g.valid = !(iniY >= maxBY - 3) && !(iniX >= maxBX - 6);
    if (maxY > maxBY) maxY = maxBY;
    if (maxX > maxBX) maxX = maxBX;
    g.x_lo = iniX + 3; g.x_hi = maxX - 3; g.y_lo = iniY + 3; g.y_hi = maxY - 3;
    if (g.x_hi - g.x_lo < 1 || g.y_hi - g.y_lo < 1) g.valid = false;      // window smaller than 7 px: FAST returns nothing
    return g;
}
__device__ __forceinline__ int nms_score(const uint8_t* sc, int pitch, const CellGeom& g, int x, int y) {
    // returns the pixel's score if it is a strict local maximum inside the interior, else 0
    const int s = sc[(size_t)y * pitch + x];
    if (s == 0) return 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            const int xx = x + dx, yy = y + dy;
            if (xx < g.x_lo || xx >= g.x_hi || yy < g.y_lo || yy >= g.y_hi) continue;
            if (sc[(size_t)yy * pitch + xx] >= s) return 0;
        }
    return s;
}
__device__ __forceinline__ void locate_cell(const LevelTable& T, int c, int& level, int& ci, int& cj) {
    level = 0;
#pragma unroll 1
    while (level + 1 < T.n && c >= T.cell_start[level + 1]) ++level;
    const int t = c - T.cell_start[level];
    cj = t % T.ncols[level];
    ci = t / T.ncols[level];
}

// counts[cell] = number of keypoints, thr[cell] = threshold used; the last block of an image scans its counts -> offsets.
__global__ void __launch_bounds__(128) k_cell_count(const __grid_constant__ OrbBatch B, const __grid_constant__ LevelTable T, int ini_th, int ncells) {
    const OrbDev& D = B.d[blockIdx.y];
    const uint8_t* __restrict__ score = D.score;
    int* __restrict__ counts = D.cell_counts;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + warp;
    if (c < ncells) {
        int level, ci, cj;
        locate_cell(T, c, level, ci, cj);
        const CellGeom g = cell_geom(T, level, ci, cj);
        int c_hi = 0, c_lo = 0;
        if (g.valid) {
            const uint8_t* sc = score + T.off[level];
            const int pitch = T.pitch[level];
            for (int y = g.y_lo; y < g.y_hi; ++y)
                for (int x = g.x_lo + lane; x < g.x_hi; x += 32) {
                    const int s = nms_score(sc, pitch, g, x, y);
                    c_lo += (s > 0);
                    c_hi += (s >= ini_th);
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { c_hi += __shfl_xor_sync(0xffffffffu, c_hi, o); c_lo += __shfl_xor_sync(0xffffffffu, c_lo, o); }
        if (lane == 0) { counts[c] = c_hi ? c_hi : c_lo; D.cell_thr[c] = c_hi ? ini_th : 1; }
    }
    // last-block-done exclusive scan
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(D.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ int part[128];
    const int per = (ncells + 127) / 128;
    const int b = threadIdx.x * per, e = min(b + per, ncells);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += ((volatile int*)counts)[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < 128; ++i) { const int v = part[i]; part[i] = acc; acc += v; }
        *D.total = acc;
        *D.ticket = 0;
    }
    __syncthreads();
    int acc = part[threadIdx.x];
    for (int i = b; i < e; ++i) { D.cell_off[i] = acc; acc += ((volatile int*)counts)[i]; }
}

// cand[k] = {level, x, y, score} in reference order: cells row-major, pixels row-major inside a cell.
__global__ void __launch_bounds__(128) k_cell_write(const __grid_constant__ OrbBatch B, const __grid_constant__ LevelTable T, int ncells) {
    const OrbDev& D = B.d[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + warp;
    if (c >= ncells) return;
    int level, ci, cj;
    locate_cell(T, c, level, ci, cj);
    const CellGeom g = cell_geom(T, level, ci, cj);
    if (!g.valid) return;
    const uint8_t* sc = D.score + T.off[level];
    const int pitch = T.pitch[level], th = D.cell_thr[c];
    int pos = D.cell_off[c];
    for (int y = g.y_lo; y < g.y_hi; ++y)
        for (int xb = g.x_lo; xb < g.x_hi; xb += 32) {
            const int x = xb + lane;
            int s = 0;
            if (x < g.x_hi) s = nms_score(sc, pitch, g, x, y);
            const bool keep = s >= th && s > 0;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int k = pos + __popc(m & ((1u << lane) - 1));
                if (k < D.cand_cap) D.cand[k] = make_int4(level, x, y, s);
            }
            pos += __popc(m);
        }
}

// ---- DistributeOctTree (src/ORBextractor.cc:483-765): one CTA per (image, level) ---------------------------------------
// Reference semantics restated as array operations.  A node list is an array in list order (index 0 = lNodes.front()).
// One "division step" divides the nodes proc[0..R] (in that order): their non-empty children n1..n4 are pushed to the
// front one after the other -- so the child created e-th ends up at position K-1-e -- and every other node keeps its
// relative order behind them.  Phase 1 (:598-669) divides every node with more than one key, in list order; phase 2
// (:677-741) divides the nodes created by the previous step in descending (size, creation sequence) order and stops
// as soon as the list has N nodes.  (The reference sorts pairs (size, ExtractorNode*): ties follow allocation
// addresses; the canonical choice is creation order)  A node's keys keep their original order in
// every child (stable partition), so "the first key of maximal response" (:747-760) is the candidate of maximal score
// with the smallest index: no per-node key list is needed, only the node position of every candidate.
__device__ __forceinline__ int qt_quadrant(int x, int y, const int4 nd) {
    const int x0 = nd.x & 0xffff, y0 = nd.x >> 16, x1 = nd.y & 0xffff, y1 = nd.y >> 16;
    const int hx = (x1 - x0 + 1) >> 1, hy = (y1 - y0 + 1) >> 1;              // ceil(static_cast<float>(d) / 2) (:485-486)
    return (x < x0 + hx ? 0 : 1) + (y < y0 + hy ? 0 : 2);                    // n1, n2, n3, n4 (:513-527)
}
__device__ __forceinline__ int4 qt_child(const int4 nd, int q, int cnt, int seq) {
    const int x0 = nd.x & 0xffff, y0 = nd.x >> 16, x1 = nd.y & 0xffff, y1 = nd.y >> 16;
    const int hx = (x1 - x0 + 1) >> 1, hy = (y1 - y0 + 1) >> 1;
    const int cx0 = (q & 1) ? x0 + hx : x0, cx1 = (q & 1) ? x1 : x0 + hx;
    const int cy0 = (q & 2) ? y0 + hy : y0, cy1 = (q & 2) ? y1 : y0 + hy;
    return make_int4(cx0 | (cy0 << 16), cx1 | (cy1 << 16), cnt, seq);
}
// exclusive prefix sum over the 1024 threads of the block; *total = sum.  sh: 33 ints.
__device__ __forceinline__ int block_scan_1024(int v, int* total, int* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                                   // sh may still be read from a previous scan
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = sh[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        sh[lane] = wi - w;
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    *total = sh[32];
    return sh[warp] + incl - v;
}
#define QT_KEY(cnt, seq, pos) (((u64)(unsigned)(cnt) << 40) | ((u64)(unsigned)(seq) << 16) | (u64)(unsigned)(pos))

extern __shared__ u64 s_qt_sort[];
__global__ void __launch_bounds__(1024) k_quadtree(const __grid_constant__ OrbBatch B, const __grid_constant__ LevelTable T,
                                                   const __grid_constant__ QtTable Q, int sort_cap) {
    __shared__ int s_scan[34];
    __shared__ int s_L, s_K, s_R, s_nE, s_seq, s_phase, s_done, s_nP, s_ER, s_kR;
    const OrbDev& D = B.d[blockIdx.y];
    const int level = blockIdx.x, tid = threadIdx.x;
    const int total_raw = *D.total;
    const int total = min(total_raw, D.cand_cap);
    if (total_raw > D.cand_cap && tid == 0) D.out[1] = OLF_ERR_CAPACITY;
    const int ncells = T.cell_start[T.n];
    const int cs = T.cell_start[level], ce = T.cell_start[level + 1];
    const int b = cs < ncells ? min(D.cell_off[cs], total) : total, e = ce < ncells ? min(D.cell_off[ce], total) : total;
    const int m = e - b, N = Q.N[level], C = Q.node_cap[level], O = Q.node_off[level];
    if (m <= 0) { if (tid == 0) D.lvl_count[level] = 0; return; }
    int4* A = D.nodes + O;
    int4* Bn = D.nodes + Q.total_cap + O;
    int* cc = D.qi + (size_t)12 * O; int* childpos = cc + 4 * C; int* rof = childpos + 4 * C; int* moved = rof + C; int* proc = moved + C; int* Epre = proc + C;
    u64* expA = D.qk + (size_t)3 * O; u64* expB = expA + C; u64* best = expB + C;
    int* cnode = D.cnode + b;
    const int4* cand = D.cand + b;
    // ---- initial nodes (:545-590)
    const int Wd = T.w[level] - 32, Hd = T.h[level] - 32;                     // maxX - minX, maxY - minY
    const int nIni = max((int)roundf(fdiv((float)Wd, (float)Hd)), 1);
    const float hX = fdiv((float)Wd, (float)nIni);
    for (int i = tid; i < nIni; i += 1024) cc[i] = 0;
    __syncthreads();
    for (int i = tid; i < m; i += 1024) {
        const int ni = min((int)fdiv((float)(cand[i].y - 16), hX), nIni - 1);
        cnode[i] = ni;
        atomicAdd(&cc[ni], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int L = 0;
        for (int i = 0; i < nIni && i < C; ++i) {
            if (cc[i] > 0) { moved[i] = L; A[L] = make_int4((int)fmul(hX, (float)i), (int)fmul(hX, (float)(i + 1)) | (Hd << 16), cc[i], i); ++L; }
            else moved[i] = -1;
        }
        s_L = L; s_seq = nIni; s_phase = 1; s_done = 0; s_nE = 0;
    }
    __syncthreads();
    for (int i = tid; i < m; i += 1024) cnode[i] = moved[cnode[i]];
    __syncthreads();
    // ---- division steps
    for (int iter = 0; iter < 8192; ++iter) {
        if (s_done) break;
        const int L = s_L, phase = s_phase, nE = s_nE;
        int nP;
        // a. nodes to divide, in processing order
        if (phase == 1) {
            const int per = (L + 1023) / 1024, pb = min(tid * per, L), pe = min(pb + per, L);
            int loc = 0;
            for (int p = pb; p < pe; ++p) loc += A[p].z >= 2;
            int run = block_scan_1024(loc, &nP, s_scan);
            for (int p = pb; p < pe; ++p) if (A[p].z >= 2) proc[run++] = p;
        } else {
            nP = nE;
            int n2 = 1; while (n2 < nE) n2 <<= 1;
            if (n2 > sort_cap) { if (tid == 0) { D.out[1] = OLF_ERR_CAPACITY; s_done = 1; } __syncthreads(); break; }
            for (int i = tid; i < n2; i += 1024) s_qt_sort[i] = i < nE ? expA[i] : 0ull;
            __syncthreads();
            for (int k = 2; k <= n2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < n2; i += 1024) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const u64 a = s_qt_sort[i], c2 = s_qt_sort[ixj];
                            const bool up = (i & k) == 0;
                            if ((a > c2) == up) { s_qt_sort[i] = c2; s_qt_sort[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int r = tid; r < nE; r += 1024) proc[r] = (int)(s_qt_sort[n2 - 1 - r] & 0xffffu);      // descending (size, seq) (:686-687)
        }
        if (tid == 0) { s_nP = nP; s_R = nP - 1; }
        __syncthreads();
        if (nP == 0) break;                                               // nothing to divide: the list cannot change any more
        // b. processing rank of every node, child counters
        for (int p = tid; p < L; p += 1024) rof[p] = 0;
        for (int j = tid; j < 4 * nP; j += 1024) cc[j] = 0;
        __syncthreads();
        for (int r = tid; r < nP; r += 1024) rof[proc[r]] = r + 1;
        __syncthreads();
        // c. children sizes (DivideNode :483-537)
        for (int i = tid; i < m; i += 1024) {
            const int p = cnode[i], r = rof[p];
            if (r) atomicAdd(&cc[4 * (r - 1) + qt_quadrant(cand[i].y - 16, cand[i].z - 16, A[p])], 1);
        }
        __syncthreads();
        // d. creation index of every child; phase 2: the step stops after the first division that reaches N nodes (:730-733)
        {
            const int per = (nP + 1023) / 1024, rb = min(tid * per, nP), re = min(rb + per, nP);
            int loc = 0;
            for (int r = rb; r < re; ++r) loc += (cc[4 * r] > 0) + (cc[4 * r + 1] > 0) + (cc[4 * r + 2] > 0) + (cc[4 * r + 3] > 0);
            int tot;
            int run = block_scan_1024(loc, &tot, s_scan);
            for (int r = rb; r < re; ++r) {
                const int k = (cc[4 * r] > 0) + (cc[4 * r + 1] > 0) + (cc[4 * r + 2] > 0) + (cc[4 * r + 3] > 0);
                Epre[r] = run; run += k;
                if (phase == 2 && L + run - (r + 1) >= N) atomicMin(&s_R, r);
            }
        }
        __syncthreads();
        const int R = s_R;
        if (tid == 0) { s_ER = Epre[R]; s_kR = (cc[4 * R] > 0) + (cc[4 * R + 1] > 0) + (cc[4 * R + 2] > 0) + (cc[4 * R + 3] > 0); }
        __syncthreads();
        const int K = s_ER + s_kR;
        // f. the nodes that are not divided keep their order behind the K new ones
        {
            const int per = (L + 1023) / 1024, pb = min(tid * per, L), pe = min(pb + per, L);
            int loc = 0;
            for (int p = pb; p < pe; ++p) loc += !(rof[p] && rof[p] - 1 <= R);
            int tot;
            int run = block_scan_1024(loc, &tot, s_scan);
            for (int p = pb; p < pe; ++p) if (!(rof[p] && rof[p] - 1 <= R)) { moved[p] = K + run; Bn[K + run] = A[p]; ++run; }
        }
        // g. children, newest first; i. those with more than one key are the next step's candidates, in creation order
        {
            const int seq0 = s_seq;
            const int per = (R + 1 + 1023) / 1024, rb = min(tid * per, R + 1), re = min(rb + per, R + 1);
            int loc = 0;
            for (int r = rb; r < re; ++r) loc += (cc[4 * r] > 1) + (cc[4 * r + 1] > 1) + (cc[4 * r + 2] > 1) + (cc[4 * r + 3] > 1);
            int nEn;
            int run = block_scan_1024(loc, &nEn, s_scan);
            for (int r = rb; r < re; ++r) {
                const int4 nd = A[proc[r]];
                int ecur = Epre[r];
                for (int q = 0; q < 4; ++q) {
                    const int cnt = cc[4 * r + q];
                    if (cnt <= 0) continue;
                    const int pos = K - 1 - ecur;
                    Bn[pos] = qt_child(nd, q, cnt, seq0 + ecur);
                    childpos[4 * r + q] = pos;
                    if (cnt > 1) expB[run++] = QT_KEY(cnt, seq0 + ecur, pos);
                    ++ecur;
                }
            }
            __syncthreads();
            // h. candidates follow their nodes
            for (int i = tid; i < m; i += 1024) {
                const int p = cnode[i], r = rof[p];
                cnode[i] = (r && r - 1 <= R) ? childpos[4 * (r - 1) + qt_quadrant(cand[i].y - 16, cand[i].z - 16, A[p])] : moved[p];
            }
            __syncthreads();
            if (tid == 0) {
                const int Ln = K + (L - (R + 1));
                s_seq = seq0 + K; s_L = Ln; s_nE = nEn;
                if (Ln >= N || Ln == L) s_done = 1;                                         // (:671-674, :736-737)
                else if (phase == 1 && Ln + 3 * nEn > N) s_phase = 2;                       // (:675)
            }
        }
        { int4* t = A; A = Bn; Bn = t; }
        { u64* t = expA; expA = expB; expB = t; }
        __syncthreads();
    }
    // ---- the best key of every node, in list order (:744-762)
    const int L = s_L;
    for (int p = tid; p < L; p += 1024) best[p] = 0ull;
    __syncthreads();
    for (int i = tid; i < m; i += 1024) atomicMax(&best[cnode[i]], ((u64)(unsigned)cand[i].w << 32) | (u64)(0xFFFFFFFFu - (unsigned)i));
    __syncthreads();
    for (int p = tid; p < L; p += 1024) D.keep[O + p] = b + (int)(0xFFFFFFFFu - (unsigned)(best[p] & 0xFFFFFFFFull));
    if (tid == 0) D.lvl_count[level] = L;
}

// ---- orientation + descriptor of the kept keypoints: one warp per keypoint -------------------------------------------
// IC_Angle (src/ORBextractor.cc:79-106): lanes = rows of the radius-15 disc, integer moments, cv::fastAtan2.
// computeOrbDescriptor (:110-149): lane i -> descriptor byte i; a = cosf(angle*pi/180), b = sinf(..) evaluated with glibc's
// own algorithm (sincosf_exact.h), single-rounded float products, cvRound.
__constant__ int c_umax[16];
__constant__ signed char c_pattern[1024];
__global__ void __launch_bounds__(256) k_orb_describe(const __grid_constant__ OrbBatch B, const __grid_constant__ LevelTable T, const __grid_constant__ QtTable Q) {
    const OrbDev& D = B.d[blockIdx.y];
    const int lane = threadIdx.x & 31;
    int start[OLF_MAX_LEVELS + 1];
    start[0] = 0;
    for (int l = 0; l < T.n; ++l) start[l + 1] = start[l] + D.lvl_count[l];
    const int n_all = start[T.n], n = min(n_all, D.out_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) { D.out[0] = n_all; if (n_all > D.out_cap) D.out[1] = OLF_ERR_CAPACITY; }
    for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < n; k += gridDim.x * 8) {
        int level = 0;
        while (k >= start[level + 1]) ++level;
        const int4 c = D.cand[D.keep[Q.node_off[level] + (k - start[level])]];
        const int pitch = T.pitch[level];
        const size_t center_off = T.off[level] + (size_t)c.z * pitch + c.y;
        int m10 = 0, m01 = 0;
        if (lane < 31) {
            const uint8_t* center = D.pyr + center_off;
            const int v = lane - 15;
            const int d = c_umax[v < 0 ? -v : v];
            const uint8_t* row = center + v * pitch;
            int rs = 0;
            for (int u = -d; u <= d; ++u) { const int p = row[u]; m10 += u * p; rs += p; }
            m01 = v * rs;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        if (lane == 0) {
            olf_keypoint kp;
            kp.x = (float)c.y; kp.y = (float)c.z;
            if (level != 0) { kp.x = fmul(kp.x, Q.scale[level]); kp.y = fmul(kp.y, Q.scale[level]); }      // (:1097-1103)
            kp.size = (float)(int)fmul(31.f, Q.scale[level]);                                               // PATCH_SIZE*mvScaleFactor[level] (:839)
            kp.angle = angle; kp.response = (float)c.w; kp.octave = level;
            D.kps[k] = kp;
        }
        const float ang = fmul(angle, 0x1.1df46ap-6f);                     // (float)(CV_PI/180.f)
        const float a = trig::cosf_exact(ang), bsn = trig::sinf_exact(ang);
        const uint8_t* center = D.blur + center_off;
        int val = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const signed char* p = c_pattern + (lane * 8 + t) * 4;
            const float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            const int r0 = __float2int_rn(fadd(fmul(x0, bsn), fmul(y0, a))), c0 = __float2int_rn(fsub(fmul(x0, a), fmul(y0, bsn)));
            const int r1 = __float2int_rn(fadd(fmul(x1, bsn), fmul(y1, a))), c1 = __float2int_rn(fsub(fmul(x1, a), fmul(y1, bsn)));
            const int t0 = center[r0 * pitch + c0], t1 = center[r1 * pitch + c1];
            val |= (t0 < t1) << t;
        }
        D.desc[(size_t)k * 32 + lane] = (uint8_t)val;
    }
}
// device trig sweep for the parity test: out[i] = {cosf_exact(x_i), sinf_exact(x_i)} for the floats with bit patterns first + i*stride
__global__ void k_trig_sweep(unsigned first, unsigned stride, unsigned count, float2* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float x = __uint_as_float(first + i * stride);
    out[i] = make_float2(trig::cosf_exact(x), trig::sinf_exact(x));
}

// ======================================================================================================
// host side
// ======================================================================================================
struct OrbImpl {
    int device = 0;
    int nfeatures, nlevels, ini_th, min_th;
    float scale_factor;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> feats_per_level;
    int umax[16];
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    SyncEvent done;
    // size-dependent state
    int img_w = 0, img_h = 0;
    LevelTable T;
    QtTable Q;
    size_t pyr_bytes = 0;
    int ncells = 0, ntiles = 0, sort_cap = 0;
    DevBuf<uint8_t> pyr, score, blur;
    DevBuf<LinCoef> coef;                    // per level: cx then cy
    std::vector<size_t> coef_off_x, coef_off_y;
    DevBuf<int> cell_counts, cell_thr, cell_off, total, cnode, qi, lvl_count, keep, out;
    DevBuf<unsigned> ticket;
    DevBuf<int4> cand, nodes;
    DevBuf<u64> qk;
    int cand_cap = 0, out_cap = 0;
    DevBuf<olf_keypoint> kps;
    DevBuf<uint8_t> desc;
    PinBuf<olf_keypoint> kps_host;
    PinBuf<uint8_t> desc_host;
    PinBuf<int> out_host;
    PinBuf<uint8_t> img_stage;               // pinned staging for pageable host images
    CUtensorMap tmaps[OLF_MAX_LEVELS]; bool has_tma = false;      // one tensor map per pyramid level (TMA tile staging of the 7x7 blur), passed by value
    int copy_cap = 0;                        // entries copied back by the last enqueue
    int last_ncand = -1;
};

static void build_lin_coefs(int src, int dst, std::vector<LinCoef>& out) {
    const double inv_scale = (double)dst / src;
    const double scale = 1.0 / inv_scale;
    out.resize(dst);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= src - 1) { f = 0; s = src - 1; }
        out[d].s = s;
        out[d].c0 = (short)lrintf((1.f - f) * 2048.f);
        out[d].c1 = (short)lrintf(f * 2048.f);
    }
}

static int orb_ensure_size(OrbImpl* h, int w, int hgt) {
    if (h->img_w == w && h->img_h == hgt) return OLF_OK;
    LevelTable& T = h->T;
    memset(&T, 0, sizeof(T));
    T.n = h->nlevels;
    size_t off = 0;
    int tiles = 0, cells = 0;
    for (int l = 0; l < h->nlevels; ++l) {
        T.w[l] = (int)lrintf((float)w * h->inv_scale[l]);            // ComputePyramid :1113-1114
        T.h[l] = (int)lrintf((float)hgt * h->inv_scale[l]);
        if (T.w[l] < 1 || T.h[l] < 1) { set_last_error("image too small for the pyramid"); return OLF_ERR_ARG; }
        T.pitch[l] = align_up(T.w[l], 64);
        T.off[l] = (unsigned)off;
        off += (size_t)T.pitch[l] * T.h[l];
        T.tiles_x[l] = (T.w[l] + TILE_W - 1) / TILE_W;
        T.tile_start[l] = tiles;
        tiles += T.tiles_x[l] * ((T.h[l] + TILE_H - 1) / TILE_H);
        // cell grid (:783-789)
        const float width = (float)(T.w[l] - 32), height = (float)(T.h[l] - 32);
        T.cell_start[l] = cells;
        if (width >= 30.f && height >= 30.f) {
            T.ncols[l] = (int)(width / 30.f); T.nrows[l] = (int)(height / 30.f);
            T.wcell[l] = (int)ceilf(width / T.ncols[l]); T.hcell[l] = (int)ceilf(height / T.nrows[l]);
            cells += T.ncols[l] * T.nrows[l];
        } else { T.ncols[l] = 1; T.nrows[l] = 0; T.wcell[l] = 1; T.hcell[l] = 1; }
    }
    T.tile_start[h->nlevels] = tiles;
    T.cell_start[h->nlevels] = cells;
    h->pyr_bytes = off; h->ntiles = tiles; h->ncells = cells;
    int rc;
    if ((rc = h->pyr.ensure(off + 256)) || (rc = h->score.ensure(off + 256)) || (rc = h->blur.ensure(off + 256))) return rc;
    // resize coefficient tables
    std::vector<LinCoef> all, tmp;
    h->coef_off_x.assign(h->nlevels, 0); h->coef_off_y.assign(h->nlevels, 0);
    for (int l = 1; l < h->nlevels; ++l) {
        build_lin_coefs(T.w[l - 1], T.w[l], tmp); h->coef_off_x[l] = all.size(); all.insert(all.end(), tmp.begin(), tmp.end());
        build_lin_coefs(T.h[l - 1], T.h[l], tmp); h->coef_off_y[l] = all.size(); all.insert(all.end(), tmp.begin(), tmp.end());
    }
    if ((rc = h->coef.ensure(std::max<size_t>(all.size(), 1)))) return rc;
    if (!all.empty()) OLF_CUDA(cudaMemcpy(h->coef.p, all.data(), all.size() * sizeof(LinCoef), cudaMemcpyHostToDevice));
    const int nc = std::max(cells, 1);
    if ((rc = h->cell_counts.ensure(nc)) || (rc = h->cell_thr.ensure(nc)) || (rc = h->cell_off.ensure(nc)) ||
        (rc = h->total.ensure(1)) || (rc = h->ticket.ensure(1)) || (rc = h->out.ensure(4)) || (rc = h->out_host.ensure(4)) || (rc = h->lvl_count.ensure(OLF_MAX_LEVELS))) return rc;
    OLF_CUDA(cudaMemset(h->ticket.p, 0, sizeof(unsigned)));
    // at most one NMS survivor per 2x2 block of the evaluated area
    h->cand_cap = (int)std::min<size_t>(off / 4 + 1024, (size_t)1 << 22);
    if ((rc = h->cand.ensure(h->cand_cap)) || (rc = h->cnode.ensure(h->cand_cap))) return rc;
    // quadtree node slots: a phase-1 pass can leave up to 4 x (N - 1) nodes
    QtTable& Q = h->Q;
    memset(&Q, 0, sizeof(Q));
    int qoff = 0, maxcap = 0;
    for (int l = 0; l < h->nlevels; ++l) {
        Q.N[l] = h->feats_per_level[l]; Q.node_cap[l] = 4 * h->feats_per_level[l] + 64; Q.node_off[l] = qoff; Q.scale[l] = h->scale[l];
        qoff += Q.node_cap[l]; maxcap = std::max(maxcap, Q.node_cap[l]);
    }
    Q.total_cap = qoff;
    if (maxcap >= 65536) { set_last_error("olf_orb: too many features per level for the device quadtree (< 16368 per level)"); return OLF_ERR_ARG; }
    h->sort_cap = 1; while (h->sort_cap < maxcap) h->sort_cap <<= 1;
    if (h->sort_cap * sizeof(u64) > 48 * 1024)
        OLF_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->sort_cap * sizeof(u64))));
    h->out_cap = qoff;
    if ((rc = h->nodes.ensure((size_t)2 * qoff)) || (rc = h->qi.ensure((size_t)12 * qoff)) || (rc = h->qk.ensure((size_t)3 * qoff)) || (rc = h->keep.ensure(qoff)) ||
        (rc = h->kps.ensure(qoff)) || (rc = h->desc.ensure((size_t)qoff * 32)) || (rc = h->kps_host.ensure(qoff)) || (rc = h->desc_host.ensure((size_t)qoff * 32))) return rc;
    if ((rc = h->img_stage.ensure((size_t)w * hgt))) return rc;
    h->has_tma = build_level_maps(h->pyr.p, T, 7, h->tmaps);
    h->img_w = w; h->img_h = hgt;
    return OLF_OK;
}

OrbImpl* orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device, cudaStream_t ext_stream) {
    if (nlevels < 1 || nlevels > OLF_MAX_LEVELS || nfeatures < 0 || scale_factor <= 1.0f || min_th < 1 || ini_th < min_th) {
        set_last_error("olf_orb_create: bad arguments"); return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_last_error("olf_orb_create: no such CUDA device (this library has no CPU path)"); return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return nullptr; }
    OrbImpl* h = new OrbImpl();
    h->device = device; h->nfeatures = nfeatures; h->nlevels = nlevels; h->ini_th = ini_th; h->min_th = min_th; h->scale_factor = scale_factor;
    // ORBextractor ctor (:417-448)
    h->scale.resize(nlevels); h->sigma2.resize(nlevels); h->inv_scale.resize(nlevels); h->inv_sigma2.resize(nlevels);
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { h->scale[i] = h->scale[i - 1] * scale_factor; h->sigma2[i] = h->scale[i] * h->scale[i]; }
    for (int i = 0; i < nlevels; i++) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    h->feats_per_level.resize(nlevels);
    const float factor = 1.0f / scale_factor;
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) { h->feats_per_level[l] = (int)lrintf(nd); sum += h->feats_per_level[l]; nd *= factor; }
    h->feats_per_level[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax (:456-471)
    const int HP = 15;
    int v, v0, vmax = (int)floorf(HP * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(HP * sqrtf(2.f) / 2);
    for (v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt((double)HP * HP - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) { while (h->umax[v0] == h->umax[v0 + 1]) ++v0; h->umax[v] = v0; ++v0; }
    if (ext_stream) { h->stream = ext_stream; h->owns_stream = false; }
    if ((!ext_stream && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) ||
        h->done.create(false) != cudaSuccess ||
        cudaMemcpyToSymbol(c_umax, h->umax, sizeof(h->umax)) != cudaSuccess ||
        cudaMemcpyToSymbol(c_pattern, OLF_BRIEF_PATTERN, 1024) != cudaSuccess) {
        set_last_error(std::string("olf_orb_create: ") + cudaGetErrorString(cudaGetLastError()));
        delete h; return nullptr;
    }
    return h;
}

cudaStream_t orb_stream(const OrbImpl* h) { return h ? h->stream : nullptr; }
void orb_destroy(OrbImpl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream && h->owns_stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->done.destroy();
    h->pyr.release(); h->score.release(); h->blur.release(); h->coef.release(); h->cell_counts.release(); h->cell_thr.release();
    h->cell_off.release(); h->total.release(); h->ticket.release(); h->cand.release(); h->cnode.release(); h->nodes.release(); h->qi.release();
    h->qk.release(); h->lvl_count.release(); h->keep.release(); h->out.release(); h->kps.release(); h->desc.release();
    h->kps_host.release(); h->desc_host.release(); h->out_host.release(); h->img_stage.release();
    delete h;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// The whole extractor for n images (one handle each, same parameters and image size), asynchronously on stream s:
// upload -> pyramid -> FAST score -> cells -> quadtree -> orientation + rBRIEF -> results to pinned staging.
int orb_enqueue(OrbImpl* const* hs, int n, const uint8_t* const* imgs, int w, int hgt, int stride, bool on_device, int cap, cudaStream_t s) {
    if (!hs || n < 1 || n > IMG_MAX_BATCH || !imgs || w <= 0 || hgt <= 0 || stride < w || cap < 0) { set_last_error("olf_orb_extract: bad arguments"); return OLF_ERR_ARG; }
    OrbImpl* h0 = hs[0];
    OLF_CUDA(cudaSetDevice(h0->device));
    int rc;
    OrbBatch B; memset(&B, 0, sizeof(B));
    BlurBatch BB; memset(&BB, 0, sizeof(BB));
    B.n = n;
    for (int k = 0; k < n; ++k) {
        OrbImpl* h = hs[k];
        if (!h || !imgs[k] || h->nfeatures != h0->nfeatures || h->nlevels != h0->nlevels || h->scale_factor != h0->scale_factor ||
            h->ini_th != h0->ini_th || h->min_th != h0->min_th || h->device != h0->device) { set_last_error("olf_orb_extract: the extractors of a batch must be configured alike"); return OLF_ERR_ARG; }
        if ((rc = orb_ensure_size(h, w, hgt))) return rc;
        const LevelTable& T = h->T;
        if (on_device) OLF_CUDA(cudaMemcpy2DAsync(h->pyr.p, T.pitch[0], imgs[k], stride, w, hgt, cudaMemcpyDeviceToDevice, s));
        else if (is_pinned_host(imgs[k])) OLF_CUDA(cudaMemcpy2DAsync(h->pyr.p, T.pitch[0], imgs[k], stride, w, hgt, cudaMemcpyHostToDevice, s));
        else {
            // pageable caller memory: stage through the handle's pinned buffer.  The previous use of the staging buffer must have
            // been consumed by the device: every entry point waits for its chain before it returns.
            for (int y = 0; y < hgt; ++y) memcpy(h->img_stage.p + (size_t)y * w, imgs[k] + (size_t)y * stride, w);
            OLF_CUDA(cudaMemcpy2DAsync(h->pyr.p, T.pitch[0], h->img_stage.p, w, w, hgt, cudaMemcpyHostToDevice, s));
        }
        OrbDev& D = B.d[k];
        D.pyr = h->pyr.p; D.score = h->score.p; D.blur = h->blur.p;
        D.cell_counts = h->cell_counts.p; D.cell_thr = h->cell_thr.p; D.cell_off = h->cell_off.p; D.total = h->total.p; D.ticket = h->ticket.p;
        D.cand = h->cand.p; D.cand_cap = h->cand_cap; D.cnode = h->cnode.p; D.nodes = h->nodes.p; D.qi = h->qi.p; D.qk = h->qk.p;
        D.lvl_count = h->lvl_count.p; D.keep = h->keep.p; D.kps = h->kps.p; D.desc = h->desc.p; D.out = h->out.p; D.out_cap = h->out_cap;
        BB.src[k] = h->pyr.p; BB.dst[k] = h->blur.p;
        OLF_CUDA(cudaMemsetAsync(h->out.p, 0, 4 * sizeof(int), s));
        if (h->ncells == 0) { OLF_CUDA(cudaMemsetAsync(h->total.p, 0, sizeof(int), s)); OLF_CUDA(cudaMemsetAsync(h->lvl_count.p, 0, OLF_MAX_LEVELS * sizeof(int), s)); }
        h->last_ncand = -1;
    }
    const LevelTable& T = h0->T;
    for (int l = 1; l < h0->nlevels; ++l) {
        dim3 b(32, 8), g((T.w[l] + 31) / 32, (T.h[l] + 7) / 8, n);
        k_resize_linear<<<g, b, 0, s>>>(B, T.off[l - 1], T.w[l - 1], T.h[l - 1], T.pitch[l - 1], T.off[l], T.w[l], T.h[l], T.pitch[l],
                                        h0->coef.p + h0->coef_off_x[l], h0->coef.p + h0->coef_off_y[l]);
    }
    k_fast_score<<<dim3(h0->ntiles, n), 256, 0, s>>>(B, T, h0->min_th);
    // ORB blur: 7x7 sigma 2 -> [18,34,48,56,48,34,18] (:1088)
    bool all_tma = n * T.n <= ORB_TMA_MAPS;
    for (int k = 0; k < n; ++k) all_tma = all_tma && hs[k]->has_tma;
    if (all_tma) {
        static thread_local TmaSet<ORB_TMA_MAPS> M;
        for (int k = 0; k < n; ++k) memcpy(&M.m[k * T.n], hs[k]->tmaps, sizeof(CUtensorMap) * T.n);
        k_blur_q8_tma<7, ORB_TMA_MAPS><<<dim3(h0->ntiles, n), 256, 0, s>>>(BB, M, T, 18, 34, 48, 56);
    }
    else k_blur_q8<7><<<dim3(h0->ntiles, n), 256, 0, s>>>(BB, T, 18, 34, 48, 56);
    int launches = (h0->nlevels - 1) + 2;
    if (h0->ncells > 0) {
        const int nb = (h0->ncells + 3) / 4;
        k_cell_count<<<dim3(nb, n), 128, 0, s>>>(B, T, h0->ini_th, h0->ncells);
        k_cell_write<<<dim3(nb, n), 128, 0, s>>>(B, T, h0->ncells);
        k_quadtree<<<dim3(h0->nlevels, n), 1024, h0->sort_cap * sizeof(u64), s>>>(B, T, h0->Q, h0->sort_cap);
        launches += 3;
    }
    k_orb_describe<<<dim3(64, n), 256, 0, s>>>(B, T, h0->Q);
    count_launches(launches + 1);
    for (int k = 0; k < n; ++k) {
        OrbImpl* h = hs[k];
        h->copy_cap = std::min(cap, h->out_cap);
        OLF_CUDA(cudaMemcpyAsync(h->out_host.p, h->out.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (h->copy_cap > 0) {
            OLF_CUDA(cudaMemcpyAsync(h->kps_host.p, h->kps.p, (size_t)h->copy_cap * sizeof(olf_keypoint), cudaMemcpyDeviceToHost, s));
            OLF_CUDA(cudaMemcpyAsync(h->desc_host.p, h->desc.p, (size_t)h->copy_cap * 32, cudaMemcpyDeviceToHost, s));
        }
    }
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}
// after the stream of orb_enqueue has been waited for: results of one image from its pinned staging
int orb_collect(OrbImpl* h, olf_keypoint* kps, uint8_t* desc, int cap, int* n) {
    if (!h || !n) return OLF_ERR_ARG;
    *n = 0;
    if (h->out_host.p[1] != 0) { set_last_error("olf_orb_extract: internal candidate / keypoint buffer overflow"); return OLF_ERR_CAPACITY; }
    const int total = h->out_host.p[0];
    if (total > cap || total > h->copy_cap) { set_last_error("olf_orb_extract: keypoint capacity too small"); return OLF_ERR_CAPACITY; }
    *n = total;
    if (total) { memcpy(kps, h->kps_host.p, (size_t)total * sizeof(olf_keypoint)); memcpy(desc, h->desc_host.p, (size_t)total * 32); }
    return OLF_OK;
}

int orb_extract(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device, olf_keypoint* kps, uint8_t* desc, int cap, int* n) {
    if (!h || !n) return OLF_ERR_ARG;
    *n = 0;
    if (!img || w <= 0 || hgt <= 0) return OLF_OK;           // _image.empty(): silent return (:1048)
    if (stride < w || !kps || !desc) { set_last_error("olf_orb_extract: bad arguments"); return OLF_ERR_ARG; }
    int rc = orb_enqueue(&h, 1, &img, w, hgt, stride, on_device, cap, h->stream);
    if (rc) return rc;
    OLF_CUDA(h->done.sync(h->stream));
    return orb_collect(h, kps, desc, cap, n);
}

int orb_level_size(const OrbImpl* h, int level, int* w, int* hh) {
    if (!h || level < 0 || level >= h->nlevels || h->img_w == 0) return OLF_ERR_ARG;
    *w = h->T.w[level]; *hh = h->T.h[level];
    return OLF_OK;
}
int orb_get_level(OrbImpl* h, int level, uint8_t* dst, int dst_stride) {
    if (!h || level < 0 || level >= h->nlevels || h->img_w == 0 || !dst) return OLF_ERR_ARG;
    OLF_CUDA(cudaSetDevice(h->device));
    OLF_CUDA(cudaMemcpy2D(dst, dst_stride, h->pyr.p + h->T.off[level], h->T.pitch[level], h->T.w[level], h->T.h[level], cudaMemcpyDeviceToHost));
    return OLF_OK;
}
// debug / parity: the FAST candidates of the last extract stay on the device; they are fetched on demand
int orb_last_candidates(OrbImpl* h, int* out, int cap, int* n) {
    if (!h || !n || h->img_w == 0) return OLF_ERR_ARG;
    OLF_CUDA(cudaSetDevice(h->device));
    int total = 0;
    OLF_CUDA(cudaMemcpy(&total, h->total.p, sizeof(int), cudaMemcpyDeviceToHost));
    *n = total;
    if (total > cap || total > h->cand_cap) return OLF_ERR_CAPACITY;
    if (total) OLF_CUDA(cudaMemcpy(out, h->cand.p, (size_t)total * sizeof(int4), cudaMemcpyDeviceToHost));
    return OLF_OK;
}
const OrbDeviceView orb_device_view(const OrbImpl* h) {
    OrbDeviceView v;
    v.pyr = h->pyr.p; v.nlevels = h->nlevels;
    for (int l = 0; l < h->nlevels; ++l) { v.w[l] = h->T.w[l]; v.h[l] = h->T.h[l]; v.pitch[l] = h->T.pitch[l]; v.off[l] = h->T.off[l]; v.scale[l] = h->scale[l]; v.inv_scale[l] = h->inv_scale[l]; }
    v.device = h->device;
    v.kps = h->kps.p; v.desc = h->desc.p; v.n = h->out.p; v.cap = h->out_cap;
    return v;
}
void orb_scale_tables(const OrbImpl* h, const float** s, const float** is, const float** s2, const float** is2, const int** fpl, int* nlevels) {
    if (s) *s = h->scale.data(); if (is) *is = h->inv_scale.data(); if (s2) *s2 = h->sigma2.data(); if (is2) *is2 = h->inv_sigma2.data();
    if (fpl) *fpl = h->feats_per_level.data(); if (nlevels) *nlevels = h->nlevels;
}
int orb_trig_sweep(unsigned first, unsigned stride, unsigned count, float* cos_sin_out, int device) {
    OLF_CUDA(cudaSetDevice(device));
    float2* d = nullptr;
    OLF_CUDA(cudaMalloc((void**)&d, (size_t)count * sizeof(float2)));
    k_trig_sweep<<<(count + 255) / 256, 256>>>(first, stride, count, d);
    cudaError_t e = cudaMemcpy(cos_sin_out, d, (size_t)count * sizeof(float2), cudaMemcpyDeviceToHost);
    cudaFree(d);
    OLF_CUDA(e);
    return OLF_OK;
}


}  // namespace olf
