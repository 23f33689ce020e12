// ORB half of the front end on sm_100a: image pyramid, FAST-9/16 score map, per-cell threshold fallback + NMS +
// ordered compaction, intensity-centroid orientation, 7x7 fixed-point blur, 256-bit rBRIEF.
// Replaces ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc:412-472, 767-855, 1045-1134).
// Arithmetic follows cv2 4.13 semantics pinned in SURVEY.md Appendix A (the reference's OpenCV is un-vendored).
// The quadtree (DistributeOctTree, :541-765) is order-defining sequential list surgery and runs on the host
// between the two GPU phases; candidates and orientations reach it through mapped pinned memory (no extra copy).
#include "common.cuh"
#include "orb.h"
#include "img_kernels.cuh"
#include "../../include/olf_brief_pattern.h"
#include <algorithm>
#include <cmath>
#include <list>
#include <functional>

namespace olf {

// ---- pyramid: cv::resize INTER_LINEAR 8UC1, one launch per level ---------------------------
// coefficient tables (x: dst_w entries, y: dst_h entries) of {src index, c0, c1} are built on the host once per size.
struct LinCoef { int s; short c0, c1; };

__global__ void k_resize_linear(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                uint8_t* __restrict__ dst, int dw, int dh, int dpitch,
                                const LinCoef* __restrict__ cx, const LinCoef* __restrict__ cy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
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
__global__ void __launch_bounds__(256) k_fast_score(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score,
                                                    const __grid_constant__ LevelTable T, int min_th) {
    __shared__ uint8_t tile[TILE_H + 6][TILE_W + 8];
    int level, tx, ty;
    locate_tile(T, blockIdx.x, level, tx, ty);
    const int w = T.w[level], h = T.h[level], pitch = T.pitch[level];
    const uint8_t* img = pyr + T.off[level];
    uint8_t* out = score + T.off[level];
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
                int A = -255, B = -255;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
                    const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
                    A = max(A, mn9);
                    B = max(B, -mx9);
                }
                s = max(A, B) - 1;
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

// counts[cell] = number of keypoints, thr[cell] = threshold used; the last block to finish scans counts -> offsets.
__global__ void __launch_bounds__(128) k_cell_count(const uint8_t* __restrict__ score, const __grid_constant__ LevelTable T,
                                                    int ini_th, int* __restrict__ counts, int* __restrict__ thr,
                                                    int* __restrict__ offsets, int* __restrict__ total,
                                                    unsigned* __restrict__ ticket, int ncells) {
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
        if (lane == 0) { counts[c] = c_hi ? c_hi : c_lo; thr[c] = c_hi ? ini_th : 1; }
    }
    // last-block-done exclusive scan
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
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
        *total = acc;
        *ticket = 0;
    }
    __syncthreads();
    int acc = part[threadIdx.x];
    for (int i = b; i < e; ++i) { offsets[i] = acc; acc += ((volatile int*)counts)[i]; }
}

// cand[k] = {level, x, y, score} in reference order: cells row-major, pixels row-major inside a cell.
__global__ void __launch_bounds__(128) k_cell_write(const uint8_t* __restrict__ score, const __grid_constant__ LevelTable T,
                                                    const int* __restrict__ thr, const int* __restrict__ offsets,
                                                    int4* __restrict__ cand, int4* __restrict__ cand_host, int cap, int ncells) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + warp;
    if (c >= ncells) return;
    int level, ci, cj;
    locate_cell(T, c, level, ci, cj);
    const CellGeom g = cell_geom(T, level, ci, cj);
    if (!g.valid) return;
    const uint8_t* sc = score + T.off[level];
    const int pitch = T.pitch[level], th = thr[c];
    int pos = offsets[c];
    for (int y = g.y_lo; y < g.y_hi; ++y)
        for (int xb = g.x_lo; xb < g.x_hi; xb += 32) {
            const int x = xb + lane;
            int s = 0;
            if (x < g.x_hi) s = nms_score(sc, pitch, g, x, y);
            const bool keep = s >= th && s > 0;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int k = pos + __popc(m & ((1u << lane) - 1));
                if (k < cap) { const int4 v = make_int4(level, x, y, s); cand[k] = v; cand_host[k] = v; }
            }
            pos += __popc(m);
        }
}

// ---- IC_Angle (src/ORBextractor.cc:79-106): one warp per candidate, lanes = rows of the radius-15 disc -----
__constant__ int c_umax[16];
__global__ void __launch_bounds__(256) k_ic_angle(const uint8_t* __restrict__ pyr, const __grid_constant__ LevelTable T,
                                                  const int4* __restrict__ cand, const int* __restrict__ total, int cap,
                                                  float* __restrict__ angle_host) {
    const int lane = threadIdx.x & 31;
    const int n = min(*total, cap);
    for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < n; k += gridDim.x * 8) {
        const int4 c = cand[k];
        const uint8_t* center = pyr + T.off[c.x] + (size_t)c.z * T.pitch[c.x] + c.y;
        int m10 = 0, m01 = 0;
        if (lane < 31) {
            const int v = lane - 15;
            const int d = c_umax[v < 0 ? -v : v];
            const uint8_t* row = center + v * T.pitch[c.x];
            int rs = 0;
            for (int u = -d; u <= d; ++u) { const int p = row[u]; m10 += u * p; rs += p; }
            m01 = v * rs;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        if (lane == 0) angle_host[k] = fast_atan2_deg((float)m01, (float)m10);
    }
}

// ---- rBRIEF (src/ORBextractor.cc:110-149): one warp per keypoint, lane i -> descriptor byte i -------------
struct KeptKp { int level, x, y; float a, b; };      // a = cosf(angle), b = sinf(angle) (host glibc)
__constant__ signed char c_pattern[1024];
__global__ void __launch_bounds__(256) k_rbrief(const uint8_t* __restrict__ blur, const __grid_constant__ LevelTable T,
                                                const KeptKp* __restrict__ kps, int n, uint8_t* __restrict__ desc) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n) return;
    const KeptKp kp = kps[k];
    const int pitch = T.pitch[kp.level];
    const uint8_t* center = blur + T.off[kp.level] + (size_t)kp.y * pitch + kp.x;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const signed char* p = c_pattern + (lane * 8 + t) * 4;
        const float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
        const int r0 = __float2int_rn(fadd(fmul(x0, kp.b), fmul(y0, kp.a))), c0 = __float2int_rn(fsub(fmul(x0, kp.a), fmul(y0, kp.b)));
        const int r1 = __float2int_rn(fadd(fmul(x1, kp.b), fmul(y1, kp.a))), c1 = __float2int_rn(fsub(fmul(x1, kp.a), fmul(y1, kp.b)));
        const int t0 = center[r0 * pitch + c0], t1 = center[r1 * pitch + c1];
        val |= (t0 < t1) << t;
    }
    desc[(size_t)k * 32 + lane] = (uint8_t)val;
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
    // size-dependent state
    int img_w = 0, img_h = 0;
    LevelTable T;
    size_t pyr_bytes = 0;
    int ncells = 0, ntiles = 0;
    DevBuf<uint8_t> pyr, score, blur;
    DevBuf<LinCoef> coef;                    // per level: cx then cy
    std::vector<size_t> coef_off_x, coef_off_y;
    DevBuf<int> cell_counts, cell_thr, cell_off, total;
    DevBuf<unsigned> ticket;
    DevBuf<int4> cand;
    PinBuf<int4> cand_host;
    PinBuf<float> angle_host;
    PinBuf<int> total_host;
    int cand_cap = 0;
    DevBuf<KeptKp> kept;
    PinBuf<KeptKp> kept_host;
    DevBuf<uint8_t> desc;
    PinBuf<uint8_t> desc_host;
    PinBuf<uint8_t> img_stage;               // pinned staging for host images
    int kept_cap = 0;
    int last_ncand = 0;
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
        (rc = h->total.ensure(1)) || (rc = h->ticket.ensure(1)) || (rc = h->total_host.ensure(1))) return rc;
    OLF_CUDA(cudaMemset(h->ticket.p, 0, sizeof(unsigned)));
    // at most one NMS survivor per 2x2 block of the evaluated area
    h->cand_cap = (int)std::min<size_t>(off / 4 + 1024, (size_t)1 << 22);
    if ((rc = h->cand.ensure(h->cand_cap)) || (rc = h->cand_host.ensure(h->cand_cap)) || (rc = h->angle_host.ensure(h->cand_cap))) return rc;
    if ((rc = h->img_stage.ensure((size_t)w * hgt))) return rc;
    h->img_w = w; h->img_h = hgt;
    return OLF_OK;
}

// ---- quadtree distribution on the host (src/ORBextractor.cc:483-765), index based ----------------------------
namespace {
struct QNode {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<int> keys;
    bool no_more = false;
    long seq = 0;
    std::list<QNode>::iterator lit;
};
struct QCand { float x, y; int score; };
void q_divide(const QNode& n, const QCand* c, QNode& n1, QNode& n2, QNode& n3, QNode& n4) {
    const int halfX = (int)ceilf((float)(n.URx - n.ULx) / 2), halfY = (int)ceilf((float)(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy; n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy; n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy; n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy; n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    const size_t m = n.keys.size();
    n1.keys.reserve(m); n2.keys.reserve(m / 2 + 1); n3.keys.reserve(m / 2 + 1); n4.keys.reserve(m / 2 + 1);
    for (int idx : n.keys) {
        const QCand& kp = c[idx];
        if (kp.x < n1.URx) { if (kp.y < n1.BRy) n1.keys.push_back(idx); else n3.keys.push_back(idx); }
        else if (kp.y < n1.BRy) n2.keys.push_back(idx);
        else n4.keys.push_back(idx);
    }
    if (n1.keys.size() == 1) n1.no_more = true;
    if (n2.keys.size() == 1) n2.no_more = true;
    if (n3.keys.size() == 1) n3.no_more = true;
    if (n4.keys.size() == 1) n4.no_more = true;
}
void quadtree(const QCand* c, int nc, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& result) {
    result.clear();
    const int nIni = std::max((int)roundf((float)(maxX - minX) / (maxY - minY)), 1);
    const float hX = (float)(maxX - minX) / nIni;
    std::list<QNode> nodes;
    std::vector<QNode*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        QNode ni;
        ni.ULx = (int)(hX * (float)i); ni.ULy = 0; ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY; ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        nodes.push_back(std::move(ni));
        ini[i] = &nodes.back();
    }
    for (int i = 0; i < nc; i++) ini[std::min((int)(c[i].x / hX), nIni - 1)]->keys.push_back(i);
    for (auto lit = nodes.begin(); lit != nodes.end();) {
        if (lit->keys.size() == 1) { lit->no_more = true; ++lit; }
        else if (lit->keys.empty()) lit = nodes.erase(lit);
        else ++lit;
    }
    std::vector<std::pair<int, QNode*>> expand;
    auto push_child = [&](QNode& n, int* n_to_expand) {
        if (n.keys.empty()) return;
        n.seq = seq++;
        const bool many = n.keys.size() > 1;
        const int sz = (int)n.keys.size();
        nodes.push_front(std::move(n));
        if (many) {
            if (n_to_expand) ++*n_to_expand;
            expand.emplace_back(sz, &nodes.front());
            nodes.front().lit = nodes.begin();
        }
    };
    bool finish = false;
    while (!finish) {
        int prev = (int)nodes.size(), n_to_expand = 0;
        expand.clear();
        for (auto lit = nodes.begin(); lit != nodes.end();) {
            if (lit->no_more) { ++lit; continue; }
            QNode n1, n2, n3, n4;
            q_divide(*lit, c, n1, n2, n3, n4);
            push_child(n1, &n_to_expand); push_child(n2, &n_to_expand); push_child(n3, &n_to_expand); push_child(n4, &n_to_expand);
            lit = nodes.erase(lit);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
        else if ((int)nodes.size() + n_to_expand * 3 > N) {
            while (!finish) {
                prev = (int)nodes.size();
                std::vector<std::pair<int, QNode*>> order = expand;
                expand.clear();
                // canonical tie-break: (size, creation sequence), SURVEY Appendix C.1
                std::sort(order.begin(), order.end(), [](const std::pair<int, QNode*>& a, const std::pair<int, QNode*>& b) {
                    return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq; });
                for (int j = (int)order.size() - 1; j >= 0; j--) {
                    QNode n1, n2, n3, n4;
                    q_divide(*order[j].second, c, n1, n2, n3, n4);
                    push_child(n1, nullptr); push_child(n2, nullptr); push_child(n3, nullptr); push_child(n4, nullptr);
                    nodes.erase(order[j].second->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev) finish = true;
            }
        }
    }
    for (const QNode& n : nodes) {
        int best = n.keys[0];
        int max_resp = c[best].score;
        for (size_t k = 1; k < n.keys.size(); k++)
            if (c[n.keys[k]].score > max_resp) { best = n.keys[k]; max_resp = c[best].score; }
        result.push_back(best);
    }
}
}  // namespace

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
    h->pyr.release(); h->score.release(); h->blur.release(); h->coef.release(); h->cell_counts.release(); h->cell_thr.release();
    h->cell_off.release(); h->total.release(); h->ticket.release(); h->cand.release(); h->cand_host.release(); h->angle_host.release();
    h->total_host.release(); h->kept.release(); h->kept_host.release(); h->desc.release(); h->desc_host.release(); h->img_stage.release();
    delete h;
}

// phase 1: upload + pyramid + FAST + cells + orientation + blur, all enqueued on the handle's stream
static int orb_enqueue_phase1(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device) {
    int rc = orb_ensure_size(h, w, hgt);
    if (rc) return rc;
    const LevelTable& T = h->T;
    cudaStream_t s = h->stream;
    if (on_device) {
        OLF_CUDA(cudaMemcpy2DAsync(h->pyr.p, T.pitch[0], img, stride, w, hgt, cudaMemcpyDeviceToDevice, s));
    } else {
        for (int y = 0; y < hgt; ++y) memcpy(h->img_stage.p + (size_t)y * w, img + (size_t)y * stride, w);
        OLF_CUDA(cudaMemcpy2DAsync(h->pyr.p, T.pitch[0], h->img_stage.p, w, w, hgt, cudaMemcpyHostToDevice, s));
    }
    for (int l = 1; l < h->nlevels; ++l) {
        dim3 b(32, 8), g((T.w[l] + 31) / 32, (T.h[l] + 7) / 8);
        k_resize_linear<<<g, b, 0, s>>>(h->pyr.p + T.off[l - 1], T.w[l - 1], T.h[l - 1], T.pitch[l - 1],
                                        h->pyr.p + T.off[l], T.w[l], T.h[l], T.pitch[l],
                                        h->coef.p + h->coef_off_x[l], h->coef.p + h->coef_off_y[l]);
    }
    k_fast_score<<<h->ntiles, 256, 0, s>>>(h->pyr.p, h->score.p, T, h->min_th);
    if (h->ncells > 0) {
        const int nb = (h->ncells + 3) / 4;
        k_cell_count<<<nb, 128, 0, s>>>(h->score.p, T, h->ini_th, h->cell_counts.p, h->cell_thr.p, h->cell_off.p, h->total.p, h->ticket.p, h->ncells);
        k_cell_write<<<nb, 128, 0, s>>>(h->score.p, T, h->cell_thr.p, h->cell_off.p, h->cand.p, h->cand_host.d, h->cand_cap, h->ncells);
        k_ic_angle<<<296, 256, 0, s>>>(h->pyr.p, T, h->cand.p, h->total.p, h->cand_cap, h->angle_host.d);
    } else {
        OLF_CUDA(cudaMemsetAsync(h->total.p, 0, sizeof(int), s));
    }
    OLF_CUDA(cudaMemcpyAsync(h->total_host.p, h->total.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    // ORB blur: 7x7 sigma 2 -> [18,34,48,56,48,34,18] (:1088)
    k_blur_q8<7><<<h->ntiles, 256, 0, s>>>(h->pyr.p, h->blur.p, T, 18, 34, 48, 56);
    count_launches((h->nlevels - 1) + 2 + (h->ncells > 0 ? 3 : 0));
    OLF_CUDA(cudaGetLastError());
    return OLF_OK;
}

int orb_extract(OrbImpl* h, const uint8_t* img, int w, int hgt, int stride, bool on_device,
                olf_keypoint* kps, uint8_t* desc, int cap, int* n, const std::function<void()>* on_phase1_enqueued) {
    if (!h || !n) return OLF_ERR_ARG;
    *n = 0;
    if (!img || w <= 0 || hgt <= 0) return OLF_OK;           // _image.empty(): silent return (:1048)
    if (stride < w || !kps || !desc) { set_last_error("olf_orb_extract: bad arguments"); return OLF_ERR_ARG; }
    OLF_CUDA(cudaSetDevice(h->device));
    int rc = orb_enqueue_phase1(h, img, w, hgt, stride, on_device);
    if (rc) { if (on_phase1_enqueued) (*on_phase1_enqueued)(); return rc; }
    {   // wait for phase 1 only: whatever another thread enqueues on a shared stream after this mark is not waited for
        cudaEvent_t ev;
        const cudaError_t e = stream_record(h->stream, &ev);
        if (on_phase1_enqueued) (*on_phase1_enqueued)();
        OLF_CUDA(e);
        OLF_CUDA(event_wait(ev));
    }
    const int ncand = *h->total_host.p;
    if (ncand > h->cand_cap) { set_last_error("FAST candidate buffer overflow"); return OLF_ERR_CAPACITY; }
    h->last_ncand = ncand;
    const LevelTable& T = h->T;
    // host: quadtree per level (cand is grouped by level because cells are enumerated level by level)
    const int4* cand = h->cand_host.p;
    const float* ang = h->angle_host.p;
    std::vector<QCand> qc;
    std::vector<int> keep;
    std::vector<KeptKp> kept;
    const float factorPI = (float)(M_PI / 180.f);
    int total = 0, pos = 0;
    for (int level = 0; level < h->nlevels; ++level) {
        const int b = pos;
        while (pos < ncand && cand[pos].x == level) ++pos;
        const int m = pos - b;
        if (m == 0) continue;
        qc.resize(m);
        for (int i = 0; i < m; ++i) qc[i] = {(float)(cand[b + i].y - 16), (float)(cand[b + i].z - 16), cand[b + i].w};
        quadtree(qc.data(), m, 16, T.w[level] - 16, 16, T.h[level] - 16, h->feats_per_level[level], keep);
        if (total + (int)keep.size() > cap) { set_last_error("olf_orb_extract: keypoint capacity too small"); return OLF_ERR_CAPACITY; }
        const int scaledPatchSize = (int)(31 * h->scale[level]);
        for (int idx : keep) {
            const int4 c = cand[b + idx];
            olf_keypoint& kp = kps[total++];
            kp.angle = ang[b + idx];
            kp.response = (float)c.w;
            kp.octave = level;
            kp.size = (float)scaledPatchSize;
            kp.x = (float)c.y; kp.y = (float)c.z;
            if (level != 0) { kp.x *= h->scale[level]; kp.y *= h->scale[level]; }
            const float a = kp.angle * factorPI;
            kept.push_back({level, c.y, c.z, cosf(a), sinf(a)});
        }
    }
    *n = total;
    if (total == 0) return OLF_OK;
    // phase 2: rBRIEF on the blurred pyramid
    if (total > h->kept_cap) {
        const int ncap = std::max(total * 2, 4096);
        if ((rc = h->kept.ensure(ncap)) || (rc = h->kept_host.ensure(ncap)) || (rc = h->desc.ensure((size_t)ncap * 32)) || (rc = h->desc_host.ensure((size_t)ncap * 32))) return rc;
        h->kept_cap = ncap;
    }
    memcpy(h->kept_host.p, kept.data(), kept.size() * sizeof(KeptKp));
    OLF_CUDA(cudaMemcpyAsync(h->kept.p, h->kept_host.p, kept.size() * sizeof(KeptKp), cudaMemcpyHostToDevice, h->stream));
    k_rbrief<<<(total + 7) / 8, 256, 0, h->stream>>>(h->blur.p, T, h->kept.p, total, h->desc.p);
    count_launches(1);
    OLF_CUDA(cudaMemcpyAsync(h->desc_host.p, h->desc.p, (size_t)total * 32, cudaMemcpyDeviceToHost, h->stream));
    OLF_CUDA(stream_sync(h->stream));
    OLF_CUDA(cudaGetLastError());
    memcpy(desc, h->desc_host.p, (size_t)total * 32);
    return OLF_OK;
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
int orb_last_candidates(OrbImpl* h, int* out, int cap, int* n) {
    if (!h || !n) return OLF_ERR_ARG;
    *n = h->last_ncand;
    if (h->last_ncand > cap) return OLF_ERR_CAPACITY;
    memcpy(out, h->cand_host.p, (size_t)h->last_ncand * sizeof(int4));
    return OLF_OK;
}
const OrbDeviceView orb_device_view(const OrbImpl* h) {
    OrbDeviceView v;
    v.pyr = h->pyr.p; v.nlevels = h->nlevels;
    for (int l = 0; l < h->nlevels; ++l) { v.w[l] = h->T.w[l]; v.h[l] = h->T.h[l]; v.pitch[l] = h->T.pitch[l]; v.off[l] = h->T.off[l]; v.scale[l] = h->scale[l]; v.inv_scale[l] = h->inv_scale[l]; }
    v.device = h->device;
    return v;
}
void orb_scale_tables(const OrbImpl* h, const float** s, const float** is, const float** s2, const float** is2, const int** fpl, int* nlevels) {
    if (s) *s = h->scale.data(); if (is) *is = h->inv_scale.data(); if (s2) *s2 = h->sigma2.data(); if (is2) *is2 = h->inv_sigma2.data();
    if (fpl) *fpl = h->feats_per_level.data(); if (nlevels) *nlevels = h->nlevels;
}

}  // namespace olf
