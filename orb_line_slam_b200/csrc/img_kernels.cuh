// Image kernels shared by the ORB and line halves: level table, 64x16 tiling, fixed-point Gaussian blur.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace olf {

// ------------------------------------------------------------------------------------------------------
// level table passed by value to the multi-level kernels
struct LevelTable {
    int n;
    int w[OLF_MAX_LEVELS], h[OLF_MAX_LEVELS], pitch[OLF_MAX_LEVELS];
    unsigned off[OLF_MAX_LEVELS];            // byte offset of the level in the pyramid / score / blur buffers
    int tiles_x[OLF_MAX_LEVELS], tile_start[OLF_MAX_LEVELS + 1];   // 64x16 tiles (score, blur kernels)
    // FAST cell grid (src/ORBextractor.cc:783-789)
    int ncols[OLF_MAX_LEVELS], nrows[OLF_MAX_LEVELS], wcell[OLF_MAX_LEVELS], hcell[OLF_MAX_LEVELS];
    int cell_start[OLF_MAX_LEVELS + 1];
};

#define TILE_W 64
#define TILE_H 16
static __device__ __forceinline__ bool has_run9(unsigned m) {           // cyclic run of >= 9 set bits in a 16-bit mask
    unsigned v = m | (m << 16);
    unsigned r = v & (v >> 1);
    r &= r >> 2;
    r &= r >> 4;          // runs of 8
    r &= v >> 8;          // runs of 9
    return r != 0;
}
static __device__ __forceinline__ void locate_tile(const LevelTable& T, int b, int& level, int& tx, int& ty) {
    level = 0;
#pragma unroll 1
    while (level + 1 < T.n && b >= T.tile_start[level + 1]) ++level;
    const int t = b - T.tile_start[level];
    tx = t % T.tiles_x[level];
    ty = t / T.tiles_x[level];
}

// ---- fixed-point separable Gaussian blur on 8U, all levels in one launch ---------------------
// H pass 8.8 (u16), V pass 16.16, rounding (acc + 2^15) >> 16, BORDER_REFLECT_101 at the level edge.
// A batch of images per launch (blockIdx.y): the rigs keep 2..8 images in flight and every kernel of this front end is
// launch-bound on one 0.9-MB image, so one launch serves all images of a call.
#define IMG_MAX_BATCH 16
struct BlurBatch { const uint8_t* src[IMG_MAX_BATCH]; uint8_t* dst[IMG_MAX_BATCH]; };
// TMA variant: the tensor maps travel as a __grid_constant__ kernel parameter (the one placement that needs no proxy fence and
// that no stray device write can reach); m[image * T.n + level]
template <int NM> struct alignas(64) TmaSet { CUtensorMap m[NM]; };
template <int K>
static __global__ void __launch_bounds__(256) k_blur_q8(const __grid_constant__ BlurBatch BB,
                                                 const __grid_constant__ LevelTable T, const int q0, const int q1, const int q2, const int q3) {
    const uint8_t* __restrict__ src = BB.src[blockIdx.y];
    uint8_t* __restrict__ dst = BB.dst[blockIdx.y];
    constexpr int R = K / 2;
    __shared__ uint8_t tile[TILE_H + 2 * R][TILE_W + 2 * R + 2];
    __shared__ uint16_t hbuf[TILE_H + 2 * R][TILE_W];
    const int q[4] = {q0, q1, q2, q3};       // q[i] = weight at distance R-i from the centre... stored outer->centre
    int level, tx, ty;
    locate_tile(T, blockIdx.x, level, tx, ty);
    const int w = T.w[level], h = T.h[level], pitch = T.pitch[level];
    const uint8_t* img = src + T.off[level];
    uint8_t* out = dst + T.off[level];
    const int x0 = tx * TILE_W, y0 = ty * TILE_H;
    for (int i = threadIdx.x; i < (TILE_H + 2 * R) * (TILE_W + 2 * R); i += 256) {
        const int r = i / (TILE_W + 2 * R), c = i % (TILE_W + 2 * R);
        const int gx = reflect101(min(x0 + c - R, w - 1 + R), w), gy = reflect101(min(y0 + r - R, h - 1 + R), h);
        tile[r][c] = img[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (TILE_H + 2 * R) * TILE_W; i += 256) {
        const int r = i / TILE_W, c = i % TILE_W;
        unsigned a = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int dk = k < R ? k : K - 1 - k; a += (unsigned)q[dk] * tile[r][c + k]; }
        hbuf[r][c] = (uint16_t)a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TILE_H * TILE_W; i += 256) {
        const int r = i / TILE_W, c = i % TILE_W;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        unsigned a = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int dk = k < R ? k : K - 1 - k; a += (unsigned)q[dk] * hbuf[r + k][c]; }
        out[(size_t)y * pitch + x] = (uint8_t)((a + (1u << 15)) >> 16);
    }
}

// The same blur with the tile + halo staged by the TMA unit: box = TMA_BOX_W x (TILE_H + 2R) bytes starting at (x0 - 16, y0 - R);
// one thread arms the mbarrier and issues cp.async.bulk.tensor.2d, everybody waits on the barrier.  TMA fills elements outside
// the image with zeros; BORDER_REFLECT_101 is restored in shared memory for the tiles that touch the image edge (the mirrored
// pixels are inside the same box).
#define TMA_BOX_W 96
#define TMA_PAD 16
template <int K, int NM>
static __global__ void __launch_bounds__(256) k_blur_q8_tma(const __grid_constant__ BlurBatch BB, const __grid_constant__ TmaSet<NM> M,
                                                     const __grid_constant__ LevelTable T, const int q0, const int q1, const int q2, const int q3) {
    constexpr int R = K / 2;
    constexpr int ROWS = TILE_H + 2 * R;
    __shared__ __align__(128) uint8_t tile[ROWS][TMA_BOX_W];
    __shared__ uint16_t hbuf[ROWS][TILE_W];
    __shared__ __align__(8) uint64_t bar;
    const int q[4] = {q0, q1, q2, q3};
    int level, tx, ty;
    locate_tile(T, blockIdx.x, level, tx, ty);
    const int w = T.w[level], h = T.h[level], pitch = T.pitch[level];
    uint8_t* out = BB.dst[blockIdx.y] + T.off[level];
    const int x0 = tx * TILE_W, y0 = ty * TILE_H;
    if (threadIdx.x == 0) tma::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_expect_tx(&bar, ROWS * TMA_BOX_W);
        tma::load_2d(&tile[0][0], &M.m[blockIdx.y * T.n + level], x0 - TMA_PAD, y0 - R, &bar);
    }
    tma::mbar_wait(&bar, 0);
    // reflect-101 fix-up of the zero-filled cells (edge tiles only): columns first, then whole rows
    const bool edge_x = x0 - R < 0 || x0 + TILE_W + R > w, edge_y = y0 - R < 0 || y0 + ROWS - R > h;
    if (edge_x) {
        for (int i = threadIdx.x; i < ROWS * (TILE_W + 2 * R); i += 256) {
            const int r = i / (TILE_W + 2 * R), c = i % (TILE_W + 2 * R);
            const int gx = x0 + c - R, gy = y0 + r - R;
            if ((gx < 0 || gx >= w) && gy >= 0 && gy < h && gx <= w - 1 + R) tile[r][c - R + TMA_PAD] = tile[r][reflect101(gx, w) - x0 + TMA_PAD];
        }
        __syncthreads();
    }
    if (edge_y) {
        for (int i = threadIdx.x; i < ROWS * (TILE_W + 2 * R); i += 256) {
            const int r = i / (TILE_W + 2 * R), c = i % (TILE_W + 2 * R);
            const int gy = y0 + r - R;
            if ((gy < 0 || gy >= h) && gy <= h - 1 + R) tile[r][c - R + TMA_PAD] = tile[reflect101(gy, h) - y0 + R][c - R + TMA_PAD];
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < ROWS * TILE_W; i += 256) {
        const int r = i / TILE_W, c = i % TILE_W;
        unsigned a = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int dk = k < R ? k : K - 1 - k; a += (unsigned)q[dk] * tile[r][c + k + TMA_PAD - R]; }
        hbuf[r][c] = (uint16_t)a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TILE_H * TILE_W; i += 256) {
        const int r = i / TILE_W, c = i % TILE_W;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        unsigned a = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const int dk = k < R ? k : K - 1 - k; a += (unsigned)q[dk] * hbuf[r + k][c]; }
        out[(size_t)y * pitch + x] = (uint8_t)((a + (1u << 15)) >> 16);
    }
}

// host helper: one tensor map per level of a pyramid-shaped buffer (box rows = TILE_H + 2 * (K / 2)); false when the driver
// offers no cuTensorMapEncodeTiled (the callers then keep the plain-load kernel)
static inline bool build_level_maps(const uint8_t* base, const LevelTable& T, int K, CUtensorMap* host_maps) {
    if (getenv("OLF_NO_TMA")) return false;
    for (int l = 0; l < T.n; ++l)
        if (!tma::encode_u8_2d(&host_maps[l], base + T.off[l], T.w[l], T.h[l], T.pitch[l], TMA_BOX_W, TILE_H + 2 * (K / 2))) return false;
    return true;
}

}  // namespace olf
