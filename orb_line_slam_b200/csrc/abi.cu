// extern "C" surface of libolf.so: thin forwarding layer over the C++ implementations (see include/olf_abi.h).
#include "common.cuh"
#include "orb.h"
#include "line.h"
#include "match.h"
#include <mutex>
#include <atomic>
#include <time.h>

namespace olf {
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<long long> g_allocs{0};
void count_allocs(int n) { g_allocs.fetch_add(n, std::memory_order_relaxed); }
long long launches_total() { return g_launches.load(); }
static int sync_mode() {
    // OLF_SYNC=spin  : busy-wait (lowest latency, one busy host core per waiting thread)
    // OLF_SYNC=block : blocking-sync event (thread sleeps until the driver's interrupt; wake-up jitter under load)
    // default        : poll the event, yielding the core between polls (bounded latency, cores stay available to other rigs)
    static const int mode = [] { const char* e = getenv("OLF_SYNC"); return !e ? 0 : (std::string(e) == "spin" ? 1 : (std::string(e) == "block" ? 2 : 0)); }();
    return mode;
}
// per-thread events of the stateless entry points (matchers); destroyed when the thread exits
struct ThreadEvents {
    cudaEvent_t ev[16] = {nullptr};
    ~ThreadEvents() { for (int d = 0; d < 16; ++d) if (ev[d]) { cudaEventDestroy(ev[d]); cudaGetLastError(); } }
};
cudaError_t stream_record(cudaStream_t s, cudaEvent_t* out) {
    static thread_local ThreadEvents te;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
    if (!te.ev[dev]) { e = cudaEventCreateWithFlags(&te.ev[dev], (sync_mode() == 2 ? cudaEventBlockingSync : 0) | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    *out = te.ev[dev];
    return cudaEventRecord(te.ev[dev], s);
}
cudaError_t event_wait(cudaEvent_t ev, bool blocking) {
    const int mode = sync_mode();
    if (mode == 2 || (blocking && mode != 1)) return cudaEventSynchronize(ev);
    for (int spins = 0;; ++spins) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
        if (mode == 1 || spins < 64) { for (int k = 0; k < 64; ++k) __builtin_ia32_pause(); }
        else { struct timespec ts = {0, 20000}; nanosleep(&ts, nullptr); }
    }
}
cudaError_t stream_sync(cudaStream_t s) {
    cudaEvent_t ev;
    const cudaError_t e = stream_record(s, &ev);
    if (e != cudaSuccess) return e;
    return event_wait(ev);
}
void set_last_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? OLF_ERR_NO_DEVICE : OLF_ERR_CUDA;
}
}  // namespace olf
using namespace olf;

extern "C" {
const char* olf_last_error(void) { return g_err.c_str(); }
long long olf_kernel_launch_count(void) { return launches_total(); }
long long olf_alloc_count(void) { return g_allocs.load(); }
int olf_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

olf_orb* olf_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device) {
    return (olf_orb*)orb_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, device);
}
void olf_orb_destroy(olf_orb* h) { orb_destroy((OrbImpl*)h); }
int olf_orb_extract(olf_orb* h, const uint8_t* img, int width, int height, int stride, olf_keypoint* kps, uint8_t* desc, int cap, int* n) {
    return orb_extract((OrbImpl*)h, img, width, height, stride, false, kps, desc, cap, n);
}
int olf_trig_sweep(unsigned first_bits, unsigned stride, unsigned count, float* cos_sin, int device) { return orb_trig_sweep(first_bits, stride, count, cos_sin, device); }
int olf_orb_extract_dev(olf_orb* h, const uint8_t* d_img, int width, int height, int stride, olf_keypoint* kps, uint8_t* desc, int cap, int* n) {
    return orb_extract((OrbImpl*)h, d_img, width, height, stride, true, kps, desc, cap, n);
}
int olf_orb_level_size(const olf_orb* h, int level, int* width, int* height) { return orb_level_size((const OrbImpl*)h, level, width, height); }
int olf_orb_get_level(olf_orb* h, int level, uint8_t* dst, int dst_stride) { return orb_get_level((OrbImpl*)h, level, dst, dst_stride); }
int olf_orb_scale_factors(const olf_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    if (!h) return OLF_ERR_ARG;
    const float *s, *is, *s2, *is2; int nl;
    orb_scale_tables((const OrbImpl*)h, &s, &is, &s2, &is2, nullptr, &nl);
    for (int i = 0; i < nl; ++i) { if (scale) scale[i] = s[i]; if (inv_scale) inv_scale[i] = is[i]; if (sigma2) sigma2[i] = s2[i]; if (inv_sigma2) inv_sigma2[i] = is2[i]; }
    return OLF_OK;
}
int olf_orb_features_per_level(const olf_orb* h, int* out) {
    if (!h || !out) return OLF_ERR_ARG;
    const int* f; int nl;
    orb_scale_tables((const OrbImpl*)h, nullptr, nullptr, nullptr, nullptr, &f, &nl);
    for (int i = 0; i < nl; ++i) out[i] = f[i];
    return OLF_OK;
}
int olf_orb_last_candidates(olf_orb* h, int* out, int cap, int* n) { return orb_last_candidates((OrbImpl*)h, out, cap, n); }

olf_line* olf_line_create(const olf_line_params* p, int device) { return (olf_line*)line_create(p, device); }
void olf_line_destroy(olf_line* h) { line_destroy((LineImpl*)h); }
int olf_line_extract(olf_line* h, const uint8_t* img, int width, int height, int stride, olf_keyline* kls, uint8_t* desc, int cap, int* n) {
    return line_extract((LineImpl*)h, img, width, height, stride, false, kls, desc, cap, n);
}
int olf_line_extract_dev(olf_line* h, const uint8_t* d_img, int width, int height, int stride, olf_keyline* kls, uint8_t* desc, int cap, int* n) {
    return line_extract((LineImpl*)h, d_img, width, height, stride, true, kls, desc, cap, n);
}
int olf_lsd_detect(olf_line* h, const uint8_t* img, int width, int height, int stride, float* segs, int cap, int* n) {
    return line_lsd_detect((LineImpl*)h, img, width, height, stride, false, segs, cap, n);
}
int olf_lbd_compute(olf_line* h, const uint8_t* img, int width, int height, int stride, const olf_keyline* kls, int n, uint8_t* desc) {
    return line_lbd_compute((LineImpl*)h, img, width, height, stride, kls, n, desc);
}
int olf_line_trace(olf_line* h, int* out, int max_rounds) { return line_trace((LineImpl*)h, out, max_rounds); }
int olf_line_last_stats(const olf_line* h, int* out8) { if (!h || !out8) return OLF_ERR_ARG; line_last_stats((const LineImpl*)h, out8); return OLF_OK; }

int olf_knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx0, int* dist0, int* idx1, int* dist1, int device) {
    return knn2_hamming(d1, n1, d2, n2, idx0, dist0, idx1, dist1, device);
}
int olf_knn2_bench(int n1, int n2, int iters, int device, double* kernel_ms, double* popc_word_pairs_per_s) { return knn2_bench(n1, n2, iters, device, kernel_ms, popc_word_pairs_per_s); }
int olf_match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* matches12, int* nmatches, int device) {
    return match_lines(d1, n1, d2, n2, nnr, 0, matches12, nmatches, device);
}
int olf_match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr_matches, int* matches12, int* nmatches, int device) {
    return match_lines(d1, n1, d2, n2, nnr, best_lr_matches, matches12, nmatches, device);
}
int olf_stereo_points(olf_orb* left, olf_orb* right, const olf_keypoint* kps_l, const uint8_t* desc_l, int n_l,
                      const olf_keypoint* kps_r, const uint8_t* desc_r, int n_r, float bf, float fx, float* u_right, float* depth) {
    return stereo_points((OrbImpl*)left, (OrbImpl*)right, kps_l, desc_l, n_l, kps_r, desc_r, n_r, bf, fx, u_right, depth);
}
int olf_stereo_lines(const olf_keyline* kls_l, const uint8_t* desc_l, int n_l, const olf_keyline* kls_r, const uint8_t* desc_r, int n_r,
                     int img_width, int img_height, const olf_line_match_params* p, int* matches12, float* disp_s_e, double* le, int device) {
    return stereo_lines(kls_l, desc_l, n_l, kls_r, desc_r, n_r, img_width, img_height, p, matches12, disp_s_e, le, device);
}
int olf_match_grid_lines(const int* lines1, const uint8_t* desc1, int n1, const olf_grid_csr* grid, const uint8_t* desc2, int n2, const double* dir2,
                         const int* window, const olf_line_match_params* p, int* matches12, int* nmatches, int device) {
    return match_grid_lines(lines1, desc1, n1, grid, desc2, n2, dir2, window, p, matches12, nmatches, device);
}
int olf_distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best, int device) { return distinctive_descriptors(desc, group_begin, n_groups, best, device); }
int olf_search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches, int device) {
    return search_by_projection_last(a, assigned_cur, cur_point, nmatches, device);
}
int olf_search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur, int* nmatches, int device) {
    return search_by_projection_map(a, assigned_cur, nmatches, device);
}
olf_vocab* olf_vocab_create(const olf_vocab_desc* v, int device) { return (olf_vocab*)vocab_create(v, device); }
void olf_vocab_destroy(olf_vocab* v) { vocab_destroy((VocabImpl*)v); }
int olf_bow_transform(olf_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id) {
    return bow_transform((VocabImpl*)v, desc, n, levelsup, word_id, weight, node_id);
}
int olf_bow_assemble(const int* word_id, const double* weight, const int* node_id, int n, int* bow_word, double* bow_value, int* n_words,
                     int* fv_node, int* fv_begin, int* fv_index, int* n_nodes) {
    return bow_assemble(word_id, weight, node_id, n, bow_word, bow_value, n_words, fv_node, fv_begin, fv_index, n_nodes);
}
int olf_search_by_bow(const olf_bow_match_args* a, int* match_f, int* nmatches, int device) { return search_by_bow(a, match_f, nmatches, device); }
int olf_search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2, int* matches12, int* nmatches, int device) { return search_by_bow_kf(a, has_point2, matches12, nmatches, device); }
int olf_window_search(const olf_window_search_args* a, int* best_idx, int* best_dist, int device) { return window_search(a, best_idx, best_dist, device); }
int olf_search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2, const olf_camera* cam,
                                  float* prev_matched, int window_size, float nn_ratio, int check_orientation, int* matches12, int* nmatches, int device) {
    return search_for_initialization(kps1, desc1, n1, kps2, desc2, n2, cam, prev_matched, window_size, nn_ratio, check_orientation, matches12, nmatches, device);
}
int olf_search_for_triangulation(const olf_triangulation_args* a, int* matches12, int* nmatches, int device) { return search_for_triangulation(a, matches12, nmatches, device); }
}
