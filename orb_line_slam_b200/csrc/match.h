// internal C++ interface of the matching half (see match.cu)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/olf_abi.h"
#include "common.cuh"
namespace olf {
struct OrbImpl;
// the calling thread's next matcher calls run on `s` (nullptr: back to the thread's own stream)
void match_use_stream(cudaStream_t s);
// the calling thread's next matcher calls use this scratch context instead of the thread's own (nullptr: back to the thread's)
struct MatchCtx;
MatchCtx* match_ctx_create(int device);
void match_ctx_destroy(MatchCtx* c);
void match_use_ctx(MatchCtx* c);
int knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx0, int* dist0, int* idx1, int* dist1, int device);
int knn2_bench(int n1, int n2, int iters, int device, double* kernel_ms, double* popc_word_pairs_per_s);
int match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int mutual, int* m12, int* nmatches, int device);
int stereo_points(OrbImpl* left, OrbImpl* right, const olf_keypoint* kl, const uint8_t* dl, int N, const olf_keypoint* kr, const uint8_t* dr, int Nr,
                  float bf, float fx, float* uRight, float* depth);
// device-resident variant used by the whole-frame driver (see match.cu): per-frame workspace, results in `out` (pinned)
// as [uRight[cap] | depth[cap]] once the stream has been waited for
struct StereoWs { DevBuf<float> u, d; DevBuf<int> sad; PinBuf<float> out; int cap = 0; };
int stereo_ws_ensure(StereoWs* ws, int cap);
void stereo_ws_release(StereoWs* ws);
int stereo_points_enqueue(StereoWs* ws, OrbImpl* left, OrbImpl* right, float bf, float fx, int cap, cudaStream_t s);
int stereo_lines_batch(int nf, const olf_keyline* const* kl, const uint8_t* const* dl, const int* n1s, const olf_keyline* const* kr, const uint8_t* const* dr, const int* n2s,
                       int img_w, int img_h, const olf_line_match_params* P, int* const* matches12, float* const* disp, double* const* le, int device);
int stereo_lines(const olf_keyline* kl, const uint8_t* dl, int n1, const olf_keyline* kr, const uint8_t* dr, int n2, int img_w, int img_h,
                 const olf_line_match_params* P, int* matches12, float* disp, double* le, int device);
int match_grid_lines(const int* lines1, const uint8_t* desc1, int n1, const olf_grid_csr* grid, const uint8_t* desc2, int n2, const double* dir2,
                     const int* window, const olf_line_match_params* P, int* matches12, int* nmatches, int device);
int distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best, int device);
int search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches, int device);
int search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur, int* nmatches, int device);
struct VocabImpl;
VocabImpl* vocab_create(const olf_vocab_desc* v, int device);
void vocab_destroy(VocabImpl* h);
int bow_transform(VocabImpl* V, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id);
int bow_assemble(const int* word_id, const double* weight, const int* node_id, int n, int* bow_word, double* bow_value, int* n_words,
                 int* fv_node, int* fv_begin, int* fv_index, int* n_nodes);
int search_by_bow(const olf_bow_match_args* a, int* match_f, int* nmatches, int device);
int search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2, int* matches12, int* nmatches, int device);
int window_search(const olf_window_search_args* a, int* best_idx, int* best_dist, int device);
int search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2, const olf_camera* cam,
                              float* prev_matched, int window_size, float nn_ratio, int check_orientation, int* m12, int* nmatches, int device);
int search_for_triangulation(const olf_triangulation_args* a, int* matches12, int* nmatches, int device);
}
