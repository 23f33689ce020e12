/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 * C interface of the CPU restatement of the reference hot path (see oracle/README.md).
 * Mirrors include/olf_abi.h with the prefix orc_ so the parity tests can call both sides alike.
 * Parity pinning: PINNED.  OpenCV halves against cv2 4.13.0 (tests/test_oracle_vs_cv2.py, tests/golden/); the vendored halves
 * (ORBextractor incl. quadtree and rBRIEF, LSD KeyLines, LBD, every matcher, stereo association) against the reference's own
 * sources compiled in oracle/_ref (make -C oracle ref; tests/test_oracle_vs_ref.py) -- the reference ships no tests or golden
 * vectors of its own.
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include "../include/olf_abi.h"
#ifdef __cplusplus
namespace orc { struct Image8; }
struct orc_orb; struct orc_line;
const orc::Image8* orc_orb_level_image(const orc_orb* h, int level);
const float* orc_orb_scales(const orc_orb* h);
const float* orc_orb_inv_scales(const orc_orb* h);
extern "C" {
#else
typedef struct orc_orb orc_orb; typedef struct orc_line orc_line;
#endif
orc_orb* orc_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast);
void orc_orb_destroy(orc_orb* h);
int orc_orb_extract(orc_orb* h, const uint8_t* img, int w, int hgt, int stride, olf_keypoint* kps, uint8_t* desc, int cap, int* n);
int orc_orb_level_size(const orc_orb* h, int level, int* w, int* hh);
int orc_orb_get_level(orc_orb* h, int level, uint8_t* dst, int dst_stride);
int orc_orb_scale_factors(const orc_orb* h, float* s, float* is, float* s2, float* is2);
int orc_orb_features_per_level(const orc_orb* h, int* out);
int orc_orb_last_candidates(orc_orb* h, int* out, int cap, int* n);

orc_line* orc_line_create(const olf_line_params* p);
void orc_line_destroy(orc_line* h);
int orc_lsd_detect(orc_line* h, const uint8_t* img, int w, int hgt, int stride, float* segs, int cap, int* n);
int orc_lsd_stats(const orc_line* h, int* out7);
int orc_keylines_from_segments(const float* segs, int nseg, int w, int h, double min_length, olf_keyline* out, int cap, int* n);
int orc_lbd_compute(orc_line* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, uint8_t* desc);
int orc_lbd_compute_float(orc_line* h, const uint8_t* img, int w, int hgt, int stride, const olf_keyline* kls, int n, float* fdesc);
int orc_line_extract(orc_line* h, const uint8_t* img, int w, int hgt, int stride, olf_keyline* kls, uint8_t* desc, int cap, int* n);

int orc_knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx0, int* dist0, int* idx1, int* dist1);
int orc_match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12, int* nmatches);
int orc_match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int* m12, int* nmatches);
int orc_stereo_points(orc_orb* left, orc_orb* right, const olf_keypoint* kl, const uint8_t* dl, int N,
                      const olf_keypoint* kr, const uint8_t* dr, int Nr, float bf, float fx, float* uRight, float* depth);
int orc_stereo_lines(const olf_keyline* kl, const uint8_t* dl, int n1, const olf_keyline* kr, const uint8_t* dr, int n2,
                     int img_w, int img_h, const olf_line_match_params* P, int* matches12, float* disp, double* le);
int orc_search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches);
int orc_search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur, int* nmatches);
/* MapPoint / MapLine::ComputeDistinctiveDescriptors (src/MapPoint.cc:254-322, src/MapLine.cc:257-322), batched */
int orc_distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best);
/* the remaining ORBmatcher overloads (src/ORBmatcher.cc:292-405, 659-1328, 1620-1747) */
int orc_window_search(const olf_window_search_args* a, int* best_idx, int* best_dist);
int orc_search_for_triangulation(const olf_triangulation_args* a, int* matches12, int* nmatches);
int orc_search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2,
                                  const olf_camera* cam, float* prev_matched, int window_size, float nn_ratio, int check_orientation, int* matches12, int* nmatches);
int orc_search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2, int* matches12, int* nmatches);
/* bag of words (oracle/bow.cpp) */
typedef struct orc_vocab orc_vocab;
orc_vocab* orc_vocab_create(const olf_vocab_desc* v);
void orc_vocab_destroy(orc_vocab* v);
int orc_bow_transform(orc_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id);
int orc_bow_assemble(const int* word_id, const double* weight, const int* node_id, int n, int* bow_word, double* bow_value, int* n_words,
                     int* fv_node, int* fv_begin, int* fv_index, int* n_nodes);
int orc_search_by_bow(const olf_bow_match_args* a, int* match_f, int* nmatches);

/* cv primitive wrappers for the cv2 pinning tests */
void orc_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, double fx, double fy, uint8_t* dst, int* dw, int* dh);
void orc_gaussian_blur(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* dst);
void orc_gauss_kernel_q8(int ksize, double sigma, int* out);
void orc_sobel3(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy);
float orc_fast_atan2(float y, float x);
/* host libm cosf / sinf of n floats (what computeOrbDescriptor calls, src/ORBextractor.cc:113-115) */
void orc_cosf_sinf(const float* x, long n, float* c, float* s);
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int th, int nms, int* xys, int cap);
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, uint8_t* score);
#ifdef __cplusplus
}
#endif
#endif
