/* olf_abi.h -- C-ABI of the B200-native ORB + line stereo front end ("olf" = ORB-Line Front end).
 *
 * This is the drop-in boundary for the ONE hot path of robotseu/ORB_Line_SLAM that this repo
 * accelerates.  The reference has no FFI layer of its own: its boundary is a set of C++
 * class surfaces (ORBextractor, Lineextractor, ORBmatcher, LineMatcher free functions, Frame::Compute*).
 * The C++ shim in orb_line_slam_b200/shim/ re-creates those surfaces with identical signatures and forwards
 * to the entry points below; each entry point cites the reference interface it replaces.
 *
 * Conventions: C linkage, POD in/out, no exceptions.  Return value 0 = OLF_OK, negative = error code.
 * Output buffers are caller-allocated with an explicit capacity; counts are returned through int*.
 * A handle owns its CUDA stream(s), device pyramids and scratch.  Distinct handles may be used
 * concurrently from different host threads; one handle is not re-entrant (same as the reference:
 * ORBextractor is stateful through mvImagePyramid, include/ORBextractor.h:92).
 * Image pointers are HOST pointers unless the function name ends in _dev; pinned (page-locked) host images are read by
 * the device directly, pageable ones go through a pinned staging buffer inside the handle.
 * Nothing CUDA lives in the caller's thread-local storage for the extractors (events belong to the handles): the
 * reference's Frame::Frame spawns fresh std::threads every frame (src/Frame.cc:164-171).
 */
#ifndef OLF_ABI_H
#define OLF_ABI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OLF_OK              0
#define OLF_ERR_ARG        -1   /* bad argument (null pointer, non-positive size, size mismatch)        */
#define OLF_ERR_CUDA       -2   /* CUDA runtime error; olf_last_error() has the text                   */
#define OLF_ERR_CAPACITY   -3   /* caller buffer or internal pool too small                            */
#define OLF_ERR_NO_DEVICE  -4   /* no CUDA device / extension built without one -- never a CPU fallback */
#define OLF_ERR_INTERNAL   -5

#define OLF_DESC_BYTES     32   /* 256-bit rBRIEF / binary LBD descriptor                              */
#define OLF_GRID_COLS      64   /* FRAME_GRID_COLS include/Frame.h:52                                  */
#define OLF_GRID_ROWS      48   /* FRAME_GRID_ROWS include/Frame.h:51                                  */
#define OLF_TH_HIGH       100   /* ORBmatcher::TH_HIGH src/ORBmatcher.cc:39                            */
#define OLF_TH_LOW         50   /* ORBmatcher::TH_LOW  src/ORBmatcher.cc:40                            */
#define OLF_HISTO_LENGTH   30   /* ORBmatcher::HISTO_LENGTH src/ORBmatcher.cc:41                       */
#define OLF_MAX_LEVELS     16
#define OLF_MAX_BATCH_FRAMES 8  /* stereo frames per olf_frontend_process_batch call */

/* cv::KeyPoint fields the reference fills (src/ORBextractor.cc:839-848, 1097-1103); class_id stays -1. */
typedef struct olf_keypoint {
    float x, y;        /* pt, level-0 coordinates                      */
    float size;        /* (int)(31*scale[octave])                      */
    float angle;       /* degrees, cv::fastAtan2                       */
    float response;    /* FAST score                                   */
    int   octave;
} olf_keypoint;

/* cv::line_descriptor::KeyLine, field for field (descriptor_custom.hpp:105-145). */
typedef struct olf_keyline {
    float angle;
    int   class_id;
    int   octave;
    float pt_x, pt_y;
    float response;
    float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int   numOfPixels;
} olf_keyline;

/* Lineextractor ctor arguments (include/LineExtractor.h:43-45) + the Config:: value it reads. */
typedef struct olf_line_params {
    int    lsd_nfeatures;      /* keep top-N by response; 0 = keep all (src/LineExtractor.cc:56) */
    double min_line_length;    /* relative to min(cols,rows)          (src/LineExtractor.cc:53)  */
    int    lsd_refine;         /* only 0 (LSD_REFINE_NONE) is supported: Examples/PL yaml files   */
    double lsd_scale;          /* 1.2 */
    double lsd_sigma_scale;    /* 0.6 */
    double lsd_quant;          /* 2.0 */
    double lsd_ang_th;         /* 22.5 */
    double lsd_log_eps;        /* unused with refine 0 */
    double lsd_density_th;     /* unused with refine 0 */
    int    lsd_n_bins;         /* 1024 */
} olf_line_params;

/* Config:: values read by the line matchers (src/LineMatcher.cpp:106,163,201,262,282; src/Frame.cc:923,954-957,1007,1043) */
typedef struct olf_line_match_params {
    int    best_lr_matches;    /* Config::bestLRMatches()  default 1    */
    double min_ratio_12_l;     /* Config::minRatio12L()    default 0.9  */
    double line_sim_th;        /* Config::lineSimTh()      default 0.75 */
    int    matching_s_ws;      /* Config::matchingSWs()    default 10   */
    double min_disp;           /* Config::minDisp()        default 1.0  */
    double line_horiz_th;      /* Config::lineHorizTh()    default 0.1  */
    double stereo_overlap_th;  /* Config::stereoOverlapTh() default 0.75 */
    double ls_min_disp_ratio;  /* Config::lsMinDispRatio() default 0.7  */
} olf_line_match_params;

/* Pinhole + stereo constants of Frame (src/Frame.cc:182-196) */
typedef struct olf_camera {
    float fx, fy, cx, cy;
    float bf;                  /* Frame::mbf */
    float min_x, max_x, min_y, max_y;   /* Frame::mnMinX.. (image bounds; 0,w,0,h when rectified) */
} olf_camera;

typedef struct olf_orb olf_orb;
typedef struct olf_line olf_line;

const char* olf_last_error(void);
int olf_device_count(void);
/* kernels launched by this library since load (bench.py reports the delta as gpu_launches) */
long long olf_kernel_launch_count(void);
/* device / pinned allocations made by this library since load (each one synchronises the device; steady state makes none) */
long long olf_alloc_count(void);

/* ---- ORBextractor (include/ORBextractor.h:52-118; src/ORBextractor.cc:412-472, 1045-1134) ---------------- */
olf_orb* olf_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int device);
void     olf_orb_destroy(olf_orb* h);
/* ORBextractor::operator() : image -> keypoints + N x 32 descriptors (src/ORBextractor.cc:1045-1107). */
int olf_orb_extract(olf_orb* h, const uint8_t* img, int width, int height, int stride,
                    olf_keypoint* kps, uint8_t* desc, int cap, int* n);
/* same, image already resident on the device (the bench's HBM-resident leg) */
int olf_orb_extract_dev(olf_orb* h, const uint8_t* d_img, int width, int height, int stride,
                        olf_keypoint* kps, uint8_t* desc, int cap, int* n);
/* ORBextractor::mvImagePyramid[level] read-back (src/Frame.cc:709,799,811,816 read it). */
int olf_orb_level_size(const olf_orb* h, int level, int* width, int* height);
int olf_orb_get_level(olf_orb* h, int level, uint8_t* dst, int dst_stride);
/* GetScaleFactors() etc. (include/ORBextractor.h:68-90): out[nlevels] */
int olf_orb_scale_factors(const olf_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
int olf_orb_features_per_level(const olf_orb* h, int* out);
/* debug/parity: FAST candidates of the last extract, before the quadtree (src/ORBextractor.cc:791-831):
 * per candidate {level, x, y, score} with x,y relative to the level image. */
int olf_orb_last_candidates(olf_orb* h, int* lvl_x_y_score /*cap x 4*/, int cap, int* n);

/* parity hook: the device evaluates the per-keypoint cosf / sinf of rBRIEF (src/ORBextractor.cc:113-115) with glibc's own
 * algorithm (csrc/sincosf_exact.h); cos_sin[2*i], cos_sin[2*i+1] = cos, sin of the float with bit pattern first_bits + i*stride */
int olf_trig_sweep(unsigned first_bits, unsigned stride, unsigned count, float* cos_sin, int device);

/* ---- Lineextractor (include/LineExtractor.h:40-72; src/LineExtractor.cc:31-67) ---------------------------- */
olf_line* olf_line_create(const olf_line_params* p, int device);
void      olf_line_destroy(olf_line* h);
/* Lineextractor::operator(): LSD detect -> top-N by response -> LBD (src/LineExtractor.cc:54-66). */
int olf_line_extract(olf_line* h, const uint8_t* img, int width, int height, int stride,
                     olf_keyline* kls, uint8_t* desc, int cap, int* n);
int olf_line_extract_dev(olf_line* h, const uint8_t* d_img, int width, int height, int stride,
                         olf_keyline* kls, uint8_t* desc, int cap, int* n);
/* cv::LineSegmentDetector::detect as configured by LSDDetectorC::detectImpl (LSDDetector_custom.cpp:246-262):
 * raw Vec4f segments (x1,y1,x2,y2) in seed order, before the KeyLine filter. */
int olf_lsd_detect(olf_line* h, const uint8_t* img, int width, int height, int stride,
                   float* segs /*cap x 4*/, int cap, int* n);
/* BinaryDescriptor::compute on caller-supplied keylines (binary_descriptor_custom.cpp:539-687). */
int olf_lbd_compute(olf_line* h, const uint8_t* img, int width, int height, int stride,
                    const olf_keyline* kls, int n, uint8_t* desc);

/* ---- Hamming matchers --------------------------------------------------------------------------------- */
/* cv::BFMatcher(NORM_HAMMING).knnMatch(d1,d2,2) as used by matchNNR (src/LineMatcher.cpp:47-49):
 * two nearest train rows per query, ties -> lowest train index.  idx1/dist1 = -1/INT_MAX-like 256*2 when n2 < 2. */
int olf_knn2_hamming(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                     int* idx0, int* dist0, int* idx1, int* dist1, int device);
/* measurement hook (BASELINE.json configs[4], the descriptor-match micro-benchmark): kernel-only time of the knn2 kernels on
 * device-resident random descriptors, and the chip's measured xor+popc+add ceiling in 32-bit word pairs per second */
int olf_knn2_bench(int n1, int n2, int iters, int device, double* kernel_ms, double* popc_word_pairs_per_s);
/* matchNNR (src/LineMatcher.cpp:42-62): matches12[n1], returns count through *nmatches. */
int olf_match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* matches12, int* nmatches, int device);
/* match(desc1,desc2,nnr,matches12) (src/LineMatcher.cpp:104-132): adds the mutual check when best_lr_matches. */
int olf_match_lines(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr_matches,
                    int* matches12, int* nmatches, int device);

/* ---- Frame::ComputeStereoMatches (src/Frame.cc:702-876) -------------------------------------------------- */
/* Needs both extractors' pyramids (mvImagePyramid) which stay on the device inside the handles. */
int olf_stereo_points(olf_orb* left, olf_orb* right,
                      const olf_keypoint* kps_l, const uint8_t* desc_l, int n_l,
                      const olf_keypoint* kps_r, const uint8_t* desc_r, int n_r,
                      float bf, float fx, float* u_right /*n_l*/, float* depth /*n_l*/);

/* ---- Frame::ComputeStereoMatches_Lines + matchGrid(lines) (src/Frame.cc:878-1000; src/LineMatcher.cpp:220-299) */
int olf_stereo_lines(const olf_keyline* kls_l, const uint8_t* desc_l, int n_l,
                     const olf_keyline* kls_r, const uint8_t* desc_r, int n_r,
                     int img_width, int img_height, const olf_line_match_params* p,
                     int* matches12 /*n_l, raw matchGrid output*/,
                     float* disp_s_e /*n_l x 2, mvDisparity_l*/, double* le /*n_l x 3, mvle_l*/, int device);

/* ---- matchGrid(lines1, desc1, grid, desc2, directions2, w, matches_12) (include/LineMatcher.h:69, src/LineMatcher.cpp:220-299) for
 * callers that build the GridStructure themselves (Frame::ComputeStereoMatches_Lines, src/Frame.cc:896-927, unchanged).
 * grid: the cells of GridStructure (include/gridStructure.h:40-58) as CSR, cell (x, y) at index x*rows + y; lines1: n1 x 4 grid
 * coordinates (start x, y, end x, y); dir2: n2 x 2 normalised directions; window: GridWindow {width.first, width.second,
 * height.first, height.second}.  Candidates are walked in ascending index (canonical order of the hash set, Appendix C.2). */
typedef struct olf_grid_csr { int rows, cols; const int* cell_begin; /* rows*cols + 1 */ const int* items; } olf_grid_csr;
int olf_match_grid_lines(const int* lines1, const uint8_t* desc1, int n1, const olf_grid_csr* grid, const uint8_t* desc2, int n2,
                         const double* dir2, const int* window, const olf_line_match_params* p, int* matches12, int* nmatches, int device);

/* ---- ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono, match12) (src/ORBmatcher.cc:1474-1618) - */
typedef struct olf_sbp_last_args {
    /* current frame (the one being filled) */
    const olf_keypoint* cur_kps; const uint8_t* cur_desc; const float* cur_u_right; int n_cur;
    olf_camera cam;
    const float* scale_factors; int nlevels;
    float Rcw[9]; float tcw[3];          /* CurrentFrame.mTcw */
    float Rlw[9]; float tlw[3];          /* LastFrame.mTcw    */
    /* last frame: one entry per keypoint i */
    const olf_keypoint* last_kps; int n_last;
    const uint8_t* last_has_point;       /* mvpMapPoints[i] != NULL && !mvbOutlier[i]       */
    const uint8_t* last_point_observed;  /* pMP->Observations() > 0                         */
    const float*   last_world_pos;       /* n_last x 3, pMP->GetWorldPos()                  */
    const uint8_t* last_point_desc;      /* n_last x 32, pMP->GetDescriptor()               */
    float th; int mono; int check_orientation;
    const uint8_t* cur_occupied;         /* optional (may be NULL): CurrentFrame.mvpMapPoints[j] && Observations()>0 on entry (:1550-1552) */
} olf_sbp_last_args;
/* assigned_cur[i] (n_last) = index of the current keypoint map point i was written to (before the
 * rotation-consistency pass) or -1; cur_point[j] (n_cur) = final index into last-frame points held by
 * current keypoint j or -1 (after ComputeThreeMaxima pruning); *nmatches = return value of the reference. */
int olf_search_by_projection_last(const olf_sbp_last_args* a, int* assigned_cur, int* cur_point, int* nmatches, int device);

/* ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:47-131) ------- */
typedef struct olf_sbp_map_args {
    const olf_keypoint* cur_kps; const uint8_t* cur_desc; const float* cur_u_right; int n_cur;
    const uint8_t* cur_occupied;         /* F.mvpMapPoints[idx] && Observations()>0 on entry */
    olf_camera cam;
    const float* scale_factors; int nlevels;
    int n_points;                        /* map points with mbTrackInView && !isBad()        */
    const float* proj_x; const float* proj_y; const float* proj_xr;   /* mTrackProjX/Y/XR     */
    const int*   pred_level;             /* mnTrackScaleLevel                                */
    const float* view_cos;               /* mTrackViewCos                                    */
    const uint8_t* point_observed;       /* Observations()>0 (so that a write blocks later)  */
    const uint8_t* point_desc;           /* n_points x 32                                    */
    float th; float nn_ratio;
} olf_sbp_map_args;
int olf_search_by_projection_map(const olf_sbp_map_args* a, int* assigned_cur /*n_points*/, int* nmatches, int device);

/* ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:254-322) and MapLine::ComputeDistinctiveDescriptors
 * (src/MapLine.cc:257-322), batched: landmark g owns the observed descriptors desc[group_begin[g] .. group_begin[g+1]) (the rows
 * pKF->mDescriptors.row(idx) of its non-bad observations, in the order of its observation map); best[g] = index inside the group
 * of the descriptor with the least median Hamming distance to all of them (first on ties), -1 for an empty group. */
int olf_distinctive_descriptors(const uint8_t* desc, const int* group_begin, int n_groups, int* best, int device);

/* ---- next row: bag of words -- Frame::ComputeBoW (src/Frame.cc:585-597) and ORBmatcher::SearchByBoW -------- */
/* DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) flattened: node 0 is the
 * root; the children of node i are children[child_begin[i] .. child_begin[i] + child_count[i]) in the order of
 * m_nodes[i].children; a node without children is a leaf = a word.  The same structure serves the line vocabulary
 * (LBD descriptors are 32 bytes too). */
typedef struct olf_vocab_desc {
    int k, L, n_nodes;
    const uint8_t* node_desc;              /* n_nodes x 32  m_nodes[i].descriptor (root: unused)              */
    const int*     child_begin;            /* n_nodes                                                         */
    const int*     child_count;            /* n_nodes       0 = leaf                                          */
    const int*     children;               /* concatenated child node ids                                     */
    const int*     word_id;                /* n_nodes       m_nodes[i].word_id of a leaf                      */
    const double*  weight;                 /* n_nodes       m_nodes[i].weight of a leaf (0 = stopped word)    */
} olf_vocab_desc;
typedef struct olf_vocab olf_vocab;
olf_vocab* olf_vocab_create(const olf_vocab_desc* v, int device);      /* copies the tree to the device */
void olf_vocab_destroy(olf_vocab* v);
/* TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup) (TemplatedVocabulary.h:1218-1260) for n
 * descriptors: tree descent by Hamming distance (FORB::distance, first minimum wins), per feature its word, the word's
 * weight and its ancestor `levelsup` levels above the leaves (levelsup >= L: the root, 0).  The O(n) assembly of
 * BowVector / FeatureVector (std::map order, double sums in feature order, L1 normalisation -- transform(features, v, fv,
 * levelsup) :1131-1194 with TF_IDF + L1_NORM) stays with the caller: olf_bow_assemble does it on the host. */
int olf_bow_transform(olf_vocab* v, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id);
/* BowVector as (word ascending, value) and FeatureVector as CSR over node ids ascending; capacities >= n.
 * n_words / n_nodes: entries written. */
int olf_bow_assemble(const int* word_id, const double* weight, const int* node_id, int n,
                     int* bow_word, double* bow_value, int* n_words,
                     int* fv_node, int* fv_begin /* n_nodes + 1 */, int* fv_index /* n */, int* n_nodes);
/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:161-290): for every vocabulary node the
 * key frame and the frame share, each key-frame feature with a good map point takes the closest still-unmatched frame
 * feature of that node (best <= TH_LOW, best < nn_ratio * second), then the rotation-histogram filter. */
typedef struct olf_bow_match_args {
    const uint8_t* kf_desc; const olf_keypoint* kf_kps_un; int n_kf;
    const uint8_t* kf_has_point;           /* n_kf  pMP != NULL && !pMP->isBad()                                */
    const int* kf_fv_node; const int* kf_fv_begin; const int* kf_fv_index; int kf_n_nodes;     /* pKF->mFeatVec */
    const uint8_t* f_desc; const olf_keypoint* f_kps; int n_f;
    const int* f_fv_node; const int* f_fv_begin; const int* f_fv_index; int f_n_nodes;         /* F.mFeatVec    */
    float nn_ratio; int check_orientation;
} olf_bow_match_args;
int olf_search_by_bow(const olf_bow_match_args* a, int* match_f /* n_f: key-frame feature index or -1 */, int* nmatches, int device);

/* ---- the remaining ORBmatcher overloads (LocalMapping / LoopClosing / relocalisation) ------------------ */
/* Grid-window search shared by Fuse(KF, MPs, th) (src/ORBmatcher.cc:827-977), Fuse(KF, Scw, MPs, th, replace) (:979-1102),
 * SearchByProjection(KF, Scw, MPs, matched, th) (:292-405), both directions of SearchBySim3 (:1104-1328) and
 * SearchByProjection(Frame, KF, found, th, ORBdist) (:1620-1747).  The caller runs the per-point prologue of the overload
 * (projection with cv::Mat float arithmetic, depth / viewing-angle tests, MapPoint::PredictScale -- O(#points) host geometry)
 * and hands over one query per surviving point; the library does what is data-parallel: KeyFrame::GetFeaturesInArea
 * (src/KeyFrame.cc:747-786; src/Frame.cc:517-570 is the same window) over the 64 x 48 grid in the reference's cell-major
 * order, the octave window, the optional chi-square reprojection gate of Fuse (:916-940), the Hamming distances, the FIRST
 * minimum (`dist < bestDist`), and -- for the two overloads that write into the searched frame while they loop -- the rule
 * that a keypoint taken by an earlier query is no longer a candidate (:360,:397; :1691,:1705).
 * best_idx[i] = keypoint of query i, or -1 when the best admissible distance exceeds max_dist (or there is no candidate);
 * best_dist[i] = its distance (256 when none). */
typedef struct olf_window_search_args {
    const olf_keypoint* kps; const uint8_t* desc; int n;      /* searched keypoints: mvKeysUn, mDescriptors                      */
    olf_camera cam;                                            /* grid geometry mnMinX.., (bf unused)                             */
    const float* u_right;                                      /* n, mvuRight: chi2_check only (may be NULL otherwise)            */
    const float* inv_level_sigma2; int nlevels;                /* mvInvLevelSigma2: chi2_check only                               */
    int n_queries;
    const float* u; const float* v; const float* radius;       /* projection and th * mvScaleFactors[nPredictedLevel]             */
    const float* ur;                                           /* u - bf * invz: chi2_check only                                  */
    const int* min_level; const int* max_level;                /* octave window of the candidate loop                             */
    const uint8_t* qdesc;                                      /* n_queries x 32, pMP->GetDescriptor()                            */
    const uint8_t* blocked;                                    /* n (may be NULL): vpMatched[idx] / mvpMapPoints[i2] on entry     */
    int sequential_blocking;                                   /* 1: an accepted keypoint blocks the queries after it             */
    int chi2_check;                                            /* 1: Fuse(KF, MPs, th) reprojection gate (7.8 stereo / 5.99 mono) */
    int max_dist;                                              /* TH_LOW, TH_HIGH or ORBdist                                      */
} olf_window_search_args;
int olf_window_search(const olf_window_search_args* a, int* best_idx, int* best_dist, int device);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:659-825): for every
 * vocabulary node both key frames share, each feature of KF1 without a map point takes the feature of KF2 (without a map
 * point) at minimal Hamming distance <= TH_LOW that passes the epipole and CheckDistEpipolarLine tests (:142-159) -- the LAST
 * one among equals (`dist > bestDist` rejects) -- then the rotation-histogram filter.  matches12[i1] = i2 or -1. */
typedef struct olf_triangulation_args {
    const olf_keypoint* kps1; const uint8_t* desc1; int n1;
    const uint8_t* skip1;                  /* n1: GetMapPoint(idx1) != NULL                                        */
    const float* u_right1;                 /* n1: mvuRight (stereo iff >= 0)                                       */
    const int* fv1_node; const int* fv1_begin; const int* fv1_index; int fv1_n_nodes;
    const olf_keypoint* kps2; const uint8_t* desc2; int n2;
    const uint8_t* skip2; const float* u_right2;
    const int* fv2_node; const int* fv2_begin; const int* fv2_index; int fv2_n_nodes;
    const float* scale_factors2; const float* level_sigma2_2; int nlevels;     /* pKF2->mvScaleFactors, mvLevelSigma2 */
    float F12[9];                          /* row major                                                            */
    float ex, ey;                          /* epipole in the second image (:666-672, computed by the caller)       */
    int only_stereo; int check_orientation;
} olf_triangulation_args;
int olf_search_for_triangulation(const olf_triangulation_args* a, int* matches12 /* n1 */, int* nmatches, int device);

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:524-657): as olf_search_by_bow, but both sides
 * need a good map point, the acceptance is `best < TH_LOW` (strict) and the result is indexed by the first key frame.
 * The argument block is the one of olf_search_by_bow (kf = pKF1, f = pKF2, f_kps = pKF2->mvKeysUn) plus has_point2. */
int olf_search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2 /* n_f */, int* matches12 /* n_kf */, int* nmatches, int device);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:407-522; monocular map initialisation):
 * every level-0 keypoint of F1 searches the window of `window_size` pixels around prev_matched[i1] among the level-0 keypoints of F2; a
 * keypoint of F2 can be taken over by a later, closer keypoint of F1 (vMatchedDistance / vnMatches21), then the ratio test and the rotation
 * histogram.  The window x Hamming candidate lists come from the device; the order-dependent take-over runs on the host in keypoint order.
 * matches12[n1]; prev_matched (n1 x 2, x y) is updated in place for the surviving matches (:516-518). */
int olf_search_for_initialization(const olf_keypoint* kps1, const uint8_t* desc1, int n1, const olf_keypoint* kps2, const uint8_t* desc2, int n2,
                                  const olf_camera* cam, float* prev_matched, int window_size, float nn_ratio, int check_orientation,
                                  int* matches12, int* nmatches, int device);

/* ---- whole stereo frame: Frame::Frame(stereo+lines) (src/Frame.cc:136-221) ----------------------------------- */
/* One rig = 2 ORB extractors + 2 line extractors on one device; olf_frontend_process runs ExtractORB(L|R) and
 * ExtractLine(L|R) on its own host threads (src/Frame.cc:164-171), then ComputeStereoMatches and
 * ComputeStereoMatches_Lines, and fills ONE fixed-capacity POD block (the unit the multi-GPU driver gathers). */
typedef struct olf_frontend_params {
    int   nfeatures; float scale_factor; int nlevels; int ini_th_fast; int min_th_fast;   /* ORBextractor ctor */
    int   has_lines;                       /* Config::hasLines() (src/LineExtractor.cc:37, src/Frame.cc:203) */
    olf_line_params line;
    olf_line_match_params line_match;
    olf_camera cam;
    int   cap_points;                      /* capacity per eye (>= nfeatures + 64: the quadtree may exceed nfeatures) */
    int   cap_lines;
} olf_frontend_params;

typedef struct olf_frame_header {          /* first bytes of the result block */
    int n_l, n_r;                          /* keypoints left / right  (Frame::N)        */
    int m_l, m_r;                          /* keylines  left / right  (Frame::N_l)      */
    int cap_points, cap_lines;
    int status;                            /* OLF_OK or the first error of the frame    */
    int reserved;
} olf_frame_header;
/* Block layout after the header (all offsets 64-byte aligned, see olf_frame_offsets):
 *  kps_l[cap_p] desc_l[cap_p*32] kps_r[cap_p] desc_r[cap_p*32] u_right[cap_p] depth[cap_p]
 *  kls_l[cap_l] ldesc_l[cap_l*32] kls_r[cap_l] ldesc_r[cap_l*32] lmatch[cap_l] ldisp[cap_l*2] lle[cap_l*3 doubles] */
typedef struct olf_frame_offsets {
    uint64_t kps_l, desc_l, kps_r, desc_r, u_right, depth, kls_l, ldesc_l, kls_r, ldesc_r, lmatch, ldisp, lle, total;
} olf_frame_offsets;
typedef struct olf_frontend olf_frontend;
int  olf_frame_layout(int cap_points, int cap_lines, olf_frame_offsets* out);
olf_frontend* olf_frontend_create(const olf_frontend_params* p, int device);
void olf_frontend_destroy(olf_frontend* h);
/* images: host pointers (on_device = 0) or device pointers (on_device = 1); result: host block of layout.total bytes */
int  olf_frontend_process(olf_frontend* h, const uint8_t* img_l, const uint8_t* img_r, int width, int height, int stride,
                          int on_device, void* result);
/* Batch entry (SURVEY 8b `olf_frame_batch_extract`, the unit of work of the multi-GPU driver, BASELINE config "8-frame
 * batch"): `nframes` INDEPENDENT stereo frames (1..max_frames <= OLF_MAX_BATCH_FRAMES) through Frame::Frame at once.  The line extraction of
 * all 2*nframes images runs as ONE chain of kernel launches (the LSD passes are latency-bound: a batch costs the latency of
 * one image), so a rig keeps 2*max_frames images in flight on ONE CUDA stream (the short ORB + stereo chain first, the line chain behind it).  Results are identical to nframes calls of
 * olf_frontend_process; results[f] is the block of frame f. */
olf_frontend* olf_frontend_create_batch(const olf_frontend_params* p, int device, int max_frames);
int  olf_frontend_process_batch(olf_frontend* h, const uint8_t* const* img_l, const uint8_t* const* img_r, int nframes,
                                int width, int height, int stride, int on_device, void* const* results);
/* the extractors of the rig, slot = 2 * frame_in_batch + eye (e.g. for olf_orb_get_level) */
olf_orb*  olf_frontend_orb(olf_frontend* h, int slot);
olf_line* olf_frontend_line(olf_frontend* h, int slot);
/* host view of the last olf_frontend_process(_batch) call, microseconds: [0] ORB + stereo-point enqueue, [1] line extraction (blocking),
 * [2] wait for the ORB chain, [3] copy-out, [4] stereo lines, [5] whole call */
int  olf_frontend_last_timing(const olf_frontend* h, int* out8);
/* last-call statistics of a line extractor: out[0] LSD rounds, out[1] waves, out[2] accepted regions, out[3] device time of
 * the region-growing chain in microseconds (CUDA events on its stream), out[4] images that shared the chain, out[5] host microseconds until the chain had finished (enqueue + wait), out[6] rectangle trig + second rectangle pass, out[7] KeyLine construction + LBD */
int  olf_line_last_stats(const olf_line* h, int* out8);

#ifdef __cplusplus
}
#endif
#endif /* OLF_ABI_H */
