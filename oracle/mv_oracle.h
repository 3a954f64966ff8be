/* mv_oracle.h -- CPU restatement ("T2") of the maveric-slam tracking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under maveric-slam_b200/ may include, link or
 * call this; it is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
 *
 * Parity status
 *   - detector post-processing, windowed matcher, RANSAC-E, pose-from-E, matmul shim:
 *     PINNED.  tests/test_oracle_vs_ref.py asserts byte equality against the reference
 *     sources themselves, compiled unmodified from /root/reference into
 *     oracle/_ref/libmaveric_ref.so (oracle/Makefile), on the reference's native
 *     24x80 shape over many seeds, and against the reference's own fixtures
 *     (tests/golden/, made by tests/golden/make_golden.py).
 *   - Gauss-Newton PnP (orc_pnp_*): PARITY UNPINNED.  The reference contains no
 *     Gauss-Newton PnP (src/pnp_solver.c is an essential-matrix RANSAC with E forced
 *     to identity; src/projection_factor.c:27-33 has only the residual).  The oracle
 *     restates the residual and the SE3 convention of the reference and is otherwise
 *     this repository's own algorithm; it is cross-checked against cv2.solvePnP in
 *     tests/test_oracle_pnp.py.
 *
 * Build: gcc -std=gnu11 -O2 -fwrapv -ffp-contract=off  (see oracle/Makefile).
 */
#ifndef MV_ORACLE_H
#define MV_ORACLE_H
#include <stdint.h>
#include <stddef.h>

/* ---- detector post-processing (src/top_N.c) ---- */
int  orc_softmax(float scale, const int8_t* semi, int cells, int* max_idx, float* probs);
/* src/run_nms.c:65-156: 2x2-quadrant NMS over the per-cell keypoints, in place.  A suppressed
 * cell gets max_idx = 64 and prob = 64.0f (run_nms.c:138-139).  events (nullable,
 * [max_events][4]) receives (x_keep, y_keep, x_suppressed, y_suppressed) per suppression in the
 * reference's order; returns the number of suppressions. */
int  orc_nms(int rows, int cols, int* max_idx, float* probs, int* events, int max_events);
int  orc_top_n(float scale, const int8_t* semi, int cells, int N, int max_valid,
               int* num_selected, int* patches, int* indices, float* probs);

/* ---- windowed matcher (src/tracking_main.c:18-43,103-194) ---- */
typedef struct {
  int rows, cols, shift_x, shift_y, radius, max_matches;
  double match_threshold, min_prob0;
} orc_match_cfg;
typedef struct {
  long long window_cells;  /* cells visited (valid or not) */
  long long pairs_256;     /* descriptor pairs evaluated over 256 dims */
  long long pairs_64;      /* ... over 64 dims */
} orc_match_stats;
int orc_match(const orc_match_cfg* cfg, const int8_t* desc0, const int8_t* desc1,
              const int* max_idx0, const float* probs0,
              int nq, const int* patches1, const int* indices1,
              float* pts0, float* pts1, int* cell0, int* query, float* score,
              orc_match_stats* stats);

/* ---- pose: src/pnp_solver.c + include/svd/svd.h ---- */
void  orc_svd3(const float A[3][3], float U[3][3], float S[3][3], float V[3][3]);
float orc_reproj_error(const float p1[2], const float p2[2], const float E[3][3]);
void  orc_normalize_points(int n, const float* pts, const float K[3][3], float* out);
int   orc_ransac_identity(int n, const float* pts1, const float* pts2, int iters, float thr,
                          int max_inliers_cap, float best_E[3][3], int* best_inliers,
                          int* num_inliers);
void  orc_recover_pose(const float E[3][3], float R1[3][3], float R2[3][3], float t[3]);

/* ---- geometry + residual (src/types.c, src/projection_factor.c) ---- */
void orc_quat_mul(const float a[4], const float b[4], float out[4]);
void orc_apply_transform(const float pose[7], const float X[3], float out[3]);
void orc_projection_error(const float pose[7], const float X[3], const float z[2],
                          const float cam[4], float err[2]);

/* ---- matmul shim (include/gemmini_functions_cpu.h) ---- */
void orc_matmul(size_t I, size_t J, size_t K, const float* A, const float* B, float* C,
                size_t sA, size_t sB, size_t sC, float a_scale, float b_scale, int tA, int tB);
void orc_matmul2(size_t I, size_t J, size_t K, const float* A, const float* B,
                 const float* D, float* C, size_t sA, size_t sB, size_t sD, size_t sC,
                 float a_scale, float b_scale, float d_scale, int tA, int tB);

/* ---- local bundle adjustment: Schur complement (src/local_bundle_adjustment.c:133-246) ---- */
void orc_invert_3x3(float* matrix, int stride);
void orc_lba_schur(int n_ldmks, int n_poses, int chunk, const float* J, float* C);
/* damped Cholesky step of the reduced camera system (the reference's cholesky() is a stub:
 * this repository's definition, parity unpinned); 1 = solved, 0 = not positive definite, d = 0 */
int orc_lba_solve(int n_poses, float damping, const float* C, float* d);

/* ---- Gauss-Newton PnP RANSAC (parity unpinned; this repository's definition) ---- */
typedef struct {
  float fx, fy, cx, cy;
  int hypotheses, sample_size, sample_iters, refine_iters;
  float gate_sq, min_depth, damping;
  uint64_t seed;
  int lanes;   /* 1: sequential sums; 2..32 (power of two): lane-strided partials + xor butterfly */
} orc_pnp_cfg;
/* corr: SoA planes X,Y,Z,u,v each `stride` floats.  out_pose[7], out_stats[4] =
 * {inliers,cost,best_h,valid}; hyp_pose (nullable) [H][8]. */
void orc_pnp_gn(const orc_pnp_cfg* cfg, int pair_index, int n, int stride, const float* corr,
                const float* init_pose, float* out_pose, float* out_stats, float* hyp_pose);

/* ---- synthetic frames (same counter-based generator as the CUDA/numpy ones) ---- */
typedef struct { uint64_t seed; int rows, cols, keypoint_permille, noise_amp; } orc_synth_cfg;
void orc_synth_frame(const orc_synth_cfg* cfg, int frame, int off_x, int off_y,
                     int8_t* semi, int8_t* desc, float* depth);

/* ---- whole path for one pair, as the timed CPU baseline ---- */
typedef struct {
  orc_match_cfg match; orc_pnp_cfg pnp;
  int top_n, max_valid, ransac_iters; float ransac_thr;
  float semi_scale;
} orc_track_cfg;
typedef struct {
  float q[4], t[3], pnp_inliers, pnp_cost;
  int num_matches, ransac_inliers, best_h, status;
} orc_pair_result;
/* scratch-free convenience: frames f0 (candidates) and f1 (queries). */
void orc_track_pair(const orc_track_cfg* cfg, int pair_index,
                    const int8_t* semi0, const int8_t* desc0, const float* depth0,
                    const int8_t* semi1, const int8_t* desc1,
                    orc_pair_result* out);
/* Runs pairs [first, first+n) of a synthetic sequence on `threads` OpenMP threads;
 * returns elapsed seconds of the tracking region only (generation excluded). */
double orc_bench_sequence(const orc_track_cfg* cfg, const orc_synth_cfg* syn,
                          const int* offsets /*[(n+1)][2]*/, int first, int n, int threads,
                          orc_pair_result* out);
#endif

/* ---- BoW word assignment (src/bow_main.c:62-125).  PARITY UNPINNED as a whole: the
 * reference program crashes as shipped and every stage reads memory as the wrong type (int8 arrays through
 * the float matmul shim :81-86, an int8 row as int* :105, 8 ints of a 4-int leaf :115).  What is pinned:
 * orc_bow_binarize / orc_bow_matching_bits equal the reference's own get_binary_descriptor (:13-41, fed the
 * descriptor widened to int) and count_matching_bits (:43-55) bit for bit (tests/test_bow.py).  The stated
 * definition of the rest:
 *   raw_j   = sum_k desc[k] * base[k][j]                      int32, exact (the matmul of :81-86)
 *   m_j     = sat_int8(rint(desc_scale * raw_j / 256))        ("mvout scale should be 1/256", int8 scores[][])
 *   score_j = scale_arr[j] * m_j + 256 * bias_arr[j]          fp32, unfused (:94)
 *   base    = first j whose score exceeds every earlier one and 0 (:92-101: strict >, max starts at 0, node 0)
 *   bits    = get_binary_descriptor over the 256 elements, 8 words, MSB first (:13-41)
 *   wid     = first leaf with the most matching bits, count_matching_bits over 8 words starting at
 *             leaf_descriptors[base][wid][0] in the FLAT array (:110-120 reads 8 ints of a 4-int leaf, i.e.
 *             the leaf and its successor; the 4 words after the last leaf are defined as 0)               */
typedef struct {
  int n_base, words_per_base;
  const int8_t* base_desc;   /* [256][n_base]  (vocabulary.h:11) */
  const float* scale;        /* [n_base]       (vocabulary.h:7)  */
  const float* bias;         /* [n_base]       (vocabulary.h:9)  */
  const int* leaves;         /* [n_base * words_per_base * 4 + 4], the last 4 zero (vocabulary.h:272) */
} orc_bow_vocab;
void orc_bow_binarize(float scale, const int8_t* feature, int* binary8);
int  orc_bow_matching_bits(const int* a, const int* b, int size);
void orc_bow_assign(const orc_bow_vocab* v, float desc_scale, const int8_t* desc, int* base, int* wid);

/* ---- landmark table: the contents of the reference's local feature pool (include/local_feature_pool.h)
 * after the per-frame sequence of local_feature_matching.c:153-163, as a map word id -> LocalFeature.  */
typedef struct { int word_id, frame_ptr, num_frames, frames[8]; float coords[3]; } orc_local_feature;
void orc_pool_observe(orc_local_feature* table, int n_words, int frame, int n, const int* word_ids, const float* coords);
void orc_pool_remove_old(orc_local_feature* table, int n_words, int current_frame);
