/* ref_harness.c -- glue that lets tests drive the UNMODIFIED reference sources
 * ("T1").  TEST INFRASTRUCTURE ONLY; built by oracle/Makefile into
 * oracle/_ref/libmaveric_ref.so together with /root/reference/src/{tracking_main,
 * pnp_solver,top_N,types,projection_factor}.c compiled where they lie.
 *
 * How the reference's main() is observed without patching it:
 *   - tracking_main.c is compiled with -Dmain=ref_tracking_main and reads its input
 *     from the globals declared in ref_shim/quantized_pair0.h, defined here;
 *   - pnp_solver.c is compiled with -Dransac_essential_matrix=ref_ransac_essential_
 *     matrix, so main()'s call lands in the capture function below, which records
 *     the match list (tracking_main.c:212) and forwards to the real routine with an
 *     inlier buffer that is large enough (main()'s own is int[10], :201);
 *   - printf is renamed to ref_printf in those two files: call_svd (pnp_solver.c:
 *     8-12) prints from inside the solver.
 */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>

#include "quantized_pair0.h"

int cell_size = 8;
int image0_rows = 192, image0_cols = 640, image0_channels = 1;
int image0_feature_rows = 24, image0_feature_cols = 80;
float image0_semi_scale, image0_desc_scale;
int8_t image0_semi[1920][65], image0_desc[1920][256];
int image1_rows = 192, image1_cols = 640, image1_channels = 1;
int image1_feature_rows = 24, image1_feature_cols = 80;
float image1_semi_scale, image1_desc_scale;
int8_t image1_semi[1920][65], image1_desc[1920][256];

static int g_verbose = 0;
int ref_printf(const char* fmt, ...) {
  if (!g_verbose) return 0;
  va_list ap;
  va_start(ap, fmt);
  int r = vprintf(fmt, ap);
  va_end(ap);
  return r;
}
void ref_set_verbose(int v) { g_verbose = v; }

/* captured at the RANSAC call */
static int g_num_matches;
static float g_pts1[150][2], g_pts2[150][2];
static float g_best_E[3][3];
static int g_inliers[1000], g_num_inliers;

void ref_ransac_essential_matrix(const int num_points, const float points1[][2],
                                 const float points2[][2], const float K[3][3],
                                 const int num_iterations, const float inlier_threshold,
                                 float best_E[3][3], int* best_inliers, int* num_inliers);
void recover_pose_from_essential_matrix(float E[3][3], float R1[3][3], float R2[3][3], float t[3]);

void ransac_essential_matrix(const int num_points, const float points1[][2],
                             const float points2[][2], const float K[3][3],
                             const int num_iterations, const float inlier_threshold,
                             float best_E[3][3], int* best_inliers, int* num_inliers) {
  g_num_matches = num_points;
  memcpy(g_pts1, points1, sizeof(float) * 2 * (size_t)num_points);
  memcpy(g_pts2, points2, sizeof(float) * 2 * (size_t)num_points);
  g_num_inliers = 0;
  /* defined result where the reference has none (no inlier / no match) */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) g_best_E[i][j] = (i == j);
  if (num_points > 0)
    ref_ransac_essential_matrix(num_points, points1, points2, K, num_iterations,
                                inlier_threshold, g_best_E, g_inliers, &g_num_inliers);
  memcpy(best_E, g_best_E, sizeof(g_best_E));
  *num_inliers = g_num_inliers;
  for (int i = 0; i < g_num_inliers && i < 10; i++) best_inliers[i] = g_inliers[i];
}

int ref_tracking_main(void);

void ref_load_pair(float semi_scale0, const int8_t* semi0, float desc_scale0, const int8_t* desc0,
                   float semi_scale1, const int8_t* semi1, float desc_scale1, const int8_t* desc1) {
  image0_semi_scale = semi_scale0; image0_desc_scale = desc_scale0;
  image1_semi_scale = semi_scale1; image1_desc_scale = desc_scale1;
  memcpy(image0_semi, semi0, sizeof(image0_semi));
  memcpy(image0_desc, desc0, sizeof(image0_desc));
  memcpy(image1_semi, semi1, sizeof(image1_semi));
  memcpy(image1_desc, desc1, sizeof(image1_desc));
}

/* Runs the reference main(); returns the number of matches and copies what it fed
 * to / got from the pose stage. */
int ref_run_tracking(float* pts1, float* pts2, int* num_inliers, int* inliers, float* best_E) {
  g_num_matches = 0;
  ref_tracking_main();
  memcpy(pts1, g_pts1, sizeof(float) * 2 * (size_t)g_num_matches);
  memcpy(pts2, g_pts2, sizeof(float) * 2 * (size_t)g_num_matches);
  if (num_inliers) *num_inliers = g_num_inliers;
  if (inliers) memcpy(inliers, g_inliers, sizeof(int) * (size_t)g_num_inliers);
  if (best_E) memcpy(best_E, g_best_E, sizeof(g_best_E));
  return g_num_matches;
}

/* ---- src/run_nms.c: main() renamed to ref_nms_main, its printf to ref_nms_printf, and its
 * three helper names that clash with tracking_main.c's prefixed with nms_ (oracle/Makefile).
 * The driver reports only through stdout: "(x y) suppressing (x y)" per suppression (:141) and
 * "x y" per surviving keypoint (:172). */
static int g_nms_events[4096][4], g_nms_n_events, g_nms_kp[1920][2], g_nms_n_kp;
int ref_nms_printf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  if (fmt[0] == '(') {
    int a = va_arg(ap, int), b = va_arg(ap, int), c = va_arg(ap, int), d = va_arg(ap, int);
    if (g_nms_n_events < 4096) {
      g_nms_events[g_nms_n_events][0] = a; g_nms_events[g_nms_n_events][1] = b;
      g_nms_events[g_nms_n_events][2] = c; g_nms_events[g_nms_n_events][3] = d;
    }
    g_nms_n_events++;
  } else {
    int a = va_arg(ap, int), b = va_arg(ap, int);
    if (g_nms_n_kp < 1920) { g_nms_kp[g_nms_n_kp][0] = a; g_nms_kp[g_nms_n_kp][1] = b; }
    g_nms_n_kp++;
  }
  va_end(ap);
  return 0;
}
int ref_nms_main(void);
/* Runs the reference's run_nms main() on image1 of the loaded pair.  events [4096][4],
 * keypoints [1920][2]; returns the number of surviving keypoints. */
int ref_run_nms(int* events, int* n_events, int* keypoints) {
  g_nms_n_events = 0; g_nms_n_kp = 0;
  ref_nms_main();
  if (events) memcpy(events, g_nms_events, sizeof(g_nms_events));
  if (n_events) *n_events = g_nms_n_events;
  if (keypoints) memcpy(keypoints, g_nms_kp, sizeof(g_nms_kp));
  return g_nms_n_kp;
}

/* The matmul shim and the local feature pool are header-only in the reference:
 * instantiate them here so tests can call the originals. */
#include "gemmini_functions_cpu.h"
#include "local_feature_pool.h"

/* local_bundle_adjustment.c runs as shipped (its main() renamed).  It hands its result, the reduced
 * camera matrix C, to cholesky() -- a stub in the reference (:88-90) that only prints; that symbol is
 * weakened in the object (oracle/Makefile) so this definition receives the matrix instead. */
static float g_lba_C[64 * 64];
static int g_lba_dim;
void cholesky(float* matrix, int dim, int stride) {
  (void)stride;
  g_lba_dim = dim;
  if (dim <= 64) memcpy(g_lba_C, matrix, sizeof(float) * (size_t)dim * (size_t)dim);
}
/* The program's input generator (:92-98, i*10+j) is weakened the same way: with `J` given, every
 * chunk's 64 x 10 column-major factor matrix is filled from it instead, so the reference's loop
 * can be observed on inputs that do not end in NaN. */
static const float* g_lba_J;
void initialize_random_matrix(float* matrix, int rows, int cols) {
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) matrix[j * rows + i] = g_lba_J ? g_lba_J[j * rows + i] : (float)(i * 10 + j);
}
int ref_lba_main(void);
void invert_3x3(float* matrix, int stride);   /* local_bundle_adjustment.c:48-75 */
/* The program reads its H_factor (:148) before writing it: 0 * H_factor is the D term of the first
 * matmul2 (:166-172), so a NaN or Inf left on the stack by whatever ran before would end up in
 * every sum.  The region its frame will occupy is zeroed first, which makes the run repeatable
 * and is the reading the restatement states (H starts at 0). */
static void __attribute__((noinline)) lba_scrub_stack(void) {
  volatile char pad[1 << 18];
  memset((void*)pad, 0, sizeof(pad));
  __asm__ volatile("" : : "r"(pad) : "memory");
}
int ref_lba_run(const float* J, float* C) {
  g_lba_dim = 0;
  g_lba_J = J;
  lba_scrub_stack();
  ref_lba_main();
  g_lba_J = 0;
  if (C) memcpy(C, g_lba_C, sizeof(float) * (size_t)g_lba_dim * (size_t)g_lba_dim);
  return g_lba_dim;
}
void ref_invert_3x3(float* matrix, int stride) { invert_3x3(matrix, stride); }
