/* maveric_slam_compat.h -- drop-in declarations for the maveric-slam tracking hot path.
 *
 * Every symbol below keeps the name, argument order and field order that the
 * reference C sources use, so a reference driver (e.g. src/tracking_main.c) links
 * against libmaveric_b200.so without source changes.  In the reference these
 * functions are *defined* in headers / per-executable .c files; here the headers
 * only declare and the definitions live in the shared library, where the compute
 * runs as sm_100a CUDA kernels (there is no CPU fallback: every entry point aborts
 * with a message on stderr if no B200-class device is usable).
 *
 * Reference interface replaced (file:line, relative to the reference tree):
 *   geometry PODs ............ include/types.h:4-23
 *   Frame / frame_create ..... include/frame.h:7-47
 *   compute_top_N/_softmax ... include/top_N.h:8-13     (bodies src/top_N.c:53-165)
 *   essential-matrix RANSAC .. include/pnp_solver.h:3-22 (bodies src/pnp_solver.c:28-194)
 *   projection factor ........ include/projection_factor.h:6-31
 *   matmul / matmul2 ......... include/gemmini_functions_cpu.h:14-19,60-66
 *   track() .................. include/tracking.h:3 (pseudocode there; SE3 stands in
 *                              for the undefined `Transform`)
 *   local feature pool ....... include/local_feature_pool.h:16-336
 */
#ifndef MAVERIC_SLAM_COMPAT_H
#define MAVERIC_SLAM_COMPAT_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- geometry PODs (types.h:4-23); quaternion order is (w, x, y, z) ---- */
#ifndef TYPES_H
#define TYPES_H
typedef struct Vector2f_t { float x, y; } Vector2f;
typedef struct Vector3f_t { float x, y, z; } Vector3f;
typedef struct Quaternionf_t { float w, x, y, z; } Quaternionf;
typedef struct SE3_t { Quaternionf q; Vector3f t; } SE3;
typedef struct Camera_t { float fx, fy, cx, cy; } Camera;

Vector2f    add_Vector2f(Vector2f a, Vector2f b, float scale);        /* types.c:3-8   */
Vector3f    add_Vector3f(Vector3f a, Vector3f b, float scale);        /* types.c:10-16 */
Quaternionf mult_Quaternionf(Quaternionf a, Quaternionf b);           /* types.c:18-25 */
Quaternionf create_Quaternionf(float w, float x, float y, float z);   /* types.c:27-34 */
Quaternionf Quaternionf_from_Vector3f(Vector3f v);                    /* types.c:36-43 */
Quaternionf conjugate_Quaternionf(Quaternionf q);                     /* types.c:45-52 */
Vector3f    Vector3f_from_Quaternionf(Quaternionf q);                 /* types.c:54-60 */
Vector3f    apply_rotation(Quaternionf q, Vector3f v);                /* types.c:62-68 */
Vector3f    apply_transform(SE3 T, Vector3f v);                       /* types.c:70-73 */
#endif

/* ---- frame container (frame.h:7-47) ---- */
#ifndef CELL_SIZE
#define CELL_SIZE 8
#endif
#ifndef DESCRIPTOR_SIZE
#define DESCRIPTOR_SIZE 256
#endif

typedef struct {
  int rows;
  int cols;
  int channels;
  const char* data;

  int num_features;
  int feature_rows;   /* image rows / 8 */
  int feature_cols;   /* image cols / 8 */
  const int* feature_xs;
  const int* feature_ys;

  float semi_scale;
  const int8_t* semi;  /* [feature_rows*feature_cols][65], cell = col*feature_rows + row */
  float desc_scale;
  const int8_t* desc;  /* [feature_rows*feature_cols][256], same cell order */
} Frame;

/* Copies the nine given fields; leaves num_features/feature_xs/feature_ys alone,
 * like frame.h:32-47. */
void frame_create(const int rows, const int cols, const int channels, const char* data,
                  const int feature_rows, const int feature_cols,
                  const float semi_scale, const int8_t* semi,
                  const float desc_scale, const int8_t* desc,
                  Frame* frame);

/* ---- detector post-processing (top_N.h:8-13) ----
 * Legacy shape: 1920 cells (24x80), at most 1000 valid cells (src/top_N.c:51,73,151).
 * compute_top_N prints "Exceed max number of features!" and exit(1)s at the 1000th
 * valid cell, as top_N.c:91-94 does.  compute_softmax adds to *num_valid (caller
 * zeroes it, tracking_main.c:85). */
void compute_top_N(float scale, int8_t semi[2400][65], int N,
                   int* num_selected, int* N_patches, int* N_indices, float* N_probs);
void compute_softmax(float scale, int8_t semi[2400][65],
                     int* num_valid, int* max_indices, float* probs);

/* ---- essential-matrix RANSAC "pnp_solver" (pnp_solver.h:3-22) ---- */
void normalize_points(const int num_points, const float points[][2],
                      const float K[3][3], float normalized_points[][2]);
void compute_essential_matrix(const int num_points,
                              const float pts1_norm[][2], const float pts2_norm[][2],
                              float E[3][3]);
float compute_reprojection_error(const float point1[2], const float point2[2],
                                 const float E[3][3]);
void ransac_essential_matrix(const int num_points,
                             const float points1[][2], const float points2[][2],
                             const float K[3][3],
                             const int num_iterations, const float inlier_threshold,
                             float best_E[3][3], int* best_inliers, int* num_inliers);
void recover_pose_from_essential_matrix(float E[3][3], float R1[3][3],
                                        float R2[3][3], float t[3]);

/* ---- projection factor (projection_factor.h:6-31) ---- */
#ifndef PROJECTION_FACTOR_H
#define PROJECTION_FACTOR_H
typedef struct ProjectionFactor_t {
  Vector3f* landmark;
  SE3* pose;
  Vector2f measurement;
  Vector2f error;
  Camera camera;
} ProjectionFactor;

ProjectionFactor* create_ProjectionFactor(Vector3f* landmark, SE3* pose,
                                          Vector2f measurement, Camera camera);
Vector2f project2d(const Vector3f trans_xyz);
Vector2f cam_project(const Vector3f trans_xyz, const Camera camera);
void compute_error_ProjectionFactor(ProjectionFactor* factor);
#endif

/* ---- matmul shim (gemmini_functions_cpu.h:8-19,60-66), row-major + strides ---- */
#ifndef GEMMINI_TYPE
#define GEMMINI_TYPE float
#endif
#ifndef elem_t
#define elem_t GEMMINI_TYPE
#endif
#ifndef scale_t
#define scale_t GEMMINI_TYPE
#endif

/* C += sA*op(A) * sB*op(B) */
void matmul(size_t dim_I, size_t dim_J, size_t dim_K,
            const elem_t* A, const elem_t* B, elem_t* C,
            size_t stride_A, size_t stride_B, size_t stride_C,
            scale_t A_scale_factor, scale_t B_scale_factor,
            bool transpose_A, bool transpose_B);
/* C = sA*op(A) * sB*op(B) + sD*D  (D == NULL: accumulate into C) */
void matmul2(size_t dim_I, size_t dim_J, size_t dim_K,
             const elem_t* A, const elem_t* B,
             const elem_t* D, elem_t* C,
             size_t stride_A, size_t stride_B, size_t stride_D, size_t stride_C,
             scale_t A_scale_factor, scale_t B_scale_factor, scale_t D_scale_factor,
             bool transpose_A, bool transpose_B);

/* ---- track() (tracking.h:3): semantics follow src/tracking_main.c:84-218 ----
 * last_frame == NULL -> identity.  window_size = 2*radius+1.  `threshold` is the
 * cosine threshold (0.9 in the reference; compared squared, in double).  The pose
 * written is the reference's pose-from-essential output: rotation R1 as a
 * quaternion, t as translation. */
void track(const Frame* last_frame, const Frame* current_frame,
           const int x_shift, const int y_shift, const int window_size,
           const float threshold, SE3* transform);

/* ---- local feature pool (local_feature_pool.h:11-336), host-side ---- */
#define MAX_LOCAL_FRAMES 8
#define MAX_LOCAL_FEATURES 1000
#define LOCAL_FEATURE_POOL_MAX_LOAD_FACTOR 0.75
#define LOCAL_FEATURE_POOL_CAPACITY 3000

typedef struct {
  int word_id;      /* -1: empty */
  int frame_ptr;
  int num_frames;
  int frames[MAX_LOCAL_FRAMES];
  Vector3f coords_3D;
} LocalFeature;

typedef struct {
  int key;
  LocalFeature value;
  bool is_occupied;
} HashEntry;

typedef struct {
  HashEntry entries[LOCAL_FEATURE_POOL_CAPACITY];
  int size;
  int capacity;
} LocalFeaturePool;

typedef struct {
  LocalFeature* feature;
  bool inserted;
} LocalFeaturePoolInsertResult;

void init_local_feature(LocalFeature* feature);
void init_local_feature_with_id(LocalFeature* feature, int word_id, int frame_num);
void update_local_feature(LocalFeature* feature, int frame_num);
bool remove_old_frame(LocalFeature* feature, int oldest_keep_frame);
void init_hash_entry(HashEntry* entry);
void delete_hash_entry(HashEntry* entry);
int  hash(int key, int capacity);
void init_local_feature_pool(LocalFeaturePool* pool);
LocalFeaturePoolInsertResult local_feature_pool_insert(LocalFeaturePool* pool,
                                                       int key, LocalFeature value);
int  chain_replacement(LocalFeaturePool* pool, int remove_index);
bool local_feature_pool_delete(LocalFeaturePool* pool, int key);
void local_feature_pool_rehash(LocalFeaturePool* pool);
float local_feature_pool_load_factor(LocalFeaturePool* pool);
void local_feature_pool_remove_old(LocalFeaturePool* pool, int current_frame_num);
void local_feature_pool_valid_keys(LocalFeaturePool* pool, int* num_keys, int* keys);
void local_feature_pool_check_invariant(LocalFeaturePool* pool, int cur_frame, bool print);

#ifdef __cplusplus
}
#endif
#endif /* MAVERIC_SLAM_COMPAT_H */
