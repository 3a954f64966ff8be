// ransac.cu -- the reference's essential-matrix RANSAC as it actually executes
// (src/pnp_solver.c:110-165): the model is the identity for every sample (:63,:81-85),
// so an iteration is an ordered inlier scan of ||E p1 - p2||^2 < threshold in pixel
// units (:89-105,:144-149), and every iteration returns the same set.  The pose is
// pose-from-E via the approximate 3x3 SVD (:168-194).  Bit-exact kernels.
//
//   K2  ransac_identity_kernel : one warp per frame pair, ordered inlier compaction
#include "mv_common.cuh"
#include "svd3.cuh"

namespace {

constexpr int kPairsPerCta = 4;  // warps

__global__ void __launch_bounds__(kPairsPerCta * 32)
ransac_identity_kernel(int n_pairs, int stride, const float* __restrict__ pts,
                       const int32_t* __restrict__ count, int iterations, float thr, int cap,
                       int32_t* __restrict__ num_inliers, int32_t* __restrict__ inliers,
                       float* __restrict__ pose) {
  const int pair = blockIdx.x * kPairsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const float E[3][3] = {{1.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 1.0f}};
  const int n = count[pair];
  const float4* p = reinterpret_cast<const float4*>(pts) + (size_t)pair * stride;
  int found = 0;
  if (iterations > 0) {
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      bool in = false;
      if (i < n) {
        const float4 v = __ldg(p + i);
        in = mvsvd::reproj_error(v.x, v.y, v.z, v.w, E) < thr;  // pnp_solver.c:145-146
      }
      const unsigned votes = __ballot_sync(0xffffffffu, in);
      const int pos = found + __popc(votes & ((1u << lane) - 1));
      if (in && pos < cap && inliers) inliers[(size_t)pair * stride + pos] = i;
      found += __popc(votes);
    }
    if (found > cap) found = cap;
  }
  if (lane == 0) {
    num_inliers[pair] = found;  // 0 where the reference leaves it unwritten
    if (pose) {
      float R1[3][3], R2[3][3], t[3];
      mvsvd::recover_pose(E, R1, R2, t);
      float* o = pose + (size_t)pair * 12;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o[r * 3 + c] = R1[r][c];
      o[9] = t[0]; o[10] = t[1]; o[11] = t[2];
    }
  }
}

}  // namespace

extern "C" mv_status mv_ransac_identity_batch(mv_ctx* ctx, int n_pairs, int stride_pts, const float* d_pts,
                                              const int32_t* d_count, int num_iterations,
                                              float inlier_threshold, int32_t* d_num_inliers,
                                              int32_t* d_inliers, float* d_pose) {
  MV_ENTER(ctx);
  if (n_pairs <= 0 || stride_pts <= 0 || !d_pts || !d_count || !d_num_inliers)
    MV_BAD_ARG(ctx, "mv_ransac_identity_batch");
  mv_prof_scope ps(ctx, "ransac");
  const int grid = (n_pairs + kPairsPerCta - 1) / kPairsPerCta;
  ransac_identity_kernel<<<grid, kPairsPerCta * 32, 0, ctx->stream>>>(
      n_pairs, stride_pts, d_pts, d_count, num_iterations, inlier_threshold,
      stride_pts < 1000 ? stride_pts : 1000 /* MAX_NUM_INLIERS, pnp_solver.c:107 */, d_num_inliers,
      d_inliers, d_pose);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}

// Host-side uses of the same code (legacy single-call symbols in api.cu)
void mv_host_recover_pose(const float E[3][3], float R1[3][3], float R2[3][3], float t[3]) {
  mvsvd::recover_pose(E, R1, R2, t);
}
