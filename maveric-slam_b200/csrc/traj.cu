// traj.cu -- pose chaining (reference: python/compute_trajectory.py:49-51,76-77), the step
// right after the multi-GPU pose gather.
//
// The reference walks the relative transforms sequentially:
//     current_pose[:3,:3] = transform[:3,:3] @ current_pose[:3,:3]
//     current_pose[:3, 3] = transform[:3, 3] + current_pose[:3, 3]
// i.e. pose_k = (R_k ... R_1, t_1 + ... + t_k): an inclusive scan under the associative
// operator (R_b, t_b) o (R_a, t_a) = (R_b R_a, t_b + t_a), a later b applied on the left.
// One CTA scans 1024 transforms per pass (warp shuffles, then the 32 warp totals), carrying
// the running pose between passes; a KITTI-00-length sequence is five passes.  float64
// throughout, like numpy; tiny next to the path it follows (tens of microseconds).
#include "mv_common.cuh"

namespace {

struct Pose {
  double R[9];
  double t[3];
};

__device__ __forceinline__ Pose pose_identity() {
  Pose p;
#pragma unroll
  for (int i = 0; i < 9; i++) p.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  p.t[0] = p.t[1] = p.t[2] = 0.0;
  return p;
}

// later o earlier
__device__ __forceinline__ Pose compose(const Pose& later, const Pose& earlier) {
  Pose o;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      o.R[i * 3 + j] = later.R[i * 3] * earlier.R[j] + later.R[i * 3 + 1] * earlier.R[3 + j] +
                       later.R[i * 3 + 2] * earlier.R[6 + j];
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = later.t[i] + earlier.t[i];
  return o;
}

__device__ __forceinline__ Pose shfl_up(const Pose& p, int delta) {
  Pose o;
#pragma unroll
  for (int i = 0; i < 9; i++) o.R[i] = __shfl_up_sync(0xffffffffu, p.R[i], delta);
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = __shfl_up_sync(0xffffffffu, p.t[i], delta);
  return o;
}

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
chain_transforms_kernel(int n, const double* __restrict__ T, double* __restrict__ traj) {
  __shared__ Pose s_warp[kScanThreads / 32];
  __shared__ Pose s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_carry = pose_identity();
    const Pose id = pose_identity();
    for (int i = 0; i < 9; i++) traj[(i / 3) * 4 + (i % 3)] = id.R[i];
    for (int i = 0; i < 3; i++) traj[i * 4 + 3] = 0.0;
  }
  __syncthreads();
  for (int base = 0; base < n; base += kScanThreads) {
    const int k = base + threadIdx.x;
    Pose p = pose_identity();
    if (k < n) {
      const double* m = T + (size_t)k * 12;
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) p.R[i * 3 + j] = m[i * 4 + j];
        p.t[i] = m[i * 4 + 3];
      }
    }
    // inclusive scan inside the warp: lane l ends with p_l o ... o p_0
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const Pose e = shfl_up(p, d);
      if (lane >= d) p = compose(p, e);
    }
    if (lane == 31) s_warp[wid] = p;
    __syncthreads();
    if (wid == 0) {   // scan of the warp totals
      Pose w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const Pose e = shfl_up(w, d);
        if (lane >= d) w = compose(w, e);
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    Pose prefix = s_carry;                                  // everything before this pass
    if (wid > 0) prefix = compose(s_warp[wid - 1], prefix); // ... and before this warp
    p = compose(p, prefix);
    if (k < n) {
      double* o = traj + (size_t)(k + 1) * 12;
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) o[i * 4 + j] = p.R[i * 3 + j];
        o[i * 4 + 3] = p.t[i];
      }
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) s_carry = p;       // padded with identities past n
    __syncthreads();
  }
}

// unit quaternion (w,x,y,z) + t  ->  row-major 3x4 [R|t]; the rotation formula of
// src/types.c:62-68 (q v q*) expanded for a unit quaternion, evaluated in float64
__global__ void results_to_transforms_kernel(int n, const mv_pair_result* __restrict__ res,
                                             double* __restrict__ T) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double w = res[k].q[0], x = res[k].q[1], y = res[k].q[2], z = res[k].q[3];
  double* o = T + (size_t)k * 12;
  o[0] = 1.0 - 2.0 * (y * y + z * z); o[1] = 2.0 * (x * y - w * z);       o[2] = 2.0 * (x * z + w * y);
  o[4] = 2.0 * (x * y + w * z);       o[5] = 1.0 - 2.0 * (x * x + z * z); o[6] = 2.0 * (y * z - w * x);
  o[8] = 2.0 * (x * z - w * y);       o[9] = 2.0 * (y * z + w * x);       o[10] = 1.0 - 2.0 * (x * x + y * y);
  o[3] = res[k].t[0]; o[7] = res[k].t[1]; o[11] = res[k].t[2];
}

}  // namespace

extern "C" mv_status mv_chain_transforms(mv_ctx* ctx, int n, const double* d_transforms, double* d_traj) {
  MV_ENTER(ctx);
  if (n < 0 || !d_traj || (n > 0 && !d_transforms)) MV_BAD_ARG(ctx, "mv_chain_transforms");
  mv_prof_scope ps(ctx, "chain");
  chain_transforms_kernel<<<1, kScanThreads, 0, ctx->stream>>>(n, d_transforms, d_traj);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}

extern "C" mv_status mv_results_to_transforms(mv_ctx* ctx, int n, const mv_pair_result* d_results,
                                              double* d_transforms) {
  MV_ENTER(ctx);
  if (n <= 0 || !d_results || !d_transforms) MV_BAD_ARG(ctx, "mv_results_to_transforms");
  results_to_transforms_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, d_results, d_transforms);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}
