// nms.cu -- 4-pixel non-maximum suppression across cell borders (reference: src/run_nms.c:
// 65-156), the stage in front of the matcher.
//
// The reference visits the corners of the cell grid sequentially, x outer / y inner, both
// inclusive of the far edge.  A corner gathers the keypoints of its up-to-four adjacent cells
// that lie within 6 px of it (:75-106) and lets the strongest one suppress the others closer
// than 4 px on both axes, repeatedly (:110-151).  Suppression rewrites the cell
// (max_idx = 64, prob = 64.0f, :138-139), and every later corner sees it, so the result depends
// on the visiting order.
//
// Order-preserving parallel form: corners (x, y) and (x', y') touch a common cell only when
// |x-x'| <= 1 and |y-y'| <= 1, and for all such pairs the reference's earlier corner has the
// smaller t = 2x + y.  Corners with equal t never share a cell.  So one warp per frame sweeps
// t = 0 .. 2*cols + rows, its lanes taking the corners of that wavefront, a __syncwarp between
// wavefronts; frames are independent.  HBM traffic is the 8 B/cell detector output, in place.
#include "mv_common.cuh"

namespace {

constexpr int kNmsWarpsPerCta = 4;

__device__ __forceinline__ void nms_corner(int xi, int yi, int rows, int cols, volatile int32_t* mi,
                                           volatile float* pr) {
  int nv = 0, patches[4], xs[4], ys[4];
  float probs[4];
#pragma unroll
  for (int xd = -1; xd <= 0; xd++) {
    const int xg = xi + xd;
#pragma unroll
    for (int yd = -1; yd <= 0; yd++) {
      const int yg = yi + yd;
      if (xg < 0 || xg >= cols || yg < 0 || yg >= rows) continue;   // :77,:81
      const int patch = xg * rows + yg;
      const int index = mi[patch];
      if (index == 64) continue;                                    // :87
      const int px = index % 8, py = index / 8;
      if (xd == -1 && px < 2) continue;                             // :94-97
      if (xd == 0 && px >= 6) continue;
      if (yd == -1 && py < 2) continue;
      if (yd == 0 && py >= 6) continue;
      patches[nv] = patch; probs[nv] = pr[patch];
      xs[nv] = xg * 8 + px; ys[nv] = yg * 8 + py;
      nv++;
    }
  }
  for (;;) {
    float max_prob = 0.0f;
    int max_index = -1;
    for (int i = 0; i < nv; i++)                                    // :115-120 (patch 0 is skipped here)
      if (patches[i] > 0 && probs[i] > max_prob) { max_prob = probs[i]; max_index = i; }
    if (max_index == -1) break;
    for (int i = 0; i < nv; i++)                                    // :125-130
      if (patches[i] >= 0 && probs[i] > max_prob) { max_prob = probs[i]; max_index = i; }
    for (int i = 0; i < nv; i++) {                                  // :132-147
      if (i == max_index || patches[i] < 0) continue;
      if (abs(xs[max_index] - xs[i]) < 4 && abs(ys[max_index] - ys[i]) < 4) {
        mi[patches[i]] = 64;
        pr[patches[i]] = 64.0f;
        patches[i] = -1; probs[i] = -1.0f;
      }
    }
    probs[max_index] = -1.0f; patches[max_index] = -1;             // :149-150
  }
}

__global__ void __launch_bounds__(kNmsWarpsPerCta * 32)
nms_wavefront_kernel(int n_frames, int rows, int cols, int32_t* __restrict__ max_idx, float* __restrict__ prob) {
  const int frame = blockIdx.x * kNmsWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (frame >= n_frames) return;
  const size_t cells = (size_t)rows * cols;
  volatile int32_t* mi = max_idx + (size_t)frame * cells;
  volatile float* pr = prob + (size_t)frame * cells;
  const int t_end = 2 * cols + rows;
  for (int t = 0; t <= t_end; t++) {
    for (int yi = lane; yi <= rows; yi += 32) {
      const int d = t - yi;
      if (d < 0 || (d & 1) || (d >> 1) > cols) continue;
      nms_corner(d >> 1, yi, rows, cols, mi, pr);
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" mv_status mv_nms_batch(mv_ctx* ctx, int n_frames, int rows, int cols, int32_t* d_max_idx, float* d_prob) {
  MV_ENTER(ctx);
  if (n_frames <= 0 || rows <= 0 || cols <= 0 || !d_max_idx || !d_prob) MV_BAD_ARG(ctx, "mv_nms_batch");
  mv_prof_scope ps(ctx, "nms");
  nms_wavefront_kernel<<<(n_frames + kNmsWarpsPerCta - 1) / kNmsWarpsPerCta, kNmsWarpsPerCta * 32, 0, ctx->stream>>>(
      n_frames, rows, cols, d_max_idx, d_prob);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}

extern "C" mv_status run_nms_ex(mv_ctx* ctx, int rows, int cols, int* max_indices, float* probs) {
  MV_ENTER(ctx);
  if (rows <= 0 || cols <= 0 || !max_indices || !probs) MV_BAD_ARG(ctx, "run_nms_ex");
  const size_t cells = (size_t)rows * cols;
  void *di, *dp;
  mv_status st;
  if ((st = mv_scratch(ctx, "nms.idx", sizeof(int32_t) * cells, &di))) return st;
  if ((st = mv_scratch(ctx, "nms.prob", sizeof(float) * cells, &dp))) return st;
  MV_CUDA(ctx, cudaMemcpyAsync(di, max_indices, sizeof(int32_t) * cells, cudaMemcpyHostToDevice, ctx->stream));
  MV_CUDA(ctx, cudaMemcpyAsync(dp, probs, sizeof(float) * cells, cudaMemcpyHostToDevice, ctx->stream));
  if ((st = mv_nms_batch(ctx, 1, rows, cols, (int32_t*)di, (float*)dp))) return st;
  MV_CUDA(ctx, cudaMemcpyAsync(max_indices, di, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost, ctx->stream));
  MV_CUDA(ctx, cudaMemcpyAsync(probs, dp, sizeof(float) * cells, cudaMemcpyDeviceToHost, ctx->stream));
  MV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MV_OK;
}
