// detector.cu -- detector post-processing kernels (reference: src/top_N.c).
//
//   K0a  softmax_cells_kernel : per-cell Taylor softmax + argmax  (top_N.c:12-49,136-165)
//   K0b  top_n_kernel         : thresholded, patch-ordered pick   (top_N.c:53-134)
//
// Both are HBM-streaming kernels: K0a reads 65 B and writes 8 B per cell.  fp32
// results must equal the reference bit for bit, so every float operation is an
// explicit round-to-nearest intrinsic (no FMA contraction) in the reference's order.
#include "mv_common.cuh"

#include <float.h>

namespace {

constexpr int kTileCells = 256;                 // cells per CTA
constexpr int kTileBytes = kTileCells * 65;     // 16640, a multiple of 16
constexpr int kTaylorTerms = 5;                 // top_N.c:7

// top_N.c:59-63
__device__ __forceinline__ void taylor_coeffs(float scale, float c[kTaylorTerms]) {
  c[0] = 1.0f;
#pragma unroll
  for (int i = 1; i < kTaylorTerms; i++) c[i] = __fdiv_rn(__fmul_rn(c[i - 1], scale), (float)i);
}

// top_N.c:12-20 (powers of x stay int32)
__device__ __forceinline__ float taylor_exp(const float c[kTaylorTerms], int x) {
  float acc = 1.0f;
  int xp = x;
#pragma unroll
  for (int i = 1; i < kTaylorTerms; i++) {
    acc = __fadd_rn(acc, __fmul_rn(c[i], (float)xp));
    xp *= x;
  }
  return acc;
}

// One CTA stages 256 consecutive cells (of the flat [n_frames*cells][65] array) into
// shared memory with 16-byte loads, then one thread walks one cell word by word;
// words without a non-negative byte (the common case, ~97 % of logits are negative)
// are skipped with one mask test.
__global__ void __launch_bounds__(kTileCells)
softmax_cells_kernel(const int8_t* __restrict__ semi, const float* __restrict__ scale,
                     long long total_cells, int cells_per_frame, int vec_ok,
                     int32_t* __restrict__ max_idx, float* __restrict__ prob,
                     int32_t* __restrict__ num_valid) {
  __shared__ __align__(16) uint8_t tile[kTileBytes + 16];
  const long long cell0 = (long long)blockIdx.x * kTileCells;
  const long long remaining = total_cells - cell0;
  const int n_here = remaining < kTileCells ? (int)remaining : kTileCells;
  const long long byte0 = cell0 * 65;
  const int n_bytes = n_here * 65;

  if (vec_ok) {
    const int n_vec = n_bytes >> 4;
    const int4* src = reinterpret_cast<const int4*>(semi + byte0);
    int4* dst = reinterpret_cast<int4*>(tile);
    for (int i = threadIdx.x; i < n_vec; i += kTileCells) dst[i] = __ldg(src + i);
    for (int i = (n_vec << 4) + threadIdx.x; i < n_bytes; i += kTileCells) tile[i] = (uint8_t)semi[byte0 + i];
  } else {
    for (int i = threadIdx.x; i < n_bytes; i += kTileCells) tile[i] = (uint8_t)semi[byte0 + i];
  }
  __syncthreads();

  const int t = threadIdx.x;
  if (t >= n_here) return;
  const long long gc = cell0 + t;
  const int frame = (int)(gc / cells_per_frame);
  float c[kTaylorTerms];
  taylor_coeffs(__ldg(scale + frame), c);

  // top_N.c:22-49
  int arg = 64;
  float top = 0.0f;
  float denom = FLT_MIN;
  const int base = t * 65;
  const int first_word = base >> 2;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(tile);
#pragma unroll 1
  for (int w = 0; w < 17; w++) {
    const uint32_t word = words[first_word + w];
    uint32_t nonneg = ~word & 0x80808080u;
    while (nonneg) {
      const int b = (__ffs(nonneg) - 1) >> 3;
      nonneg &= nonneg - 1;
      const int ch = ((first_word + w) << 2) + b - base;
      if (ch < 0 || ch > 64) continue;
      const int x = (int)((word >> (8 * b)) & 0xFF);
      const float e = taylor_exp(c, x);
      if (ch != 64 && e > top) {
        top = e;
        arg = ch;
      }
      denom = __fadd_rn(denom, e);
    }
  }
  const float p = __fdiv_rn(top, denom);
  max_idx[gc] = arg;
  prob[gc] = (arg != 64) ? p : -1.0f;  // top_N.c:156-163

  if (num_valid) {
    const unsigned active = __activemask();
    const int lead_frame = __shfl_sync(active, frame, __ffs(active) - 1);
    const bool uniform = __all_sync(active, frame == lead_frame);
    if (uniform) {
      const unsigned votes = __ballot_sync(active, arg != 64);
      if ((threadIdx.x & 31) == (__ffs(active) - 1) && votes) atomicAdd(num_valid + frame, __popc(votes));
    } else if (arg != 64) {
      atomicAdd(num_valid + frame, 1);
    }
  }
}

// ---- K0b -----------------------------------------------------------------
constexpr int kTopNThreads = 512;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* warp_sums) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < (kTopNThreads / 32) ? warp_sums[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += n;
    }
    if (lane < (kTopNThreads / 32)) warp_sums[lane] = si - s;  // exclusive warp offsets
    if (lane == 31) *total = si;
  }
  __syncthreads();
  const int out = warp_sums[wid] + inc - v;
  __syncthreads();
  return out;
}

// One CTA per frame.  Pass 1: count the valid cells and their prob range.  Pass 2:
// ordered compaction of the first top_n cells whose prob clears the interpolated cut.
__global__ void __launch_bounds__(kTopNThreads)
top_n_kernel(const int32_t* __restrict__ max_idx, const float* __restrict__ prob, int cells,
             int top_n, int max_valid, float valid_gt /* round_down(0.01) */,
             int32_t* __restrict__ q_patch, int32_t* __restrict__ q_idx, float* __restrict__ q_prob,
             int32_t* __restrict__ q_count, int32_t* __restrict__ overflow) {
  __shared__ int s_warp[kTopNThreads / 32];
  __shared__ float s_hi[kTopNThreads / 32], s_lo[kTopNThreads / 32];
  __shared__ int s_total, s_nv, s_base;
  __shared__ float s_cut;

  const int f = blockIdx.x;
  const int32_t* mi = max_idx + (size_t)f * cells;
  const float* pr = prob + (size_t)f * cells;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  int nv = 0;
  float hi = 0.0f, lo = FLT_MAX;  // top_N.c:69
  for (int p = threadIdx.x; p < cells; p += kTopNThreads) {
    const float v = pr[p];
    if (mi[p] != 64 && v > valid_gt) {  // top_N.c:77
      nv++;
      hi = fmaxf(hi, v);
      lo = fminf(lo, v);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  }
  if (lane == 0) { s_warp[wid] = nv; s_hi[wid] = hi; s_lo[wid] = lo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tn = 0;
    float th = 0.0f, tl = FLT_MAX;
    for (int w = 0; w < kTopNThreads / 32; w++) {
      tn += s_warp[w];
      th = fmaxf(th, s_hi[w]);
      tl = fminf(tl, s_lo[w]);
    }
    s_nv = tn;
    float cut = -FLT_MAX;  // nv <= N: take every valid cell (top_N.c:98-106)
    if (tn > top_n) {      // top_N.c:108-109
      const float split = __fdiv_rn((float)top_n, (float)tn);
      cut = __fadd_rn(__fmul_rn(th, split), __fmul_rn(tl, __fsub_rn(1.0f, split)));
    }
    s_cut = cut;
    s_base = 0;
  }
  __syncthreads();
  const int total_valid = s_nv;
  if (total_valid >= max_valid) {  // top_N.c:91-94: the reference exits here
    if (threadIdx.x == 0) {
      q_count[f] = 0;
      if (overflow) overflow[f] = 1;
    }
    return;
  }
  const float cut = s_cut;

  int32_t* op = q_patch + (size_t)f * top_n;
  int32_t* oi = q_idx + (size_t)f * top_n;
  float* opr = q_prob + (size_t)f * top_n;
  for (int p0 = 0; p0 < cells; p0 += kTopNThreads) {
    const int p = p0 + threadIdx.x;
    int take = 0, ch = 64;
    float v = 0.0f;
    if (p < cells) {
      ch = mi[p];
      v = pr[p];
      take = (ch != 64 && v > valid_gt && v >= cut) ? 1 : 0;  // top_N.c:121
    }
    const int base = s_base;
    const int pos = base + block_exclusive_scan(take, &s_total, s_warp);
    if (take && pos < top_n) {
      op[pos] = p;
      oi[pos] = ch;
      opr[pos] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + s_total;
    __syncthreads();
    if (s_base >= top_n) break;  // top_N.c:128-130
  }
  if (threadIdx.x == 0) {
    q_count[f] = s_base < top_n ? s_base : top_n;
    if (overflow) overflow[f] = 0;
  }
}

}  // namespace

extern "C" mv_status mv_softmax_batch(mv_ctx* ctx, int n_frames, int cells, const int8_t* d_semi,
                                      const float* d_semi_scale, int32_t* d_max_idx, float* d_prob,
                                      int32_t* d_num_valid) {
  if (!ctx) return MV_ERR_BAD_ARG;
  if (n_frames <= 0 || cells <= 0 || !d_semi || !d_semi_scale || !d_max_idx || !d_prob)
    MV_BAD_ARG(ctx, "mv_softmax_batch");
  const long long total = (long long)n_frames * cells;
  const int grid = (int)((total + kTileCells - 1) / kTileCells);
  const int vec_ok = (reinterpret_cast<uintptr_t>(d_semi) & 15) == 0;
  if (d_num_valid) MV_CUDA(ctx, cudaMemsetAsync(d_num_valid, 0, sizeof(int32_t) * n_frames, ctx->stream));
  mv_prof_scope ps(ctx, "detect");
  softmax_cells_kernel<<<grid, kTileCells, 0, ctx->stream>>>(d_semi, d_semi_scale, total, cells, vec_ok,
                                                             d_max_idx, d_prob, d_num_valid);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}

extern "C" mv_status mv_top_n_batch(mv_ctx* ctx, int n_frames, int cells, int top_n, int max_valid,
                                    const int32_t* d_max_idx, const float* d_prob, int32_t* d_q_patch,
                                    int32_t* d_q_idx, float* d_q_prob, int32_t* d_q_count,
                                    int32_t* d_overflow) {
  if (!ctx) return MV_ERR_BAD_ARG;
  if (n_frames <= 0 || cells <= 0 || top_n <= 0 || max_valid <= 0 || !d_max_idx || !d_prob ||
      !d_q_patch || !d_q_idx || !d_q_prob || !d_q_count)
    MV_BAD_ARG(ctx, "mv_top_n_batch");
  mv_prof_scope ps(ctx, "topn");
  top_n_kernel<<<n_frames, kTopNThreads, 0, ctx->stream>>>(d_max_idx, d_prob, cells, top_n, max_valid,
                                                           mv_round_down(0.01), d_q_patch, d_q_idx,
                                                           d_q_prob, d_q_count, d_overflow);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}
