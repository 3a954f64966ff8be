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

// The per-cell part, on a staged tile (see softmax_cells_kernel for the method).  Every thread of
// the CTA must call it (warp votes inside); `live` threads own a cell.
__device__ __forceinline__ void softmax_tile_cells(const uint8_t* tile, int n_here, long long cell0,
                                                   const float* __restrict__ scale, int cells_per_frame,
                                                   int32_t* __restrict__ max_idx, float* __restrict__ prob,
                                                   int32_t* __restrict__ num_valid) {
  const int t = threadIdx.x;
  const bool live = t < n_here;
  const long long gc = cell0 + t;
  // 32-bit division whenever the cell index fits (a 64-bit one costs ~40 instructions per cell)
  const int frame = !live ? 0
                    : (gc < 0x7fffffffLL ? (int)((unsigned)gc / (unsigned)cells_per_frame) : (int)(gc / cells_per_frame));
  float c[kTaylorTerms];
  taylor_coeffs(__ldg(scale + frame), c);

  // ---- 1. mask of the non-negative channels (bit ch of m2:m1:m0)
  const int base = t * 65;
  const int first_word = base >> 2;
  const int skew = base & 3;                    // the cell starts `skew` bytes into its first word
  const uint32_t* words = reinterpret_cast<const uint32_t*>(tile);
  uint32_t m0 = 0, m1 = 0, m2 = 0;
  if (live) {
#pragma unroll
    for (int w = 0; w < 17; w++) {
      const uint32_t word = words[first_word + w];
      // sign bits of the four bytes -> bits 0..3 (byte b -> bit b)
      const uint32_t nib = ((((~word) >> 7) & 0x01010101u) * 0x01020408u) >> 24;
      if (w < 8) m0 |= nib << (4 * w);
      else if (w < 16) m1 |= nib << (4 * (w - 8));
      else m2 |= nib;
    }
    // drop the `skew` bytes in front of the cell, keep channels 0..64
    m0 = __funnelshift_r(m0, m1, skew);
    m1 = __funnelshift_r(m1, m2, skew);
    m2 = (m2 >> skew) & 1u;
  }

  // ---- 2. top_N.c:22-49, candidates in ascending channel order
  int arg = 64;
  float top = 0.0f;
  float denom = FLT_MIN;
  while (__any_sync(0xffffffffu, (m0 | m1 | m2) != 0)) {
    int ch = -1;
    if (m0) { ch = __ffs(m0) - 1; m0 &= m0 - 1; }
    else if (m1) { ch = 32 + __ffs(m1) - 1; m1 &= m1 - 1; }
    else if (m2) { ch = 64; m2 = 0; }
    if (ch >= 0) {
      const int x = tile[base + ch];            // non-negative int8
      const float e = taylor_exp(c, x);
      if (ch != 64 && e > top) {
        top = e;
        arg = ch;
      }
      denom = __fadd_rn(denom, e);
    }
  }
  if (!live) return;
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

// One CTA stages 256 consecutive cells (of the flat [n_frames*cells][65] array) into
// shared memory with 16-byte loads; one thread then owns one cell.
//
// ~97 % of the logits are negative and skipped by the reference (top_N.c:29), but which ones
// differs from cell to cell, so walking the 65 bytes and branching per byte makes a warp
// execute the Taylor exponential once per (lane, byte) pair in turn (measured: 837 warp
// instructions per 32 cells, issue-bound at 33 % of HBM).  Instead (55 % of HBM; a bulk-async-copy
// pipelined variant of the same kernel was measured and was not faster -- the rest is the
// strided shared-memory walk, not the loads):
//   1. branch-free: the sign bits of the cell's 17 words are compressed into a 65-bit mask of
//      its non-negative channels (one multiply gathers a word's four sign bits);
//   2. the warp then loops as often as its busiest lane has candidates (typically 3-6), each
//      lane taking its next channel in ascending order -- the order the reference adds the
//      denominator in (top_N.c:26-44) -- so every exponential is evaluated 32 lanes wide.
template <int kTileCells>  // kTileCells * 65 must be a multiple of 16
__global__ void __launch_bounds__(kTileCells)
softmax_cells_kernel(const int8_t* __restrict__ semi, const float* __restrict__ scale,
                     long long total_cells, int cells_per_frame, int vec_ok,
                     int32_t* __restrict__ max_idx, float* __restrict__ prob,
                     int32_t* __restrict__ num_valid) {
  constexpr int kTileBytes = kTileCells * 65;
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

  softmax_tile_cells(tile, n_here, cell0, scale, cells_per_frame, max_idx, prob, num_valid);
}

// ---- K0b -----------------------------------------------------------------
// One CTA per frame: pass 1 counts the valid cells and their prob range (block reduce), pass 2 is
// a ballot/scan-ordered compaction of the first top_n cells whose prob clears the interpolated
// cut, 256 cells per round, stopping once top_n are taken.  A frame is two coalesced sweeps of
// its 8 B/cell detector output.
constexpr int kTopThreads = 256;
__global__ void __launch_bounds__(kTopThreads)
top_n_kernel(const int32_t* __restrict__ max_idx, const float* __restrict__ prob, int cells,
             int top_n, int max_valid, float valid_gt /* round_down(0.01) */,
             int32_t* __restrict__ q_patch, int32_t* __restrict__ q_idx, float* __restrict__ q_prob,
             int32_t* __restrict__ q_count, int32_t* __restrict__ overflow) {
  __shared__ int s_nv[kTopThreads / 32];
  __shared__ float s_hi[kTopThreads / 32], s_lo[kTopThreads / 32];
  __shared__ int s_cnt[kTopThreads / 32];
  const int f = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t* mi = max_idx + (size_t)f * cells;
  const float* pr = prob + (size_t)f * cells;

  int nv = 0;
  float hi = 0.0f, lo = FLT_MAX;  // top_N.c:69
  for (int p = threadIdx.x; p < cells; p += kTopThreads) {
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
  if (lane == 0) { s_nv[wid] = nv; s_hi[wid] = hi; s_lo[wid] = lo; }
  __syncthreads();
  nv = 0; hi = 0.0f; lo = FLT_MAX;
#pragma unroll
  for (int w = 0; w < kTopThreads / 32; w++) {
    nv += s_nv[w];
    hi = fmaxf(hi, s_hi[w]);
    lo = fminf(lo, s_lo[w]);
  }
  if (nv >= max_valid) {  // top_N.c:91-94: the reference exits here
    if (threadIdx.x == 0) {
      q_count[f] = 0;
      if (overflow) overflow[f] = 1;
    }
    return;
  }
  float cut = -FLT_MAX;  // nv <= N: take every valid cell (top_N.c:98-106)
  if (nv > top_n) {      // top_N.c:108-109
    const float split = __fdiv_rn((float)top_n, (float)nv);
    cut = __fadd_rn(__fmul_rn(hi, split), __fmul_rn(lo, __fsub_rn(1.0f, split)));
  }

  int32_t* op = q_patch + (size_t)f * top_n;
  int32_t* oi = q_idx + (size_t)f * top_n;
  float* opr = q_prob + (size_t)f * top_n;
  int taken = 0;
  for (int p0 = 0; p0 < cells && taken < top_n; p0 += kTopThreads) {  // top_N.c:116-133
    const int p = p0 + threadIdx.x;
    int ch = 64;
    float v = 0.0f;
    bool take = false;
    if (p < cells) {
      ch = mi[p];
      v = pr[p];
      take = ch != 64 && v > valid_gt && v >= cut;  // top_N.c:121
    }
    const unsigned votes = __ballot_sync(0xffffffffu, take);
    if (lane == 0) s_cnt[wid] = __popc(votes);
    __syncthreads();
    int before = taken, round = 0;
#pragma unroll
    for (int w = 0; w < kTopThreads / 32; w++) {
      const int c = s_cnt[w];
      before += w < wid ? c : 0;
      round += c;
    }
    const int pos = before + __popc(votes & ((1u << lane) - 1));
    if (take && pos < top_n) {
      op[pos] = p;
      oi[pos] = ch;
      opr[pos] = v;
    }
    taken += round;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    q_count[f] = taken < top_n ? taken : top_n;
    if (overflow) overflow[f] = 0;
  }
}

}  // namespace

extern "C" mv_status mv_softmax_batch(mv_ctx* ctx, int n_frames, int cells, const int8_t* d_semi,
                                      const float* d_semi_scale, int32_t* d_max_idx, float* d_prob,
                                      int32_t* d_num_valid) {
  MV_ENTER(ctx);
  if (n_frames <= 0 || cells <= 0 || !d_semi || !d_semi_scale || !d_max_idx || !d_prob)
    MV_BAD_ARG(ctx, "mv_softmax_batch");
  const long long total = (long long)n_frames * cells;
  const int vec_ok = (reinterpret_cast<uintptr_t>(d_semi) & 15) == 0;
  if (d_num_valid) MV_CUDA(ctx, cudaMemsetAsync(d_num_valid, 0, sizeof(int32_t) * n_frames, ctx->stream));
  mv_prof_scope ps(ctx, "detect");
  softmax_cells_kernel<256><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
      d_semi, d_semi_scale, total, cells, vec_ok, d_max_idx, d_prob, d_num_valid);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}

extern "C" mv_status mv_top_n_batch(mv_ctx* ctx, int n_frames, int cells, int top_n, int max_valid,
                                    const int32_t* d_max_idx, const float* d_prob, int32_t* d_q_patch,
                                    int32_t* d_q_idx, float* d_q_prob, int32_t* d_q_count,
                                    int32_t* d_overflow) {
  MV_ENTER(ctx);
  if (n_frames <= 0 || cells <= 0 || top_n <= 0 || max_valid <= 0 || !d_max_idx || !d_prob ||
      !d_q_patch || !d_q_idx || !d_q_prob || !d_q_count)
    MV_BAD_ARG(ctx, "mv_top_n_batch");
  mv_prof_scope ps(ctx, "topn");
  top_n_kernel<<<n_frames, kTopThreads, 0, ctx->stream>>>(d_max_idx, d_prob, cells, top_n, max_valid,
                                                           mv_round_down(0.01), d_q_patch, d_q_idx,
                                                           d_q_prob, d_q_count, d_overflow);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}
