// match_tc.cu -- windowed int8 descriptor search on the 5th-generation tensor cores
// (reference: src/tracking_main.c:18-43,103-194; same results as match.cu, bit for bit).
//
// The reference evaluates a 64-dimension dot product for every candidate of a query but the
// first one (squared_dist, tracking_main.c:33-41), so the bulk of the search is an int8 GEMM
//     D[query][cell] = sum_{k<64} desc1[query][k] * desc0[cell][k]
// over the cells of the query's search window.  A work item is a tile of 128 consecutive
// queries of one frame pair; its candidates are the full cell columns spanned by the union of
// the queries' windows, "the window as a TMA box" (SURVEY §7):
//
//   warp 0      TMA producer: one 4-D box [64 B][rows][Cx columns][1 frame] of frame 0's
//               descriptor tensor per chunk, SWIZZLE_64B, into a 6-deep shared-memory ring
//   warp 1      MMA issuer: per chunk two tcgen05.mma.kind::i8 (M=128, N=Cx*rows<=256, K=32)
//               into one of two 256-column TMEM accumulator stages
//   warps 2-5   query warps, thread <-> query row, one tile ahead of the epilogue: the first 64
//               bytes of the query descriptor into the A operand (same swizzle), the validity
//               bits of the tile's cell range, and the query's leading 256-dimension
//               evaluation (squared_dist's first call, tracking_main.c:21-32, sticky while the
//               candidate norm is 0) with dp4a -- its global-load latency hides behind the
//               previous tile's epilogue; the per-row state goes to shared memory
//   warps 6-13  epilogue, thread <-> query row (TMEM lane), two threads per row splitting the
//               32-column blocks: tcgen05.ld, then a BRANCH-FREE pass that squares each
//               accumulator, applies window/validity mask and key filter and collects the
//               few survivors in a bitmask; only survivors take the reference's exact float
//               path (see "filter" below), re-read one TMEM column at a time
//
// The distance matrix never leaves TMEM.
//
// Filter.  For the non-leading candidates of a query the score (tracking_main.c:154) is
//     s = float(int32(dot*dot)) / float(int32(norm_F * norm_q64))
// with a denominator that is constant per query.  int->float and IEEE division are monotone,
// so s is non-decreasing in key = n (denominator >= 0) or key = ~n (denominator < 0), n the
// wrapped int32 square.  A candidate whose key does not exceed the largest key seen so far can
// neither pass the threshold nor beat the current best (strict '>' at :155-156), so the exact
// float path runs only on strict prefix maxima of the key, which start above a conservative
// bound derived from the acceptance threshold.  Ties keep the smaller cell index = the
// earlier candidate in the reference's x-outer / y-inner scan.
#include "mv_common.cuh"
#include "sm100_ptx.cuh"

namespace {

using namespace sm100;

constexpr int kTileQ = 128;
constexpr int kBStages = 6;
constexpr int kBStageBytes = 256 * 64;
constexpr int kAStages = 4;
constexpr int kAStageBytes = kTileQ * 64;
constexpr int kMaxAccStages = 16;   // accumulator stages: 512 TMEM columns / columns per chunk (TcGeom)
constexpr int kTmemCols = 512;
constexpr int kQueryWarps = 4;
constexpr int kQueryThreads = 32 * kQueryWarps;     // == kTileQ: one thread per query row
constexpr int kEpiWarp0 = 2 + kQueryWarps;
constexpr int kEpiWarps = 16;
constexpr int kParts = kEpiWarps / 4;   // threads per query row, each takes every kParts-th 32-column block
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
constexpr int kVPadWords = 12;   // zero words after a frame's validity bits (chunk overrun + funnel)

struct TcGeom {
  int rows, cols, cells, shift_x, shift_y, radius, top_n;
  int cx;             // cell columns per chunk
  int n_chunk;        // UMMA N: cx*rows rounded up to 16
  int acc_stride;     // TMEM columns per accumulator stage: n_chunk rounded up to 32
  int acc_stages;     // kTmemCols / acc_stride, at most kMaxAccStages
  int vwords;         // validity words per frame (incl. padding)
  int tiles_per_pair, n_items;
  float accept_gt;    // (double)s > thr^2  <=>  s > accept_gt
  float prob_lt;      // (double)p < min    <=>  p < prob_lt
};

// bit c of frame f's word array: cell c is a candidate (tracking_main.c:142,146)
__global__ void valid_bits_kernel(int cells, int vwords, float prob_lt, const int32_t* __restrict__ max_idx,
                                  const float* __restrict__ prob, uint32_t* __restrict__ vbits) {
  const int f = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool v = false;
  if (c < cells) v = (max_idx[(size_t)f * cells + c] != 64) && !(prob[(size_t)f * cells + c] < prob_lt);
  const unsigned w = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && (c >> 5) < vwords) vbits[(size_t)f * vwords + (c >> 5)] = w;
}

struct TileSpan {
  int f0, f1, q0, n_rows;   // n_rows == 0: nothing to do
  int X0, n_chunks;         // first candidate column, chunks of cx columns
};

// Warp-collective: the tile's query count and the cell columns its windows span.
__device__ __forceinline__ TileSpan compute_tile_span(const TcGeom& g, int item, const int32_t* __restrict__ f0_of,
                                              const int32_t* __restrict__ f1_of,
                                              const int32_t* __restrict__ q_patch,
                                              const int32_t* __restrict__ q_count) {
  TileSpan t;
  const int pair = item / g.tiles_per_pair;
  t.q0 = (item - pair * g.tiles_per_pair) * kTileQ;
  t.f0 = f0_of ? f0_of[pair] : pair;
  t.f1 = f1_of ? f1_of[pair] : pair + 1;
  const int nq = min(q_count[t.f1], g.top_n);
  t.n_rows = max(0, min(kTileQ, nq - t.q0));
  int xmin = 0x7fffffff, xmax = -1;
  for (int r = threadIdx.x & 31; r < t.n_rows; r += 32) {
    const int x = q_patch[(size_t)t.f1 * g.top_n + t.q0 + r] / g.rows;
    xmin = min(xmin, x);
    xmax = max(xmax, x);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
  }
  t.X0 = max(xmin + g.shift_x - g.radius, 0);
  const int X1 = min(xmax + g.shift_x + g.radius, g.cols - 1);
  t.n_chunks = (t.n_rows > 0 && X1 >= t.X0) ? (X1 - t.X0 + g.cx) / g.cx : 0;
  return t;
}

// The spans of all tiles, computed once by a small kernel (one warp per tile): every role of the
// matcher then gets a tile's span with one 16-byte load instead of a chain of dependent loads
// and shuffles per role and tile.
__global__ void tile_spans_kernel(TcGeom g, const int32_t* __restrict__ f0_of, const int32_t* __restrict__ f1_of,
                                  const int32_t* __restrict__ q_patch, const int32_t* __restrict__ q_count,
                                  int4* __restrict__ spans) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= g.n_items) return;
  const TileSpan t = compute_tile_span(g, item, f0_of, f1_of, q_patch, q_count);
  if ((threadIdx.x & 31) == 0) spans[item] = make_int4(t.n_rows, t.X0, t.n_chunks, 0);
}

__device__ __forceinline__ TileSpan tile_span(const TcGeom& g, int item, const int32_t* __restrict__ f0_of,
                                              const int32_t* __restrict__ f1_of, const int4* __restrict__ spans) {
  TileSpan t;
  const int pair = item / g.tiles_per_pair;
  t.q0 = (item - pair * g.tiles_per_pair) * kTileQ;
  t.f0 = f0_of ? f0_of[pair] : pair;
  t.f1 = f1_of ? f1_of[pair] : pair + 1;
  const int4 s = __ldg(spans + item);
  t.n_rows = s.x; t.X0 = s.y; t.n_chunks = s.z;
  return t;
}

__device__ __forceinline__ int first_set_in_range(const uint32_t* sv, int base_bit, int lo, int hi) {
  int p = lo - base_bit;
  const int e = hi - base_bit;
  while (p <= e) {
    const uint32_t w = sv[p >> 5] >> (p & 31);
    if (w) {
      const int q = p + __ffs(w) - 1;
      return q <= e ? q + base_bit : -1;
    }
    p = (p | 31) + 1;
  }
  return -1;
}

// tracking_main.c:154 with defined (two's complement) wrap
__device__ __forceinline__ float wrapped_cos2(int dot, int n_cand, int n_query) {
  const int num = (int)((unsigned)dot * (unsigned)dot);
  const int den = (int)((unsigned)n_cand * (unsigned)n_query);
  return __fdiv_rn(__int2float_rn(num), __int2float_rn(den));
}

__device__ __forceinline__ int dp4a4(const int4& a, const int4& b, int acc) {
  acc = __dp4a(a.x, b.x, acc);
  acc = __dp4a(a.y, b.y, acc);
  acc = __dp4a(a.z, b.z, acc);
  return __dp4a(a.w, b.w, acc);
}

struct MergeSlot {
  float s;
  int cell;
};

// Per-row state the query warps hand to the epilogue (32 bytes).
struct __align__(16) RowInfo {
  int xwin;       // x_lo | x_hi << 16   (x_hi < x_lo: empty window / inactive row)
  int ywin;       // y_lo | y_hi << 16
  int lead_end;   // last cell scored over 256 dims (-1: none)
  int bcell;      // best so far (-1: none)
  float bs;       // its score
  float den_f;    // float(int32(norm_F * norm_q64)), the denominator of every later score
  int curmax;     // start of the key filter
  int flip;       // 0 / -1: key = n ^ flip
};

__global__ void __launch_bounds__(kThreads, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcGeom g, const int32_t* __restrict__ f0_of,
                const int32_t* __restrict__ f1_of, const int8_t* __restrict__ desc,
                const uint32_t* __restrict__ vbits, const int32_t* __restrict__ q_patch,
                const int32_t* __restrict__ q_count, const int4* __restrict__ spans,
                int32_t* __restrict__ best_cell, float* __restrict__ best_score, size_t part_stride,
                int* abort_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = sB + kBStages * kBStageBytes;
  uint32_t* sV = reinterpret_cast<uint32_t*>(sA + kAStages * kAStageBytes);   // [kAStages][vstride]
  const int vstride = (g.vwords + 3) & ~3;
  RowInfo* sR = reinterpret_cast<RowInfo*>(sV + kAStages * ((g.vwords + 3) & ~3));   // [kAStages][kTileQ]

  __shared__ uint64_t bar_full_b[kBStages], bar_empty_b[kBStages];
  __shared__ uint64_t bar_full_a[kAStages], bar_empty_a[kAStages];
  __shared__ uint64_t bar_acc_full[kMaxAccStages], bar_acc_empty[kMaxAccStages];
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBStages; i++) { mbar_init(smem_u32(&bar_full_b[i]), 1); mbar_init(smem_u32(&bar_empty_b[i]), 1); }
    for (int i = 0; i < kAStages; i++) {
      mbar_init(smem_u32(&bar_full_a[i]), kQueryThreads);
      mbar_init(smem_u32(&bar_empty_a[i]), 1 + kEpiWarps);
    }
    for (int i = 0; i < g.acc_stages; i++) { mbar_init(smem_u32(&bar_acc_full[i]), 1); mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t box_bytes = 64u * (uint32_t)g.rows * (uint32_t)g.cx;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    uint32_t chunk = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      const TileSpan t = tile_span(g, item, f0_of, f1_of, spans);
      if (lane == 0) {
        for (int c = 0; c < t.n_chunks; c++, chunk++) {
          const uint32_t s = chunk % kBStages, ph = (chunk / kBStages) & 1;
          mbar_wait(smem_u32(&bar_empty_b[s]), ph ^ 1, abort_flag, 1, 256);
          mbar_expect_tx(smem_u32(&bar_full_b[s]), box_bytes);
          tma_load_4d(smem_u32(sB + s * kBStageBytes), &tmap, smem_u32(&bar_full_b[s]), 0, 0, t.X0 + c * g.cx, t.f0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_s8(kTileQ, g.n_chunk);
    uint32_t chunk = 0, tile = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      const TileSpan t = tile_span(g, item, f0_of, f1_of, spans);
      if (t.n_rows == 0) continue;
      if (lane == 0) {
        const uint32_t a = tile % kAStages, aph = (tile / kAStages) & 1;
        mbar_wait(smem_u32(&bar_full_a[a]), aph, abort_flag, 2, 128);
        const uint32_t a_addr = smem_u32(sA + a * kAStageBytes);
        for (int c = 0; c < t.n_chunks; c++, chunk++) {
          const uint32_t s = chunk % kBStages, ph = (chunk / kBStages) & 1;
          const uint32_t acc = chunk % (uint32_t)g.acc_stages, accph = (chunk / (uint32_t)g.acc_stages) & 1;
          mbar_wait(smem_u32(&bar_acc_empty[acc]), accph ^ 1, abort_flag, 3, 64);
          mbar_wait(smem_u32(&bar_full_b[s]), ph, abort_flag, 4);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(sB + s * kBStageBytes);
          const uint32_t d_addr = tmem_base + acc * (uint32_t)g.acc_stride;
          umma_s8(d_addr, umma_desc_k_sw64(a_addr), umma_desc_k_sw64(b_addr), idesc, 0);
          umma_s8(d_addr, umma_desc_k_sw64(a_addr + 32), umma_desc_k_sw64(b_addr + 32), idesc, 1);
          umma_commit(smem_u32(&bar_empty_b[s]));
          umma_commit(smem_u32(&bar_acc_full[acc]));
        }
        if (t.n_chunks > 0) umma_commit(smem_u32(&bar_empty_a[a]));
        else mbar_arrive(smem_u32(&bar_empty_a[a]));
      }
      __syncwarp();
      tile++;
    }
  } else if (warp < kEpiWarp0) {
    // ------------------------------------------------------------ query warps (one tile ahead)
    const int row = threadIdx.x - 64;
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      const TileSpan t = tile_span(g, item, f0_of, f1_of, spans);
      if (t.n_rows == 0) continue;
      const uint32_t a = tile % kAStages, aph = (tile / kAStages) & 1;
      mbar_wait(smem_u32(&bar_empty_a[a]), aph ^ 1, abort_flag, 5, 512);
#ifdef MV_TC_TRACE
      const long long cq0 = clock64();
#endif

      // validity words of the tile's cell range (plus the overrun of the last chunk)
      const int w_lo = (t.X0 * g.rows) >> 5;
      const int w_hi = min(g.vwords - 1, (((t.X0 + t.n_chunks * g.cx) * g.rows + 63) >> 5) + 1);
      const uint32_t* vsrc = vbits + (size_t)t.f0 * g.vwords;
      uint32_t* sv = sV + a * vstride;
      for (int w = w_lo + row; w <= w_hi; w += kQueryThreads) sv[w - w_lo] = vsrc[w];
      const int base_bit = w_lo << 5;

      // this row's query: A operand row and search window
      const bool active = row < t.n_rows;
      int cell1 = 0, x_lo = 0, x_hi = -1, y_lo = 0, y_hi = -1;
      int4 q4[4] = {make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0), make_int4(0, 0, 0, 0)};
      if (active) {
        cell1 = q_patch[(size_t)t.f1 * g.top_n + t.q0 + row];
        const int qx = cell1 / g.rows, qy = cell1 - qx * g.rows;
        x_lo = max(qx + g.shift_x - g.radius, 0);
        x_hi = min(qx + g.shift_x + g.radius, g.cols - 1);
        y_lo = max(qy + g.shift_y - g.radius, 0);
        y_hi = min(qy + g.shift_y + g.radius, g.rows - 1);
        if (y_hi < y_lo) { x_lo = 0; x_hi = -1; }
      }
      const int4* qp = reinterpret_cast<const int4*>(desc + ((size_t)t.f1 * g.cells + cell1) * 256);
      if (active) {
#pragma unroll
        for (int c = 0; c < 4; c++) q4[c] = __ldg(qp + c);
      }
      {
        uint8_t* dst = sA + a * kAStageBytes;
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<int4*>(dst + sw64_offset(row, c)) = q4[c];
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kQueryThreads) : "memory");   // validity words visible

      // Leading candidates (tracking_main.c:21-32): every valid window cell, in scan order, is
      // scored over 256 dims until one has a non-zero norm.  The search for the next cell is
      // per lane; the evaluation is warp-convergent so each round is one batch of loads.
      RowInfo ri;
      ri.xwin = (x_lo & 0xffff) | (x_hi << 16);
      ri.ywin = (y_lo & 0xffff) | (y_hi << 16);
      ri.lead_end = -1; ri.bcell = -1; ri.bs = 0.0f; ri.den_f = 1.0f; ri.curmax = 0x7fffffff; ri.flip = 0;
      {
        const int8_t* d0 = desc + (size_t)t.f0 * g.cells * 256;
        int n_cand = 0, nq64 = 0;
        int sx = x_lo, slo = x_lo * g.rows + y_lo;   // search cursor: column and first cell to test
        bool searching = x_hi >= x_lo;
        while (true) {
          int c = -1;
          while (searching && c < 0) {
            c = first_set_in_range(sv, base_bit, slo, sx * g.rows + y_hi);
            if (c < 0) {
              sx++;
              slo = sx * g.rows + y_lo;
              searching = sx <= x_hi;
            }
          }
          if (!__any_sync(0xffffffffu, c >= 0)) break;
          if (c >= 0) {
            const int4* cp = reinterpret_cast<const int4*>(d0 + (size_t)c * 256);
            int dot = 0, nc = 0, nq = 0;
#pragma unroll
            for (int kb = 0; kb < 16; kb += 8) {
              int4 cv[8], qv[8];
#pragma unroll
              for (int k = 0; k < 8; k++) { cv[k] = __ldg(cp + kb + k); qv[k] = __ldg(qp + kb + k); }
#pragma unroll
              for (int k = 0; k < 8; k++) {
                dot = dp4a4(cv[k], qv[k], dot);
                nc = dp4a4(cv[k], cv[k], nc);
                nq = dp4a4(qv[k], qv[k], nq);
                if (kb == 0 && k == 3) nq64 = nq;
              }
            }
            n_cand = nc;
            const float s = wrapped_cos2(dot, nc, nq);
            if (s > g.accept_gt && (ri.bcell < 0 || s > ri.bs)) { ri.bs = s; ri.bcell = c; }
            ri.lead_end = c;
            slo = c + 1;
            searching = n_cand == 0;   // sticky zero norm: the next valid cell is a leading one too
          }
        }
        if (n_cand != 0) {
          const int den_i = (int)((unsigned)n_cand * (unsigned)nq64);
          ri.den_f = __int2float_rn(den_i);
          // conservative start of the key filter: every key <= curmax has s <= accept_gt
          if (den_i > 0) {
            const double b = floor((double)g.accept_gt * (double)ri.den_f) - 256.0;
            ri.curmax = b < -2147483648.0 ? (int)0x80000000 : (b > 2147483647.0 ? 0x7fffffff : (int)b);
          } else if (den_i < 0) {
            ri.flip = -1;
            const double b = -ceil((double)g.accept_gt * (double)ri.den_f) - 257.0;
            ri.curmax = b < -2147483648.0 ? (int)0x80000000 : (b > 2147483647.0 ? 0x7fffffff : (int)b);
          } else {
            ri.curmax = 0;   // s = +inf only for n > 0
          }
        }
        // n_cand == 0: every valid candidate had a zero norm and was scored above
      }
      {
        int4* rdst = reinterpret_cast<int4*>(sR + a * kTileQ + row);
        rdst[0] = make_int4(ri.xwin, ri.ywin, ri.lead_end, ri.bcell);
        rdst[1] = make_int4(__float_as_int(ri.bs), __float_as_int(ri.den_f), ri.curmax, ri.flip);
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bar_full_a[a]));
#ifdef MV_TC_TRACE
      if (row == 0) MV_TC_TRACE_ADD(8, cq0);
#endif
      tile++;
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int ew = warp - kEpiWarp0;
    const int qd = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = ew >> 2;         // which of the row's kParts threads
    const int row = qd * 32 + lane;
    uint32_t chunk = 0, tile = 0;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
#ifdef MV_TC_TRACE
      const long long cs0 = clock64();
#endif
      const TileSpan t = tile_span(g, item, f0_of, f1_of, spans);
#ifdef MV_TC_TRACE
      if (ew == 0 && lane == 0) MV_TC_TRACE_ADD(11, cs0);
      const long long cp0 = clock64();
#endif
      if (t.n_rows == 0) continue;
      const int pair = item / g.tiles_per_pair;
      const uint32_t a = tile % kAStages, aph = (tile / kAStages) & 1;
      mbar_wait(smem_u32(&bar_full_a[a]), aph, abort_flag, 6);
      const uint32_t* sv = sV + a * vstride;
      const int base_bit = ((t.X0 * g.rows) >> 5) << 5;

      const bool active = row < t.n_rows;
      const int4 r0 = reinterpret_cast<const int4*>(sR + a * kTileQ + row)[0];
      const int4 r1 = reinterpret_cast<const int4*>(sR + a * kTileQ + row)[1];
      const int x_lo = r0.x & 0xffff, x_hi = r0.x >> 16, y_lo = r0.y & 0xffff, y_hi = r0.y >> 16;
      const int lead_end = r0.z;
      int bcell = r0.w;
      float bs = __int_as_float(r1.x);
      const float den_f = __int_as_float(r1.y);
      int curmax = r1.z;
      const int flip = r1.w;
      // warp-uniform column range of this warp's windows (blocks outside it are skipped outright)
      int wx_lo = x_hi >= x_lo ? x_lo : (1 << 20);
      int wx_hi = x_hi >= x_lo ? x_hi : -1;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        wx_lo = min(wx_lo, __shfl_xor_sync(0xffffffffu, wx_lo, o));
        wx_hi = max(wx_hi, __shfl_xor_sync(0xffffffffu, wx_hi, o));
      }

      // ---- the tile's chunks
#ifdef MV_TC_TRACE
      if (ew == 0 && lane == 0) MV_TC_TRACE_ADD(13, cp0);
#endif
      for (int c = 0; c < t.n_chunks; c++, chunk++) {
        const uint32_t acc = chunk % (uint32_t)g.acc_stages, accph = (chunk / (uint32_t)g.acc_stages) & 1;
#ifdef MV_TC_TRACE
        const long long cw0 = clock64();
#endif
        mbar_wait(smem_u32(&bar_acc_full[acc]), accph, abort_flag, 7);
#ifdef MV_TC_TRACE
        if (ew == 0 && lane == 0) MV_TC_TRACE_ADD(14, cw0);
#endif
#ifdef MV_TC_TRACE
        const long long ce0 = clock64();
#endif
        tc_fence_after();
        const int chunk_x0 = t.X0 + c * g.cx;
        const int chunk_cell = chunk_x0 * g.rows;
        const int col_limit = g.cx * g.rows;
        // columns of this chunk some window of the warp can touch: [wc_lo, wc_hi)
        const int wc_lo = max(0, (wx_lo - chunk_x0) * g.rows);
        const int wc_hi = min(col_limit, (wx_hi - chunk_x0 + 1) * g.rows);
        // this thread's blocks: half, half+kParts, ... restricted to that range
        int col = (wc_lo / (32 * kParts)) * (32 * kParts) + half * 32;
        if (col + 32 <= wc_lo) col += 32 * kParts;
        int x0b = chunk_x0 + col / g.rows, y0b = col - (col / g.rows) * g.rows;   // cell coords of column `col`
        const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16) + acc * (uint32_t)g.acc_stride;
        for (; col < wc_hi; col += 32 * kParts) {
          const int cb = chunk_cell + col;
          // validity of the block's 32 cells (uniform), minus the padding columns
          const int o = cb - base_bit;
          uint32_t vm = __funnelshift_r(sv[o >> 5], sv[(o >> 5) + 1], o & 31);
          if (col_limit - col < 32) vm &= (1u << (col_limit - col)) - 1u;
          // this row's window inside the block
          uint32_t wm = 0;
          {
            int x = x0b, off = -y0b;
            while (off < 32) {
              if (x >= x_lo && x <= x_hi) {
                const int lo = max(off + y_lo, 0), hi = min(off + y_hi, 31);
                if (lo <= hi) wm |= (0xffffffffu >> (31 - (hi - lo))) << lo;
              }
              x++;
              off += g.rows;
            }
          }
          uint32_t m = wm & vm;
          if (lead_end >= cb) m = (lead_end - cb >= 31) ? 0u : (m & ~((2u << (lead_end - cb)) - 1u));
          if (__any_sync(0xffffffffu, m != 0)) {
            int v[32];
            tmem_ld_32x32(t_row + col, v);
            tmem_ld_wait();
            // branch-free: which in-window, valid columns exceed the key filter?
            // key > curmax  <=>  n > curmax (flip = 0)  /  n < ~curmax (flip = -1); the flipped
            // lanes test n <= ~curmax, a superset the exact path re-checks: one compare with the
            // lane's flip folded in as a predicate, no per-column xor.
            const bool flipb = flip != 0;
            const int cm = flipb ? ~curmax : curmax;
            uint32_t trig = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const int n = (int)((unsigned)v[j] * (unsigned)v[j]);
              trig |= ((n > cm) != flipb) ? (1u << j) : 0u;
            }
            trig &= m;
            // survivors (rare): exact score, in column order, one TMEM column at a time
            uint32_t any = __reduce_or_sync(0xffffffffu, trig);
            while (any) {
              const int j = __ffs(any) - 1;
              any &= any - 1;
              int vj;
              tmem_ld_32x1(t_row + col + j, vj);
              tmem_ld_wait();
              if (trig & (1u << j)) {
                const int n = (int)((unsigned)vj * (unsigned)vj);
                const int key = n ^ flip;
                if (key > curmax) {
                  curmax = key;
                  const float s = __fdiv_rn(__int2float_rn(n), den_f);
                  if (s > g.accept_gt && (bcell < 0 || s > bs)) { bs = s; bcell = cb + j; }
                }
              }
            }
          }
          y0b += 32 * kParts;
          while (y0b >= g.rows) { y0b -= g.rows; x0b++; }
        }
        tc_fence_before();
        __syncwarp();
#ifdef MV_TC_TRACE
        if (lane == 0) MV_TC_TRACE_ADD(9 + (ew == 0 ? 0 : 1), ce0);
#endif
        if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
      }

      // ---- every one of the row's kParts threads reports its own best candidate; the emit kernel
      // merges them (larger score, ties to the earlier cell).  No barrier between the epilogue warps:
      // a warp that is done with the tile moves on to the next one.
#ifdef MV_TC_TRACE
      const long long cm0 = clock64();
#endif
      if (active) {
        const size_t out = (size_t)half * part_stride + (size_t)pair * g.top_n + t.q0 + row;
        best_cell[out] = bcell;
        best_score[out] = bs;
      }
      __syncwarp();
#ifdef MV_TC_TRACE
      if (ew == 0 && lane == 0) MV_TC_TRACE_ADD(12, cm0);
#endif
      if (lane == 0) mbar_arrive(smem_u32(&bar_empty_a[a]));
      tile++;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
#ifdef MV_TC_TRACE
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(&mv_tc_trace[15], 1ull) + 1;
    if (done == gridDim.x) {
      printf("tc trace (Mcycles summed over grid): empty_b %.2f full_a(mma) %.2f acc_empty %.2f full_b %.2f empty_a(query) %.2f "
             "full_a(epi, x256 thr) %.2f acc_full(epi, x256 thr) %.2f | query tile body %.2f epi chunk w6 %.2f epi chunk others(x7) %.2f | epi w6: tile_span %.2f merge+store %.2f prechunk(full_a wait, row info) %.2f acc_full wait w6 %.2f items %d\n",
             mv_tc_trace[1] * 1e-6, mv_tc_trace[2] * 1e-6, mv_tc_trace[3] * 1e-6, mv_tc_trace[4] * 1e-6, mv_tc_trace[5] * 1e-6 / 128,
             mv_tc_trace[6] * 1e-6 / 256, mv_tc_trace[7] * 1e-6 / 256, mv_tc_trace[8] * 1e-6, mv_tc_trace[9] * 1e-6,
             mv_tc_trace[10] * 1e-6 / 7, mv_tc_trace[11] * 1e-6, mv_tc_trace[12] * 1e-6, mv_tc_trace[13] * 1e-6, mv_tc_trace[14] * 1e-6, g.n_items);
      for (int i = 0; i < 16; i++) mv_tc_trace[i] = 0;
    }
  }
#endif
}

}  // namespace

int mv_match_tc_parts() { return kParts; }

mv_status mv_match_tc_launch(mv_ctx* ctx, const mv_match_params* p, int n_frames, int n_pairs, int top_n,
                             const int32_t* d_f0, const int32_t* d_f1, const int8_t* d_desc,
                             const int32_t* d_max_idx, const float* d_prob, const int32_t* d_q_patch,
                             const int32_t* d_q_count, int32_t* d_best_cell, float* d_best_score) {
  // d_best_cell / d_best_score: [mv_match_tc_parts()][n_pairs][top_n], one candidate per epilogue part
  if (n_frames <= 0) MV_BAD_ARG(ctx, "tensor-core matcher: n_frames must be given");
  if (p->rows > 256) MV_BAD_ARG(ctx, "tensor-core matcher: rows <= 256 (one cell column per TMA box row block)");
  const double thr2 = p->match_threshold * p->match_threshold;
  if (!(thr2 >= 0.0)) MV_BAD_ARG(ctx, "tensor-core matcher: match_threshold^2 must be >= 0");
  mv_tmap_encode_fn encode = mv_get_tmap_encode();
  if (!encode) {
    snprintf(ctx->err, sizeof(ctx->err), "cuTensorMapEncodeTiled not available from the driver");
    return MV_ERR_CUDA;
  }
  TcGeom g;
  g.rows = p->rows; g.cols = p->cols; g.cells = p->rows * p->cols;
  g.shift_x = p->shift_x; g.shift_y = p->shift_y; g.radius = p->radius; g.top_n = top_n;
  g.cx = 256 / p->rows;
  if (const char* e = getenv("MV_TC_CX")) {   // cell columns per chunk (A/B knob; any value gives the same bytes)
    const int v = atoi(e);
    if (v >= 1 && v * p->rows <= 256) g.cx = v;
  }
  if (g.cx > p->cols) g.cx = p->cols;
  g.n_chunk = (g.cx * p->rows + 15) & ~15;
  g.acc_stride = (g.n_chunk + 31) & ~31;
  g.acc_stages = kTmemCols / g.acc_stride < kMaxAccStages ? kTmemCols / g.acc_stride : kMaxAccStages;
  g.vwords = (g.cells + 31) / 32 + kVPadWords + (g.cx * p->rows + 31) / 32;
  g.tiles_per_pair = (top_n + kTileQ - 1) / kTileQ;
  g.n_items = n_pairs * g.tiles_per_pair;
  g.accept_gt = mv_round_down(thr2);
  g.prob_lt = mv_round_up(p->min_prob0);

  void* vb = nullptr; int* flag = nullptr;
  mv_status st = mv_scratch(ctx, "match.vbits", sizeof(uint32_t) * (size_t)n_frames * g.vwords, &vb);
  if (st) return st;
  st = mv_abort_flag(ctx, &flag);
  if (st) return st;
  void* spans = nullptr;
  st = mv_scratch(ctx, "match.tc_spans", sizeof(int4) * (size_t)g.n_items, &spans);
  if (st) return st;

  CUtensorMap tmap;
  const cuuint64_t dims[4] = {256, (cuuint64_t)p->rows, (cuuint64_t)p->cols, (cuuint64_t)n_frames};
  const cuuint64_t strides[3] = {256, (cuuint64_t)p->rows * 256, (cuuint64_t)g.cells * 256};
  const cuuint32_t box[4] = {64, (cuuint32_t)p->rows, (cuuint32_t)g.cx, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(d_desc), dims, strides, box,
                             es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(ctx->err, sizeof(ctx->err), "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return MV_ERR_CUDA;
  }

  mv_prof_scope ps(ctx, "match");
  MV_CUDA(ctx, cudaMemsetAsync(vb, 0, sizeof(uint32_t) * (size_t)n_frames * g.vwords, ctx->stream));
  {
    dim3 grid((g.cells + 255) / 256, n_frames);
    valid_bits_kernel<<<grid, 256, 0, ctx->stream>>>(g.cells, g.vwords, g.prob_lt, d_max_idx, d_prob, (uint32_t*)vb);
    MV_CHECK_LAUNCH(ctx);
  }
  tile_spans_kernel<<<(g.n_items + 7) / 8, 256, 0, ctx->stream>>>(g, d_f0, d_f1, d_q_patch, d_q_count, (int4*)spans);
  MV_CHECK_LAUNCH(ctx);
  const size_t smem = 1024 + (size_t)kBStages * kBStageBytes + (size_t)kAStages * kAStageBytes +
                      sizeof(uint32_t) * (size_t)kAStages * ((g.vwords + 3) & ~3) +
                      sizeof(RowInfo) * kAStages * kTileQ;
  if (smem > 227 * 1024) MV_BAD_ARG(ctx, "tensor-core matcher: grid too large for the shared-memory validity window");
  MV_CUDA(ctx, cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = g.n_items < ctx->sm_count ? g.n_items : ctx->sm_count;
  match_tc_kernel<<<grid, kThreads, smem, ctx->stream>>>(tmap, g, d_f0, d_f1, d_desc, (const uint32_t*)vb, d_q_patch,
                                                        d_q_count, (const int4*)spans, d_best_cell, d_best_score, (size_t)n_pairs * (size_t)top_n, flag);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}
