// synth.cu -- device-side generator of the synthetic KITTI-shaped frames used by
// bench.py and the tests.  Same counter-based definition as maveric-slam_b200/synth.py
// (byte-identical output): world-anchored keypoints and base descriptors, per-frame
// logit background and descriptor noise.  Layout per frame is the reference's
// (python/superpoint_inference.py:630-664): semi [cells][65], desc [cells][256],
// cell = col*rows + row.
#include "mv_common.cuh"

namespace {

__device__ __forceinline__ int desc_lut(unsigned u) {
  const int tail[10] = {-128, -111, -97, -85, -70, 70, 85, 97, 111, 127};
  return u < 246u ? (int)(u % 67u) - 33 : tail[u - 246u];
}

// one thread per (frame, cell, 8-byte group); groups 0..31 = descriptor, 32..40 = logits
__global__ void synth_frames_kernel(unsigned long long ms, int rows, int cols, int permille, int amp,
                                    int first_frame, int n_frames, const int32_t* __restrict__ off,
                                    int8_t* __restrict__ semi, int8_t* __restrict__ desc,
                                    float* __restrict__ depth) {
  const int cells = rows * cols;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_frames * cells * 41;
  if (tid >= total) return;
  const int g = (int)(tid % 41);
  const long long fc = tid / 41;
  const int cell = (int)(fc % cells);
  const int fl = (int)(fc / cells);
  const int frame = first_frame + fl;
  const int x = cell / rows, y = cell - x * rows;
  const unsigned long long wx = (unsigned long long)(long long)(x + off[2 * fl] + (1 << 23));
  const unsigned long long wy = (unsigned long long)(long long)(y + off[2 * fl + 1] + (1 << 23));

  if (g < 32) {
    const unsigned long long hb = mv_ctr(ms, 3, wx, wy, (unsigned long long)g);
    const unsigned long long hn = mv_ctr(ms, 4, (unsigned long long)frame, (unsigned long long)cell,
                                         (unsigned long long)g);
    unsigned long long packed = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int base = desc_lut((unsigned)((hb >> (8 * b)) & 0xFF));
      const int noise = amp > 0 ? (int)(((hn >> (8 * b)) & 0xFF) % (unsigned)(2 * amp + 1)) - amp : 0;
      int v = base + noise;
      v = v > 127 ? 127 : (v < -128 ? -128 : v);
      packed |= (unsigned long long)(unsigned char)(signed char)v << (8 * b);
    }
    *reinterpret_cast<unsigned long long*>(desc + ((size_t)fl * cells + cell) * 256 + g * 8) = packed;
    return;
  }
  const int sg = g - 32;
  const unsigned long long k = mv_ctr(ms, 1, wx, wy, 0);
  const bool is_kp = (int)(k % 1000ull) < permille;
  const int kp_ch = (int)((k >> 16) & 63);
  const int kp_val = 14 + (int)((k >> 24) % 29ull);
  const int kp_dust = -20 + (int)((k >> 32) % 45ull);
  const int bg_dust = 10 + (int)((k >> 32) % 30ull);
  if (sg == 0 && depth)
    depth[(size_t)fl * cells + cell] =
        __fadd_rn(4.0f, __fmul_rn((float)((k >> 40) & 0xFFFF), 36.0f / 65536.0f));
  const unsigned long long h = mv_ctr(ms, 2, (unsigned long long)frame, (unsigned long long)cell,
                                      (unsigned long long)sg);
  int8_t* srow = semi + ((size_t)fl * cells + cell) * 65;
  for (int b = 0; b < 8; b++) {
    const int ch = sg * 8 + b;
    if (ch >= 65) break;
    const unsigned u = (unsigned)((h >> (8 * b)) & 0xFF);
    int v = u == 255u ? (ch % 6) * 3 : -30 - (int)(u % 70u);
    if (ch == 64) v = is_kp ? kp_dust : bg_dust;
    else if (is_kp && ch == kp_ch) v = kp_val;
    srow[ch] = (int8_t)v;
  }
}

}  // namespace

extern "C" mv_status mv_synth_frames(mv_ctx* ctx, const mv_synth_params* p, int first_frame, int n_frames,
                                     const int32_t* d_off, int8_t* d_semi, int8_t* d_desc, float* d_depth) {
  MV_ENTER(ctx);
  if (!p || n_frames <= 0 || !d_off || !d_semi || !d_desc || p->rows <= 0 || p->cols <= 0)
    MV_BAD_ARG(ctx, "mv_synth_frames");
  if (reinterpret_cast<uintptr_t>(d_desc) & 7) MV_BAD_ARG(ctx, "mv_synth_frames: d_desc must be 8-byte aligned");
  const long long total = (long long)n_frames * p->rows * p->cols * 41;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) MV_BAD_ARG(ctx, "mv_synth_frames: too many frames for one launch");
  mv_prof_scope ps(ctx, "synth");
  synth_frames_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(
      mv_sm64(p->seed), p->rows, p->cols, p->keypoint_permille, p->noise_amp, first_frame, n_frames, d_off,
      d_semi, d_desc, d_depth);
  MV_CHECK_LAUNCH(ctx);
  return MV_OK;
}
