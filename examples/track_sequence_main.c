/* track_sequence_main.c -- a C99 host program on the batched C ABI (include/maveric_b200.h).
 *
 * The reference's hot-path driver is src/tracking_main.c: one main() that runs one frame pair
 * from a compiled-in header.  This is the same driver for a sequence: frames come from a file,
 * every pair (f, f+1) goes through detector -> windowed match -> RANSAC-E -> Gauss-Newton PnP
 * in one call, and one 64-byte record per pair comes back.  Plain C, host buffers only (no CUDA
 * header, no C++ type): the library stages the frames itself (mv_track_sequence_host).
 *
 *   track_sequence <frames.bin> <results.bin> [top_n max_valid max_matches hypotheses]
 *
 * frames.bin:  int32 n_frames, rows, cols, 0
 *              float  semi_scale[n_frames]
 *              int8   semi [n_frames][rows*cols][65]     cells column-major (patch = x*rows + y)
 *              int8   desc [n_frames][rows*cols][256]
 *              float  depth[n_frames][rows*cols]
 * results.bin: mv_pair_result[n_frames - 1]
 * stdout:      one line per pair, as tracking_main.c prints its counts and pose
 *
 * Build (what `make -C examples` runs):
 *   gcc -std=c99 -Wall -I../include track_sequence_main.c -L../maveric-slam_b200 -lmaveric_b200 \
 *       -Wl,-rpath,'$ORIGIN/../maveric-slam_b200' -o track_sequence
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "maveric_b200.h"

static int fail(const char* what) {
  fprintf(stderr, "track_sequence: %s\n", what);
  return 2;
}

static void* read_block(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (p && fread(p, 1, bytes, f) != bytes) {
    free(p);
    p = NULL;
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 3 && argc != 7)
    return fail("usage: track_sequence <frames.bin> <results.bin> [top_n max_valid max_matches hypotheses]");
  FILE* f = fopen(argv[1], "rb");
  if (!f) return fail("cannot open the frames file");
  int32_t hdr[4];
  if (fread(hdr, sizeof(int32_t), 4, f) != 4 || hdr[0] < 2 || hdr[1] <= 0 || hdr[2] <= 0) {
    fclose(f);
    return fail("bad header (need n_frames >= 2, rows > 0, cols > 0)");
  }
  const int n_frames = hdr[0], rows = hdr[1], cols = hdr[2];
  const size_t cells = (size_t)rows * cols, n = (size_t)n_frames;
  float* scale = (float*)read_block(f, sizeof(float) * n);
  int8_t* semi = (int8_t*)read_block(f, n * cells * 65);
  int8_t* desc = (int8_t*)read_block(f, n * cells * 256);
  float* depth = (float*)read_block(f, sizeof(float) * n * cells);
  fclose(f);
  if (!scale || !semi || !desc || !depth) return fail("frames file is shorter than its header says");

  mv_ctx* ctx = NULL;
  mv_status st = mv_ctx_create(0, &ctx);
  if (st != MV_OK) return fail(mv_status_str(st)); /* no sm_100 GPU: there is no CPU fallback */

  mv_track_params p;
  mv_track_params_default(&p, rows, cols); /* tracking_main.c's constants: N=100, r=4, shift (4,4), 150 matches */
  if (argc == 7) {
    p.top_n = atoi(argv[3]);
    p.max_valid = atoi(argv[4]);
    p.match.max_matches = atoi(argv[5]);
    p.pnp.hypotheses = atoi(argv[6]);
  }

  mv_pair_result* res = (mv_pair_result*)calloc(n - 1, sizeof(*res));
  unsigned long long up = 0, down = 0;
  st = mv_track_sequence_host(ctx, &p, n_frames, semi, scale, desc, depth, res, &up, &down);
  if (st != MV_OK) {
    fprintf(stderr, "track_sequence: %s: %s\n", mv_status_str(st), mv_last_error(ctx));
    mv_ctx_destroy(ctx);
    return 1;
  }
  for (int i = 0; i < n_frames - 1; i++)
    printf("pair %d: num_matches = %d, num_inliers = %d, pnp_inliers = %d, q = [%g %g %g %g], t = [%g %g %g]%s\n", i,
           res[i].num_matches, res[i].ransac_inliers, (int)res[i].pnp_inliers, res[i].q[0], res[i].q[1], res[i].q[2],
           res[i].q[3], res[i].t[0], res[i].t[1], res[i].t[2], res[i].status ? "  (max_valid overflow)" : "");
  fprintf(stderr, "track_sequence: %d pairs, %llu bytes to the GPU, %llu back, %llu kernel launches\n", n_frames - 1, up,
          down, mv_ctx_launch_count(ctx));
  mv_ctx_destroy(ctx);

  f = fopen(argv[2], "wb");
  if (!f || fwrite(res, sizeof(*res), n - 1, f) != n - 1) return fail("cannot write the results file");
  fclose(f);
  free(res); free(scale); free(semi); free(desc); free(depth);
  return 0;
}
