// svd3.cuh -- 3x3 SVD with approximate Givens rotations and a bit-trick reciprocal
// square root, restating what the reference's vendored include/svd/svd.h computes
// (McAdams, Selle, Tamstorf, Teran, Sifakis, TR1690; svd.h:358-405) so that
// recover_pose_from_essential_matrix (src/pnp_solver.c:168-194) returns the same
// bits on host and device.  Every fp32 operation is an explicit RN intrinsic in the
// reference's evaluation order; compiled for both host and device.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace mvsvd {

#define MVF __host__ __device__ __forceinline__
// Device: explicit round-to-nearest intrinsics (never contracted into FMA).
// Host: plain fp32 ops pinned through a volatile so the host compiler cannot fuse them.
MVF float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
MVF float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
MVF float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
MVF float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
MVF int f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}
MVF float i2f(int i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}
// 5.828427124 is a double literal in svd.h:22, so this comparison is double math
MVF bool gamma_lt(float sh, float ch2) { return 5.828427124 * (double)sh * (double)sh < (double)ch2; }

struct Sym3 { float d0, o10, d1, o20, o21, d2; };

// svd.h:37-48
MVF float rsqrt_one(float x) {
  const float half = fmul(0.5f, x);
  const float y = i2f(0x5f375a82 - (f2i(x) >> 1));
  return fmul(y, fsub(1.5f, fmul(fmul(half, y), y)));
}
// svd.h:54-62
MVF float rsqrt_two(float x) {
  const float half = fmul(0.5f, x);
  float y = i2f(0x5f37599e - (f2i(x) >> 1));
  y = fmul(y, fsub(1.5f, fmul(fmul(half, y), y)));
  y = fmul(y, fsub(1.5f, fmul(fmul(half, y), y)));
  return y;
}

// svd.h:148-216, one Jacobi step on the leading 2x2 block + cyclic relabel
MVF void jacobi_step(int ax, int ay, int az, Sym3& s, float q[4]) {
  float ch = fmul(2.0f, fsub(s.d0, s.d1));
  float sh = s.o10;
  const float ch2 = fmul(ch, ch), sh2 = fmul(sh, sh);
  const bool keep = gamma_lt(sh, ch2);
  const float w = rsqrt_one(fadd(ch2, sh2));
  ch = keep ? fmul(w, ch) : (float)0.923879532;
  sh = keep ? fmul(w, sh) : (float)0.3826834323;

  const float cc = fmul(ch, ch), ss = fmul(sh, sh);
  const float nrm = fadd(cc, ss);
  const float a = fdiv(fsub(cc, ss), nrm);
  const float b = fdiv(fmul(fmul(2.0f, sh), ch), nrm);
  const float nb = -b;

  const float p00 = s.d0, p10 = s.o10, p11 = s.d1, p20 = s.o20, p21 = s.o21, p22 = s.d2;
  const float t0 = fadd(fmul(a, p00), fmul(b, p10));     // a*s11 + b*s21
  const float t1 = fadd(fmul(a, p10), fmul(b, p11));     // a*s21 + b*s22
  const float t2 = fadd(fmul(nb, p00), fmul(a, p10));    // -b*s11 + a*s21
  const float t3 = fadd(fmul(nb, p10), fmul(a, p11));    // -b*s21 + a*s22
  const float n00 = fadd(fmul(a, t0), fmul(b, t1));
  const float n10 = fadd(fmul(a, t2), fmul(b, t3));
  const float n11 = fadd(fmul(nb, t2), fmul(a, t3));
  const float n20 = fadd(fmul(a, p20), fmul(b, p21));
  const float n21 = fadd(fmul(nb, p20), fmul(a, p21));
  const float n22 = p22;

  float tv[3] = {fmul(q[0], sh), fmul(q[1], sh), fmul(q[2], sh)};
  sh = fmul(sh, q[3]);
  q[0] = fmul(q[0], ch); q[1] = fmul(q[1], ch); q[2] = fmul(q[2], ch); q[3] = fmul(q[3], ch);
  q[az] = fadd(q[az], sh);
  q[3] = fsub(q[3], tv[az]);
  q[ax] = fadd(q[ax], tv[ay]);
  q[ay] = fsub(q[ay], tv[ax]);

  s.d0 = n11;
  s.o10 = n21; s.d1 = n22;
  s.o20 = n10; s.o21 = n20; s.d2 = n00;
}

MVF float sq3(float x, float y, float z) { return fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)); }

MVF void neg_swap_cols(bool c, float M[3][3], int i, int j) {
  for (int r = 0; r < 3; r++) {
    const float z = -M[r][i];
    M[r][i] = c ? M[r][j] : M[r][i];
    M[r][j] = c ? z : M[r][j];
  }
}

// svd.h:277-291
MVF void qr_givens(float pivot, float below, float& ch, float& sh) {
  const float eps = (float)1e-6;
  const float sum = fadd(fmul(pivot, pivot), fmul(below, below));
  const float rho = fmul(sum, rsqrt_two(sum));
  float s = rho > eps ? below : 0.0f;
  float c = fadd(fabsf(pivot), fmaxf(rho, eps));
  if (pivot < 0.0f) { const float t = s; s = c; c = t; }
  const float w = rsqrt_one(fadd(fmul(c, c), fmul(s, s)));
  ch = fmul(c, w);
  sh = fmul(s, w);
}

MVF float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return fadd(fadd(fmul(a0, b0), fmul(a1, b1)), fmul(a2, b2));
}

// svd.h:358-405
MVF void svd3(const float A[3][3], float U[3][3], float S[3][3], float V[3][3]) {
  float ata[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) ata[i][j] = dot3(A[0][i], A[0][j], A[1][i], A[1][j], A[2][i], A[2][j]);
  Sym3 s = {ata[0][0], ata[1][0], ata[1][1], ata[2][0], ata[2][1], ata[2][2]};
  float q[4] = {0.0f, 0.0f, 0.0f, 1.0f};
  for (int sweep = 0; sweep < 4; sweep++) {
    jacobi_step(0, 1, 2, s, q);
    jacobi_step(1, 2, 0, s, q);
    jacobi_step(2, 0, 1, s, q);
  }
  {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z), xz = fmul(x, z), xy = fmul(x, y),
                yz = fmul(y, z), wx = fmul(w, x), wy = fmul(w, y), wz = fmul(w, z);
    V[0][0] = fsub(1.0f, fmul(2.0f, fadd(yy, zz))); V[0][1] = fmul(2.0f, fsub(xy, wz)); V[0][2] = fmul(2.0f, fadd(xz, wy));
    V[1][0] = fmul(2.0f, fadd(xy, wz)); V[1][1] = fsub(1.0f, fmul(2.0f, fadd(xx, zz))); V[1][2] = fmul(2.0f, fsub(yz, wx));
    V[2][0] = fmul(2.0f, fsub(xz, wy)); V[2][1] = fmul(2.0f, fadd(yz, wx)); V[2][2] = fsub(1.0f, fmul(2.0f, fadd(xx, yy)));
  }
  float B[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i][j] = dot3(A[i][0], V[0][j], A[i][1], V[1][j], A[i][2], V[2][j]);
  {
    float r0 = sq3(B[0][0], B[1][0], B[2][0]);
    float r1 = sq3(B[0][1], B[1][1], B[2][1]);
    float r2 = sq3(B[0][2], B[1][2], B[2][2]);
    bool c = r0 < r1;
    neg_swap_cols(c, B, 0, 1); neg_swap_cols(c, V, 0, 1);
    if (c) { const float t = r0; r0 = r1; r1 = t; }
    c = r0 < r2;
    neg_swap_cols(c, B, 0, 2); neg_swap_cols(c, V, 0, 2);
    if (c) { const float t = r0; r0 = r2; r2 = t; }
    c = r1 < r2;
    neg_swap_cols(c, B, 1, 2); neg_swap_cols(c, V, 1, 2);
  }
  float ch1, sh1, ch2, sh2, ch3, sh3, a, b;
  float R[3][3], T[3][3];
  qr_givens(B[0][0], B[1][0], ch1, sh1);
  a = fsub(1.0f, fmul(fmul(2.0f, sh1), sh1));
  b = fmul(fmul(2.0f, ch1), sh1);
  for (int j = 0; j < 3; j++) {
    R[0][j] = fadd(fmul(a, B[0][j]), fmul(b, B[1][j]));
    R[1][j] = fadd(fmul(-b, B[0][j]), fmul(a, B[1][j]));
    R[2][j] = B[2][j];
  }
  qr_givens(R[0][0], R[2][0], ch2, sh2);
  a = fsub(1.0f, fmul(fmul(2.0f, sh2), sh2));
  b = fmul(fmul(2.0f, ch2), sh2);
  for (int j = 0; j < 3; j++) {
    T[0][j] = fadd(fmul(a, R[0][j]), fmul(b, R[2][j]));
    T[1][j] = R[1][j];
    T[2][j] = fadd(fmul(-b, R[0][j]), fmul(a, R[2][j]));
  }
  qr_givens(T[1][1], T[2][1], ch3, sh3);
  a = fsub(1.0f, fmul(fmul(2.0f, sh3), sh3));
  b = fmul(fmul(2.0f, ch3), sh3);
  for (int j = 0; j < 3; j++) {
    S[0][j] = T[0][j];
    S[1][j] = fadd(fmul(a, T[1][j]), fmul(b, T[2][j]));
    S[2][j] = fadd(fmul(-b, T[1][j]), fmul(a, T[2][j]));
  }
  // svd.h:341-355 (products associate left to right)
  const float s1 = fmul(sh1, sh1), s2 = fmul(sh2, sh2), s3 = fmul(sh3, sh3);
  const float m1 = fadd(-1.0f, fmul(2.0f, s1));   // -1 + 2*sh12
  const float m2 = fadd(-1.0f, fmul(2.0f, s2));
  const float m3 = fadd(-1.0f, fmul(2.0f, s3));
  const float p2 = fsub(1.0f, fmul(2.0f, s2));    // 1 - 2*sh22
  U[0][0] = fmul(m1, m2);
  U[0][1] = fadd(fmul(fmul(fmul(fmul(fmul(4.0f, ch2), ch3), m1), sh2), sh3),
                 fmul(fmul(fmul(2.0f, ch1), sh1), m3));
  U[0][2] = fsub(fmul(fmul(fmul(fmul(4.0f, ch1), ch3), sh1), sh3),
                 fmul(fmul(fmul(fmul(2.0f, ch2), m1), sh2), m3));
  U[1][0] = fmul(fmul(fmul(2.0f, ch1), sh1), p2);
  U[1][1] = fadd(fmul(fmul(fmul(fmul(fmul(fmul(-8.0f, ch1), ch2), ch3), sh1), sh2), sh3), fmul(m1, m3));
  U[1][2] = fadd(fmul(fmul(-2.0f, ch3), sh3),
                 fmul(fmul(4.0f, sh1),
                      fadd(fmul(fmul(ch3, sh1), sh3), fmul(fmul(fmul(ch1, ch2), sh2), m3))));
  U[2][0] = fmul(fmul(2.0f, ch2), sh2);
  U[2][1] = fmul(fmul(fmul(2.0f, ch3), p2), sh3);
  U[2][2] = fmul(m2, m3);
}

// pnp_solver.c:168-194
MVF void recover_pose(const float E[3][3], float R1[3][3], float R2[3][3], float t[3]) {
  float U[3][3], S[3][3], V[3][3];
  svd3(E, U, S, V);
  const float W[3][3] = {{0, -1, 0}, {1, 0, 0}, {0, 0, 1}};
  const float Wt[3][3] = {{0, 1, 0}, {-1, 0, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R1[i][j] = dot3(U[i][0], W[0][j], U[i][1], W[1][j], U[i][2], W[2][j]);
      R2[i][j] = dot3(U[i][0], Wt[0][j], U[i][1], Wt[1][j], U[i][2], Wt[2][j]);
    }
  for (int i = 0; i < 3; i++) t[i] = U[i][2];
}

// pnp_solver.c:89-105
MVF float reproj_error(float x1, float y1, float x2, float y2, const float E[3][3]) {
  const float h1[3] = {x1, y1, 1.0f}, h2[3] = {x2, y2, 1.0f};
  float err = 0.0f;
  for (int i = 0; i < 3; i++) {
    const float proj = fadd(fadd(fmul(E[i][0], h1[0]), fmul(E[i][1], h1[1])), fmul(E[i][2], h1[2]));
    const float d = fsub(proj, h2[i]);
    err = fadd(err, fmul(d, d));
  }
  return err;
}

}  // namespace mvsvd
