// ffma_peak.cu -- FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on this GPU: 8 or 16 independent
// accumulator chains per thread, 4..32 warps per SM.  Result on the B200 pool (profiles/r01/ffma_peak.txt):
// both forms saturate at ~72-73 TFLOP/s, i.e. FFMA2 halves issue slots but not FMA-pipe time.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c){ f2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int CH> __global__ void k1(float* out, int iters, float a, float b) {
  float acc[CH];
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fma1(acc[i], a, b);
  float s = 0; for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __global__ void k2(float* out, int iters, float a, float b) {
  f2 acc[CH];
  f2 A, B; asm("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a), "f"(a)); asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b), "f"(b));
  for (int i = 0; i < CH; i++) asm("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"((float)threadIdx.x + i), "f"((float)i));
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fma2(acc[i], A, B);
  float s = 0; for (int i = 0; i < CH; i++) { float l, h; asm("mov.b64 {%0,%1}, %2;" : "=f"(l), "=f"(h) : "l"(acc[i])); s += l + h; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no CUDA device\n"); return 1; }
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = 128, blocks = 148 * (warps * 32 / threads);
    float m1 = timeit([&] { k1<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float m2 = timeit([&] { k2<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float m3 = timeit([&] { k1<16><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float m4 = timeit([&] { k2<16><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    double l1 = (double)blocks * threads * iters * 8 / (m1 * 1e-3) / 1e12 * 2, l2 = (double)blocks * threads * iters * 8 * 2 / (m2 * 1e-3) / 1e12 * 2;
    double l3 = (double)blocks * threads * iters * 16 / (m3 * 1e-3) / 1e12 * 2, l4 = (double)blocks * threads * iters * 16 * 2 / (m4 * 1e-3) / 1e12 * 2;
    printf("warps/SM=%2d  FFMA x8: %.1f TF  FFMA2 x8: %.1f TF  FFMA x16: %.1f TF  FFMA2 x16: %.1f TF\n", warps, l1, l2, l3, l4);
  }
  return 0;
}
