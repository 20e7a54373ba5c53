// device_utils.cuh -- small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// image index n (0..N-1, mode k = n - N/2) -> fine-grid index k mod nf
__device__ __forceinline__ int mode_to_fine(int n, int N, int nf) {
  int k = n - N / 2;
  return k < 0 ? k + nf : k;
}
// fine-grid index l -> image index n or -1 when l is outside the kept modes
__device__ __forceinline__ int fine_to_mode(int l, int N, int nf) {
  if (l < N - N / 2) return l + N / 2;
  if (l >= nf - N / 2) return l - nf + N / 2;
  return -1;
}
