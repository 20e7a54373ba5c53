// DFMA issue rate of one SM sub-partition on sm_100a: 32 independent accumulators per thread, the operand pattern
// of double_rows.cu's visit loop (4 multipliers x 8 multiplicands).  nvcc -arch=sm_100a -O3 -o dfma_rate dfma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int WARPS_PER_SMSP>
__global__ void __launch_bounds__(128 * WARPS_PER_SMSP) k(double* out, int iters, double a0, double a1, double e0) {
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = i;
  double a[4] = {a0, a1, a0 + 1, a1 + 1}, e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = e0 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[c * 4 + r] = fma(a[r], e[c], acc[c * 4 + r]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W>
void run(int sms) {
  double* out;
  cudaMalloc(&out, (size_t)sms * 128 * W * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<W><<<sms, 128 * W>>>(out, 100, 1.0, 2.0, 3.0);
  cudaEventRecord(e0);
  k<W><<<sms, 128 * W>>>(out, iters, 1.0, 2.0, 3.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)sms * 128 * W * iters * 32;
  printf("warps/SMSP %d: %.3f ms, %.2f TFLOP/s, %.2f DFMA lanes / clk / SM at 1.9 GHz\n", W, ms, 2 * dfma / ms * 1e-9,
         dfma / (ms * 1e-3) / sms / 1.9e9);
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<1>(sms);
  run<2>(sms);
  run<4>(sms);
  return 0;
}
