// Issue rate of packed fma.rn.f32x2 (SASS FFMA2) against scalar FFMA on one B200 -- the denominator of the
// row kernels' FMA-pipe floor in DESIGN.md section 3.3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_rate tools/ffma2_rate.cu && /tmp/ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(512) k_rate(float* out, int iters, float seed, float mf, float cf) {
  // 16 independent chains per thread: latency (4 cycles) is covered even by a single warp per scheduler
  unsigned long long a[16];
  float s[16];
  for (int i = 0; i < 16; ++i) {
    s[i] = seed + i + threadIdx.x;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(s[i]));
  }
  unsigned long long m, c;
  // multiplier and addend come in as kernel arguments: register operands (the immediate form of FFMA
  // issues at twice the rate of the three-register form and would flatter the scalar figure)
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(mf));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(cf));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
      else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(mf), "f"(cf));
    }
  }
  float r = 0.f;
  for (int i = 0; i < 16; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    r += lo + hi + s[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int PACKED>
void run(const char* name, int sms, float clock_ghz_hint) {
  const int blocks = sms * 4, threads = 512, iters = 20000;
  float* d;
  cudaMalloc(&d, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_rate<PACKED><<<blocks, threads>>>(d, 100, 1.f, 1.0000001f, 1e-9f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_rate<PACKED><<<blocks, threads>>>(d, iters, 1.f, 1.0000001f, 1e-9f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_inst = (double)blocks * threads / 32 * iters * 16;
  const double per_s = warp_inst / (ms * 1e-3);
  printf("{\"op\": \"%s\", \"ms\": %.3f, \"warp_inst_per_s\": %.4g, \"warp_inst_per_clk_per_smsp_at_%.2fGHz\": %.3f, "
         "\"fp32_fma_per_s\": %.4g}\n",
         name, ms, per_s, clock_ghz_hint, per_s / (sms * 4.0 * clock_ghz_hint * 1e9), per_s * 32 * (PACKED ? 2 : 1));
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz * 1e-6f;
  run<0>("fma.rn.f32 (FFMA)", p.multiProcessorCount, ghz);
  run<1>("fma.rn.f32x2 (FFMA2)", p.multiProcessorCount, ghz);
  return 0;
}
