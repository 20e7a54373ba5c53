// fftbench.cu -- cuFFT layout probe for the oversampled-grid FFT (design input, not product code).
// Times, for an n^3 complex64 grid: (a) contiguous batched 3-D C2C, (b) coil-interleaved layout
// (istride = T, idist = 1), (c) a zero-padding-aware split: 2-D (y,x) FFTs on half of the z planes
// + strided 1-D FFTs along z.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fftbench.cu -lcufft
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define FK(x) do { cufftResult r = (x); if (r != CUFFT_SUCCESS) { printf("CUFFT %d at %d\n", (int)r, __LINE__); exit(1);} } while (0)

static float time_exec(cufftHandle* plans, int nplans, cufftComplex** ptrs, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; ++w) for (int i = 0; i < nplans; ++i) FK(cufftExecC2C(plans[i], ptrs[i], ptrs[i], CUFFT_FORWARD));
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) for (int i = 0; i < nplans; ++i) FK(cufftExecC2C(plans[i], ptrs[i], ptrs[i], CUFFT_FORWARD));
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 512;
  int Tmax = argc > 2 ? atoi(argv[2]) : 8;
  size_t grid = (size_t)n * n * n;
  cufftComplex* d;
  CK(cudaMalloc(&d, grid * Tmax * sizeof(cufftComplex)));
  CK(cudaMemset(d, 0, grid * Tmax * sizeof(cufftComplex)));
  double gb = grid * 8.0 / 1e9;
  printf("n=%d grid=%.3f GB Tmax=%d\n", n, gb, Tmax);
  int dims[3] = {n, n, n};
  for (int T = 1; T <= Tmax; T *= 2) {
    cufftHandle p; size_t ws = 0;
    FK(cufftCreate(&p));
    FK(cufftMakePlanMany(p, 3, dims, NULL, 1, (int)grid, NULL, 1, (int)grid, CUFFT_C2C, T, &ws));
    cufftComplex* ptr = d;
    float ms = time_exec(&p, 1, &ptr, 5);
    printf("contiguous 3D  T=%2d: %8.3f ms  = %.3f ms/grid  (%.0f GB/s at 2 passes r+w) ws=%.2f GB\n", T, ms, ms / T,
           4 * gb * T / (ms * 1e-3), ws / 1e9);
    cufftDestroy(p);
  }
  for (int T = 2; T <= Tmax; T *= 4) {
    cufftHandle p; size_t ws = 0;
    FK(cufftCreate(&p));
    cufftResult r = cufftMakePlanMany(p, 3, dims, dims, T, 1, dims, T, 1, CUFFT_C2C, T, &ws);
    if (r != CUFFT_SUCCESS) { printf("interleaved T=%d: plan failed %d\n", T, (int)r); continue; }
    cufftComplex* ptr = d;
    float ms = time_exec(&p, 1, &ptr, 3);
    printf("interleaved 3D T=%2d: %8.3f ms  = %.3f ms/grid ws=%.2f GB\n", T, ms, ms / T, ws / 1e9);
    cufftDestroy(p);
  }
  {  // pruned: 2-D FFT over (y,x) on the two z slabs of n/4 planes + strided 1-D along z
    cufftHandle p2, pz; size_t ws = 0;
    int d2[2] = {n, n};
    FK(cufftCreate(&p2));
    FK(cufftMakePlanMany(p2, 2, d2, NULL, 1, n * n, NULL, 1, n * n, CUFFT_C2C, n / 4, &ws));
    FK(cufftCreate(&pz));
    int d1[1] = {n};
    int emb[1] = {n};
    FK(cufftMakePlanMany(pz, 1, d1, emb, n * n, 1, emb, n * n, 1, CUFFT_C2C, n * n, &ws));
    cufftHandle plans[3] = {p2, p2, pz};
    cufftComplex* ptrs[3] = {d, d + (size_t)(n - n / 4) * n * n, d};
    float ms = time_exec(plans, 3, ptrs, 5);
    cufftHandle pl2[2] = {p2, p2};
    float ms2 = time_exec(pl2, 2, ptrs, 5);
    cufftHandle plz[1] = {pz};
    cufftComplex* pz_ptr[1] = {d};
    float msz = time_exec(plz, 1, pz_ptr, 5);
    printf("pruned (2D on n/2 planes + strided z): %.3f ms/grid  [2D half %.3f, z pass %.3f]\n", ms, ms2, msz);
    // x-only contiguous pass and y-only strided pass for reference
    cufftHandle px; FK(cufftCreate(&px));
    FK(cufftMakePlanMany(px, 1, d1, NULL, 1, n, NULL, 1, n, CUFFT_C2C, n * n, &ws));
    cufftHandle plx[1] = {px};
    float msx = time_exec(plx, 1, pz_ptr, 5);
    printf("1D x pass (contiguous, n*n batch): %.3f ms (%.0f GB/s)\n", msx, 2 * gb / (msx * 1e-3));
    printf("1D z pass (stride n*n):            %.3f ms (%.0f GB/s)\n", msz, 2 * gb / (msz * 1e-3));
  }
  // plain device copy for the same box
  {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    if (Tmax >= 2) {
      CK(cudaMemcpy(d + grid, d, grid * 8, cudaMemcpyDeviceToDevice));
      cudaEventRecord(a);
      for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(d + grid, d, grid * 8, cudaMemcpyDeviceToDevice));
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("D2D copy of one grid: %.3f ms (%.0f GB/s r+w)\n", ms / 5, 2 * gb / (ms / 5 * 1e-3));
    }
  }
  return 0;
}
