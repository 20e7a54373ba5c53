// setpts.cu -- kernel K1: fold the non-uniform points onto the oversampled grid, compute the
// footprint origin / first-tap offset, the bin key, and the stable sort by bin key.
//
// Replaces what finufft does inside `Plan.setpts` (call site
// src/mrinufft/operators/interfaces/finufft.py:55-62).  The bit-exact contract with the CPU
// reference sort is written in oracle/es_nufft.py::fold_points / bin_sort:
//
//   xd = double(x); t = xd * INV_2PI; t = t - floor(t + 0.5); g = (t + 0.5) * nf
//   g >= nf -> g - nf ;  g < 0 -> 0
//   i1 = ceil(g - w/2); x1 = float(i1 - g); origin = i1 mod nf
//
// every step one IEEE-754 double operation rounded to nearest (explicit __d*_rn intrinsics, so
// the compiler cannot contract them into FMAs).
//
// Bins are "pencils": one cell along every axis but the fastest, BX = 16 cells along the fastest
// axis, each split in two sub-bins by `cross` (does the footprint leave the 16-cell tile?).  The key
// orders them so that, for a fixed slow origin, x-tile and cross flag, consecutive middle origins
// are consecutive keys (3-D: key = ((o0 * nbx + bx) * 2 + cross) * nf1 + o1;
// 2-D: key = (bx * 2 + cross) * nf0 + o0; 1-D: key = bx).  The row kernels (spread_rows.cu) rely
// on that order: the points visiting a grid row are a handful of contiguous key ranges.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

__device__ __forceinline__ void fold_one(float x, int nf, int w, int* origin, float* x1) {
  const double INV_2PI = 0.15915494309189535;
  double xd = (double)x;
  double t = __dmul_rn(xd, INV_2PI);
  t = __dsub_rn(t, floor(__dadd_rn(t, 0.5)));
  double g = __dmul_rn(__dadd_rn(t, 0.5), (double)nf);
  if (g >= (double)nf) g = __dsub_rn(g, (double)nf);
  if (!(g >= 0.0)) g = 0.0;  // also catches NaN
  double i1 = ceil(__dsub_rn(g, 0.5 * (double)w));
  *x1 = (float)__dsub_rn(i1, g);
  int o = (int)i1;
  if (o < 0) o += nf;
  if (o >= nf) o -= nf;
  *origin = o;
}

// Bin key.  x-tiles of BX cells; `cross` = the footprint [ox, ox + w) leaves the point's own tile
// (the last tile may be shorter than BX).
__device__ __forceinline__ int make_key(const Geom& g, const int* o) {
  if (g.dim == 1) return o[0] / g.bin[0];
  const int ox = o[g.dim - 1];
  const int BX = g.bin[g.dim - 1];
  const int nfx = g.nf[g.dim - 1];
  const int bx = ox / BX;
  const int rest = nfx - bx * BX;
  const int tlen = rest < BX ? rest : BX;
  const int cross = (ox - bx * BX + g.w - 1 >= tlen) ? 1 : 0;
  if (g.dim == 3) return ((o[0] * g.nbins[2] + bx) * 2 + cross) * g.nf[1] + o[1];
  return (bx * 2 + cross) * g.nf[0] + o[0];
}

__global__ void __launch_bounds__(256)
k_fold(const float* __restrict__ xyz, long long M, Geom g, int32_t* __restrict__ o0,
       int32_t* __restrict__ o1, int32_t* __restrict__ o2, float* __restrict__ f0,
       float* __restrict__ f1, float* __restrict__ f2, int32_t* __restrict__ key,
       int32_t* __restrict__ iota) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int o[3] = {0, 0, 0};
  float f[3] = {0.f, 0.f, 0.f};
  for (int a = 0; a < g.dim; ++a) fold_one(xyz[j * g.dim + a], g.nf[a], g.w, &o[a], &f[a]);
  o0[j] = o[0];
  f0[j] = f[0];
  if (g.dim > 1) {
    o1[j] = o[1];
    f1[j] = f[1];
  }
  if (g.dim > 2) {
    o2[j] = o[2];
    f2[j] = f[2];
  }
  key[j] = make_key(g, o);
  iota[j] = (int32_t)j;
}

__global__ void __launch_bounds__(256)
k_gather_sorted(long long M, int dim, const int32_t* __restrict__ perm,
                const int32_t* __restrict__ o0, const int32_t* __restrict__ o1,
                const int32_t* __restrict__ o2, const float* __restrict__ f0,
                const float* __restrict__ f1, const float* __restrict__ f2,
                int32_t* __restrict__ so0, int32_t* __restrict__ so1, int32_t* __restrict__ so2,
                float* __restrict__ sf0, float* __restrict__ sf1, float* __restrict__ sf2) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  int j = perm[s];
  so0[s] = o0[j];
  sf0[s] = f0[j];
  if (dim > 1) {
    so1[s] = o1[j];
    sf1[s] = f1[j];
  }
  if (dim > 2) {
    so2[s] = o2[j];
    sf2[s] = f2[j];
  }
}

// bin_start[b] = first sorted position whose key is >= b  (bin_start[nbins] = M).
__global__ void __launch_bounds__(256)
k_bin_start(long long M, const int32_t* __restrict__ key_s, int32_t* __restrict__ bin_start,
            long long nbins) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > M) return;
  long long lo = (s == 0) ? 0 : (long long)key_s[s - 1] + 1;
  long long hi = (s == M) ? nbins : (long long)key_s[s];
  for (long long b = lo; b <= hi; ++b) bin_start[b] = (int32_t)s;
}

static int ensure_point_capacity(b200_plan* p, long long M) {
  if (M <= p->Mcap) return B200_OK;
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  for (int a = 0; a < 3; ++a) {
    fr(p->d_org_u[a]);
    fr(p->d_x1_u[a]);
    fr(p->d_org_s[a]);
    fr(p->d_x1_s[a]);
    p->d_org_u[a] = p->d_org_s[a] = nullptr;
    p->d_x1_u[a] = p->d_x1_s[a] = nullptr;
  }
  fr(p->d_key_u);
  fr(p->d_key_s);
  fr(p->d_perm);
  fr(p->d_iota);
  fr(p->d_sort_tmp);
  fr(p->d_ksp_tmp);
  fr(p->d_pipe_tmp);
  p->d_key_u = p->d_key_s = p->d_perm = p->d_iota = nullptr;
  p->d_sort_tmp = nullptr;
  p->d_ksp_tmp = nullptr;
  p->ksp_tmp_bytes = 0;
  p->d_pipe_tmp = nullptr;
  p->Mcap = 0;
  size_t n = (size_t)M;
  for (int a = 0; a < p->g.dim; ++a) {
    CUDA_TRY(cudaMalloc(&p->d_org_u[a], n * 4));
    CUDA_TRY(cudaMalloc(&p->d_x1_u[a], n * 4));
    CUDA_TRY(cudaMalloc(&p->d_org_s[a], n * 4));
    CUDA_TRY(cudaMalloc(&p->d_x1_s[a], n * 4));
  }
  CUDA_TRY(cudaMalloc(&p->d_key_u, n * 4));
  CUDA_TRY(cudaMalloc(&p->d_key_s, n * 4));
  CUDA_TRY(cudaMalloc(&p->d_perm, n * 4));
  CUDA_TRY(cudaMalloc(&p->d_iota, n * 4));
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)M);
  p->sort_tmp_bytes = tmp;
  CUDA_TRY(cudaMalloc(&p->d_sort_tmp, tmp));
  p->Mcap = M;
  return B200_OK;
}

int k1_setpts(b200_plan* p, const float* xyz, cudaStream_t st) {
  const long long M = p->M;
  B200_TRY(ensure_point_capacity(p, M));
  if (M == 0) {
    CUDA_TRY(cudaMemsetAsync(p->d_bin_start, 0, (size_t)(p->nbins_tot + 1) * 4, st));
    return B200_OK;
  }
  const int nb = ceil_div(M, 256);
  k_fold<<<nb, 256, 0, st>>>(xyz, M, p->g, p->d_org_u[0], p->d_org_u[1], p->d_org_u[2],
                             p->d_x1_u[0], p->d_x1_u[1], p->d_x1_u[2], p->d_key_u, p->d_iota);
  CHECK_LAUNCH();
  // number of significant key bits
  int bits = 1;
  while (bits < 31 && (1LL << bits) < p->nbins_tot) ++bits;
  size_t tmp = p->sort_tmp_bytes;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(p->d_sort_tmp, tmp, p->d_key_u, p->d_key_s, p->d_iota,
                                           p->d_perm, (int)M, 0, bits, st));
  g_kernel_launches += 3;  // CUB launches its histogram / onesweep kernels
  k_gather_sorted<<<nb, 256, 0, st>>>(M, p->g.dim, p->d_perm, p->d_org_u[0], p->d_org_u[1],
                                      p->d_org_u[2], p->d_x1_u[0], p->d_x1_u[1], p->d_x1_u[2],
                                      p->d_org_s[0], p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],
                                      p->d_x1_s[1], p->d_x1_s[2]);
  CHECK_LAUNCH();
  k_bin_start<<<ceil_div(M + 1, 256), 256, 0, st>>>(M, p->d_key_s, p->d_bin_start, p->nbins_tot);
  CHECK_LAUNCH();
  return B200_OK;
}

// ---------------------------------------------------------------- sort of points whose origins are known
// The complex128 path folds its points in double (double_path.cu) and hands the footprint origins over in
// p->d_org_u[]: bin key, stable sort, sorted origins and bin offsets as above (the float offsets are not
// used by that path).
__global__ void __launch_bounds__(256)
k_key_from_origins(long long M, Geom g, const int32_t* __restrict__ o0, const int32_t* __restrict__ o1,
                   const int32_t* __restrict__ o2, int32_t* __restrict__ key, int32_t* __restrict__ iota) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int o[3] = {o0[j], g.dim > 1 ? o1[j] : 0, g.dim > 2 ? o2[j] : 0};
  key[j] = make_key(g, o);
  iota[j] = (int32_t)j;
}

int k1_reserve_points(b200_plan* p, long long M) { return ensure_point_capacity(p, M); }

int k1_sort_origins(b200_plan* p, cudaStream_t st) {
  const long long M = p->M;
  if (M == 0) {
    CUDA_TRY(cudaMemsetAsync(p->d_bin_start, 0, (size_t)(p->nbins_tot + 1) * 4, st));
    return B200_OK;
  }
  const int nb = ceil_div(M, 256);
  k_key_from_origins<<<nb, 256, 0, st>>>(M, p->g, p->d_org_u[0], p->d_org_u[1], p->d_org_u[2], p->d_key_u,
                                         p->d_iota);
  CHECK_LAUNCH();
  int bits = 1;
  while (bits < 31 && (1LL << bits) < p->nbins_tot) ++bits;
  size_t tmp = p->sort_tmp_bytes;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(p->d_sort_tmp, tmp, p->d_key_u, p->d_key_s, p->d_iota, p->d_perm,
                                           (int)M, 0, bits, st));
  g_kernel_launches += 3;
  k_gather_sorted<<<nb, 256, 0, st>>>(M, p->g.dim, p->d_perm, p->d_org_u[0], p->d_org_u[1], p->d_org_u[2],
                                      p->d_x1_u[0], p->d_x1_u[1], p->d_x1_u[2], p->d_org_s[0], p->d_org_s[1],
                                      p->d_org_s[2], p->d_x1_s[0], p->d_x1_s[1], p->d_x1_s[2]);
  CHECK_LAUNCH();
  k_bin_start<<<ceil_div(M + 1, 256), 256, 0, st>>>(M, p->d_key_s, p->d_bin_start, p->nbins_tot);
  CHECK_LAUNCH();
  return B200_OK;
}
