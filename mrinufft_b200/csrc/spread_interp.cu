// spread_interp.cu -- kernels K2 (type-1 spreading) and K3 (type-2 interpolation),
// point-driven variants ("method 1"): one thread per sorted non-uniform point, loop over the
// batched transforms, exponential-of-semicircle weights evaluated once per point by Horner
// (float32 polynomial fitted at plan time).  These are the simple, always-correct variants that
// the tiled kernels (spread_tiled.cu) are checked against on the device.
//
// Replaces finufft's spread/interp stage reached through `Plan.execute_adjoint` / `Plan.execute`
// (src/mrinufft/operators/interfaces/finufft.py:69,76); algorithm per
// docs/explanations/nufft.rst:253-309.
#include "common.cuh"
#include "device_utils.cuh"

#define PD_THREADS 128

// weights for this thread's point: sw[(a * MAXW + i) * PD_THREADS + tid]
template <int DIM>
__global__ void __launch_bounds__(PD_THREADS)
k_spread_pd(Geom g, long long M, int T, const float* __restrict__ poly,
            const int32_t* __restrict__ perm, const int32_t* __restrict__ o0,
            const int32_t* __restrict__ o1, const int32_t* __restrict__ o2,
            const float* __restrict__ f0, const float* __restrict__ f1,
            const float* __restrict__ f2, const float2* __restrict__ ksp,
            const float* __restrict__ density, float2* __restrict__ fw) {
  extern __shared__ float smem[];
  float* spoly = smem;                               // (deg+1)*w
  float* sw = smem + (B200_MAX_DEG + 1) * B200_MAX_W;  // DIM * w * PD_THREADS
  const int w = g.w;
  for (int i = threadIdx.x; i < (g.deg + 1) * w; i += blockDim.x) spoly[i] = poly[i];
  __syncthreads();
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int tid = threadIdx.x;
  int org[3] = {0, 0, 0};
  {
    const int32_t* op[3] = {o0, o1, o2};
    const float* fp[3] = {f0, f1, f2};
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      org[a] = op[a][s];
      float z = fmaf(2.f, fp[a][s], (float)(w - 1));
      for (int i = 0; i < w; ++i) {
        float acc = spoly[g.deg * w + i];
        for (int k = g.deg - 1; k >= 0; --k) acc = fmaf(acc, z, spoly[k * w + i]);
        sw[(a * w + i) * PD_THREADS + tid] = acc;
      }
    }
  }
  const int j = perm[s];
  const float dens = density ? density[j] : 1.f;
  for (int t = 0; t < T; ++t) {
    float2 c = ksp[(long long)t * M + j];
    c.x *= dens;
    c.y *= dens;
    float2* fwt = fw + (long long)t * g.nftot;
    if (DIM == 1) {
      for (int i = 0; i < w; ++i) {
        int x = org[0] + i;
        if (x >= g.nf[0]) x -= g.nf[0];
        float wt = sw[i * PD_THREADS + tid];
        atomicAdd(&fwt[x], make_float2(c.x * wt, c.y * wt));
      }
    } else if (DIM == 2) {
      for (int iy = 0; iy < w; ++iy) {
        int y = org[0] + iy;
        if (y >= g.nf[0]) y -= g.nf[0];
        float wy = sw[iy * PD_THREADS + tid];
        float2 cy = make_float2(c.x * wy, c.y * wy);
        long long rowb = (long long)y * g.nf[1];
        for (int ix = 0; ix < w; ++ix) {
          int x = org[1] + ix;
          if (x >= g.nf[1]) x -= g.nf[1];
          float wx = sw[(w + ix) * PD_THREADS + tid];
          atomicAdd(&fwt[rowb + x], make_float2(cy.x * wx, cy.y * wx));
        }
      }
    } else {
      for (int iz = 0; iz < w; ++iz) {
        int z = org[0] + iz;
        if (z >= g.nf[0]) z -= g.nf[0];
        float wz = sw[iz * PD_THREADS + tid];
        for (int iy = 0; iy < w; ++iy) {
          int y = org[1] + iy;
          if (y >= g.nf[1]) y -= g.nf[1];
          float wzy = wz * sw[(w + iy) * PD_THREADS + tid];
          float2 cy = make_float2(c.x * wzy, c.y * wzy);
          long long rowb = ((long long)z * g.nf[1] + y) * g.nf[2];
          for (int ix = 0; ix < w; ++ix) {
            int x = org[2] + ix;
            if (x >= g.nf[2]) x -= g.nf[2];
            float wx = sw[(2 * w + ix) * PD_THREADS + tid];
            atomicAdd(&fwt[rowb + x], make_float2(cy.x * wx, cy.y * wx));
          }
        }
      }
    }
  }
}

// Interpolation; the epilogue optionally fuses the data-consistency residual (kernel K5):
//   out = scale * interp            (obs == nullptr)
//   out = (scale * interp - obs)    (obs != nullptr)   [density is applied by the spreader]
template <int DIM>
__global__ void __launch_bounds__(PD_THREADS)
k_interp_pd(Geom g, long long M, int T, const float* __restrict__ poly,
            const int32_t* __restrict__ perm, const int32_t* __restrict__ o0,
            const int32_t* __restrict__ o1, const int32_t* __restrict__ o2,
            const float* __restrict__ f0, const float* __restrict__ f1,
            const float* __restrict__ f2, const float2* __restrict__ fw,
            float2* __restrict__ ksp, float scale, const float2* __restrict__ obs) {
  extern __shared__ float smem[];
  float* spoly = smem;
  float* sw = smem + (B200_MAX_DEG + 1) * B200_MAX_W;
  const int w = g.w;
  for (int i = threadIdx.x; i < (g.deg + 1) * w; i += blockDim.x) spoly[i] = poly[i];
  __syncthreads();
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int tid = threadIdx.x;
  int org[3] = {0, 0, 0};
  {
    const int32_t* op[3] = {o0, o1, o2};
    const float* fp[3] = {f0, f1, f2};
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      org[a] = op[a][s];
      float z = fmaf(2.f, fp[a][s], (float)(w - 1));
      for (int i = 0; i < w; ++i) {
        float acc = spoly[g.deg * w + i];
        for (int k = g.deg - 1; k >= 0; --k) acc = fmaf(acc, z, spoly[k * w + i]);
        sw[(a * w + i) * PD_THREADS + tid] = acc;
      }
    }
  }
  const int j = perm[s];
  for (int t = 0; t < T; ++t) {
    const float2* fwt = fw + (long long)t * g.nftot;
    float2 acc = make_float2(0.f, 0.f);
    if (DIM == 1) {
      for (int i = 0; i < w; ++i) {
        int x = org[0] + i;
        if (x >= g.nf[0]) x -= g.nf[0];
        float wt = sw[i * PD_THREADS + tid];
        float2 v = __ldg(&fwt[x]);
        acc.x = fmaf(v.x, wt, acc.x);
        acc.y = fmaf(v.y, wt, acc.y);
      }
    } else if (DIM == 2) {
      for (int iy = 0; iy < w; ++iy) {
        int y = org[0] + iy;
        if (y >= g.nf[0]) y -= g.nf[0];
        float wy = sw[iy * PD_THREADS + tid];
        long long rowb = (long long)y * g.nf[1];
        float2 r = make_float2(0.f, 0.f);
        for (int ix = 0; ix < w; ++ix) {
          int x = org[1] + ix;
          if (x >= g.nf[1]) x -= g.nf[1];
          float wx = sw[(w + ix) * PD_THREADS + tid];
          float2 v = __ldg(&fwt[rowb + x]);
          r.x = fmaf(v.x, wx, r.x);
          r.y = fmaf(v.y, wx, r.y);
        }
        acc.x = fmaf(r.x, wy, acc.x);
        acc.y = fmaf(r.y, wy, acc.y);
      }
    } else {
      for (int iz = 0; iz < w; ++iz) {
        int z = org[0] + iz;
        if (z >= g.nf[0]) z -= g.nf[0];
        float wz = sw[iz * PD_THREADS + tid];
        float2 p = make_float2(0.f, 0.f);
        for (int iy = 0; iy < w; ++iy) {
          int y = org[1] + iy;
          if (y >= g.nf[1]) y -= g.nf[1];
          float wy = sw[(w + iy) * PD_THREADS + tid];
          long long rowb = ((long long)z * g.nf[1] + y) * g.nf[2];
          float2 r = make_float2(0.f, 0.f);
          for (int ix = 0; ix < w; ++ix) {
            int x = org[2] + ix;
            if (x >= g.nf[2]) x -= g.nf[2];
            float wx = sw[(2 * w + ix) * PD_THREADS + tid];
            float2 v = __ldg(&fwt[rowb + x]);
            r.x = fmaf(v.x, wx, r.x);
            r.y = fmaf(v.y, wx, r.y);
          }
          p.x = fmaf(r.x, wy, p.x);
          p.y = fmaf(r.y, wy, p.y);
        }
        acc.x = fmaf(p.x, wz, acc.x);
        acc.y = fmaf(p.y, wz, acc.y);
      }
    }
    acc.x *= scale;
    acc.y *= scale;
    const long long oi = (long long)t * M + j;
    if (obs) {
      float2 y = obs[oi];
      acc.x -= y.x;
      acc.y -= y.y;
    }
    ksp[oi] = acc;
  }
}

static size_t pd_smem(const Geom& g) {
  return ((B200_MAX_DEG + 1) * B200_MAX_W + (size_t)g.dim * g.w * PD_THREADS) * sizeof(float);
}

int spread_point_driven(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                        cudaStream_t st) {
  const long long M = p->M;
  if (M == 0) return B200_OK;
  const int nb = ceil_div(M, PD_THREADS);
  const size_t sm = pd_smem(p->g);
#define LAUNCH_SPREAD(D)                                                                        \
  k_spread_pd<D><<<nb, PD_THREADS, sm, st>>>(p->g, M, T, p->d_poly, p->d_perm, p->d_org_s[0],   \
                                             p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],        \
                                             p->d_x1_s[1], p->d_x1_s[2], ksp, density, fw)
  if (p->g.dim == 1) LAUNCH_SPREAD(1);
  else if (p->g.dim == 2) LAUNCH_SPREAD(2);
  else LAUNCH_SPREAD(3);
#undef LAUNCH_SPREAD
  CHECK_LAUNCH();
  return B200_OK;
}

int interp_point_driven(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                        const float2* obs, cudaStream_t st) {
  const long long M = p->M;
  if (M == 0) return B200_OK;
  const int nb = ceil_div(M, PD_THREADS);
  const size_t sm = pd_smem(p->g);
#define LAUNCH_INTERP(D)                                                                        \
  k_interp_pd<D><<<nb, PD_THREADS, sm, st>>>(p->g, M, T, p->d_poly, p->d_perm, p->d_org_s[0],   \
                                             p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],        \
                                             p->d_x1_s[1], p->d_x1_s[2], fw, ksp, scale, obs)
  if (p->g.dim == 1) LAUNCH_INTERP(1);
  else if (p->g.dim == 2) LAUNCH_INTERP(2);
  else LAUNCH_INTERP(3);
#undef LAUNCH_INTERP
  CHECK_LAUNCH();
  return B200_OK;
}
