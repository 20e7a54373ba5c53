// gridops.cu -- the HBM-bound passes around cuFFT.
//
// K4a  pad + deapodise (+ sensitivity-map multiply): image -> oversampled grid.  Replaces
//      finufft's deconvolve/zero-pad step inside `Plan.execute` (finufft.py:76) fused with the
//      `coil_img = smaps[idx] * x` pass of `_op_sense`
//      (src/mrinufft/operators/base.py:988-993; cufinufft.py:497).
// K4b  crop + deapodise (+ conj(smaps) multiply + coil accumulation + norm scale): oversampled
//      grid -> image.  Replaces finufft's deconvolve step of `execute_adjoint` (finufft.py:69)
//      fused with `_coil_combine_kernel` (src/mrinufft/operators/gpu_utils.py:12-24,
//      toeplitz.py:299-321) and `ret *= inv_norm_factor` (base.py:1034).
// Both are single passes: every oversampled-grid element is written (K4a) or the kept modes are
// read (K4b) exactly once.
#include "common.cuh"
#include "device_utils.cuh"

#define GO_THREADS 256

// One thread per pair of consecutive fine-grid elements along the fastest axis (16-byte stores).
__global__ void __launch_bounds__(GO_THREADS)
k_pad(Geom g, int T, const float2* __restrict__ img, const float2* __restrict__ smaps,
      const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ d2,
      float2* __restrict__ fw, int conj_smaps) {
  const int nfx = g.nf[g.dim - 1];
  const int Nx = g.N[g.dim - 1];
  const long long npairs = g.nftot / 2;  // nf is even along every axis
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  const int hx = nfx / 2;
  int lx = (int)(q % hx) * 2;
  long long rest = q / hx;
  // decode slow axes and their image indices
  long long nrow = 0;  // linear image row index over the slow axes
  float dsl = 1.f;
  bool inside = true;
  if (g.dim == 3) {
    int l1 = (int)(rest % g.nf[1]);
    int l0 = (int)(rest / g.nf[1]);
    int n0 = fine_to_mode(l0, g.N[0], g.nf[0]);
    int n1 = fine_to_mode(l1, g.N[1], g.nf[1]);
    inside = (n0 >= 0) && (n1 >= 0);
    if (inside) {
      nrow = (long long)n0 * g.N[1] + n1;
      dsl = d0[n0] * d1[n1];
    }
  } else if (g.dim == 2) {
    int l0 = (int)rest;
    int n0 = fine_to_mode(l0, g.N[0], g.nf[0]);
    inside = n0 >= 0;
    if (inside) {
      nrow = n0;
      dsl = d0[n0];
    }
  }
  const float* dx = (g.dim == 3) ? d2 : (g.dim == 2 ? d1 : d0);
  int nx0 = inside ? fine_to_mode(lx, Nx, nfx) : -1;
  int nx1 = inside ? fine_to_mode(lx + 1, Nx, nfx) : -1;
  const long long fidx = rest * nfx + lx;
  if (nx0 < 0 && nx1 < 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t)
      *reinterpret_cast<float4*>(&fw[(long long)t * g.nftot + fidx]) = z;
    return;
  }
  const float w0 = nx0 >= 0 ? dsl * dx[nx0] : 0.f;
  const float w1 = nx1 >= 0 ? dsl * dx[nx1] : 0.f;
  const long long i0 = nrow * Nx + (nx0 >= 0 ? nx0 : 0);
  const long long i1 = nrow * Nx + (nx1 >= 0 ? nx1 : 0);
  if (smaps) {
    float2 a0 = nx0 >= 0 ? cscale(img[i0], w0) : make_float2(0.f, 0.f);
    float2 a1 = nx1 >= 0 ? cscale(img[i1], w1) : make_float2(0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float2* sm = smaps + (long long)t * g.Ntot;
      float2 s0 = nx0 >= 0 ? sm[i0] : make_float2(0.f, 0.f);
      float2 s1 = nx1 >= 0 ? sm[i1] : make_float2(0.f, 0.f);
      float2 v0 = conj_smaps ? cmul_conj(a0, s0) : cmul(a0, s0);
      float2 v1 = conj_smaps ? cmul_conj(a1, s1) : cmul(a1, s1);
      *reinterpret_cast<float4*>(&fw[(long long)t * g.nftot + fidx]) =
          make_float4(v0.x, v0.y, v1.x, v1.y);
    }
  } else {
    for (int t = 0; t < T; ++t) {
      const float2* im = img + (long long)t * g.Ntot;
      float2 v0 = nx0 >= 0 ? cscale(im[i0], w0) : make_float2(0.f, 0.f);
      float2 v1 = nx1 >= 0 ? cscale(im[i1], w1) : make_float2(0.f, 0.f);
      *reinterpret_cast<float4*>(&fw[(long long)t * g.nftot + fidx]) =
          make_float4(v0.x, v0.y, v1.x, v1.y);
    }
  }
}

// One thread per image element.
__global__ void __launch_bounds__(GO_THREADS)
k_crop(Geom g, int T, const float2* __restrict__ fw, const float2* __restrict__ smaps,
       const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ d2,
       float2* __restrict__ img, int accumulate, float scale, int conj_smaps) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.Ntot) return;
  long long fidx;
  float dd = scale;
  if (g.dim == 3) {
    int n2 = (int)(n % g.N[2]);
    long long r = n / g.N[2];
    int n1 = (int)(r % g.N[1]);
    int n0 = (int)(r / g.N[1]);
    dd *= d0[n0] * d1[n1] * d2[n2];
    fidx = ((long long)mode_to_fine(n0, g.N[0], g.nf[0]) * g.nf[1] +
            mode_to_fine(n1, g.N[1], g.nf[1])) * g.nf[2] + mode_to_fine(n2, g.N[2], g.nf[2]);
  } else if (g.dim == 2) {
    int n1 = (int)(n % g.N[1]);
    int n0 = (int)(n / g.N[1]);
    dd *= d0[n0] * d1[n1];
    fidx = (long long)mode_to_fine(n0, g.N[0], g.nf[0]) * g.nf[1] +
           mode_to_fine(n1, g.N[1], g.nf[1]);
  } else {
    dd *= d0[n];
    fidx = mode_to_fine((int)n, g.N[0], g.nf[0]);
  }
  if (smaps) {
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      float2 v = fw[(long long)t * g.nftot + fidx];
      float2 s = smaps[(long long)t * g.Ntot + n];
      float2 pr = conj_smaps ? cmul(v, s) : cmul_conj(v, s);
      acc.x += pr.x;
      acc.y += pr.y;
    }
    acc = cscale(acc, dd);
    if (accumulate) {
      float2 o = img[n];
      acc.x += o.x;
      acc.y += o.y;
    }
    img[n] = acc;
  } else {
    for (int t = 0; t < T; ++t) {
      float2 v = cscale(fw[(long long)t * g.nftot + fidx], dd);
      const long long oi = (long long)t * g.Ntot + n;
      if (accumulate) {
        float2 o = img[oi];
        v.x += o.x;
        v.y += o.y;
      }
      img[oi] = v;
    }
  }
}

__global__ void __launch_bounds__(GO_THREADS)
k_pipe_div(long long M, float* __restrict__ d, const float2* __restrict__ ksp) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  float2 v = ksp[j];
  d[j] = d[j] / sqrtf(v.x * v.x + v.y * v.y);
}

__global__ void __launch_bounds__(GO_THREADS)
k_r2c(long long M, const float* __restrict__ d, float2* __restrict__ out) {
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  out[j] = make_float2(d[j], 0.f);
}

// fw[t][i] *= kern[i]: the Toeplitz spectrum multiply of the cuFFT path (the fused FFT passes apply
// it in the store of their last pass instead)
__global__ void __launch_bounds__(GO_THREADS)
k_mul_real_kernel(long long n, int T, float2* __restrict__ fw, const float* __restrict__ kern) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float k = kern[i];
  for (int t = 0; t < T; ++t) {
    float2 v = fw[(long long)t * n + i];
    v.x *= k;
    v.y *= k;
    fw[(long long)t * n + i] = v;
  }
}

int k_mul_real(b200_plan* p, float2* fw, const float* kern, int T, cudaStream_t st) {
  k_mul_real_kernel<<<ceil_div(p->g.nftot, GO_THREADS), GO_THREADS, 0, st>>>(p->g.nftot, T, fw, kern);
  CHECK_LAUNCH();
  return B200_OK;
}

int k4a_pad(b200_plan* p, const float2* img, const float2* smaps, float2* fw, int T,
            int conj_smaps, cudaStream_t st) {
  const long long npairs = p->g.nftot / 2;
  k_pad<<<ceil_div(npairs, GO_THREADS), GO_THREADS, 0, st>>>(
      p->g, T, img, smaps, p->dvec(0), p->dvec(1), p->dvec(2), fw, conj_smaps);
  CHECK_LAUNCH();
  return B200_OK;
}

int k4b_crop(b200_plan* p, const float2* fw, const float2* smaps, float2* img, int T,
             int accumulate, float scale, int conj_smaps, cudaStream_t st) {
  k_crop<<<ceil_div(p->g.Ntot, GO_THREADS), GO_THREADS, 0, st>>>(
      p->g, T, fw, smaps, p->dvec(0), p->dvec(1), p->dvec(2), img, accumulate, scale, conj_smaps);
  CHECK_LAUNCH();
  return B200_OK;
}

int k_pipe_update(b200_plan* p, float* d, const float2* ksp, cudaStream_t st) {
  if (p->M == 0) return B200_OK;
  k_pipe_div<<<ceil_div(p->M, GO_THREADS), GO_THREADS, 0, st>>>(p->M, d, ksp);
  CHECK_LAUNCH();
  return B200_OK;
}

int k_real_to_cpx(b200_plan* p, const float* d, float2* out, cudaStream_t st) {
  if (p->M == 0) return B200_OK;
  k_r2c<<<ceil_div(p->M, GO_THREADS), GO_THREADS, 0, st>>>(p->M, d, out);
  CHECK_LAUNCH();
  return B200_OK;
}
