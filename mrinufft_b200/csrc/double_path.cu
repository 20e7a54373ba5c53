// double_path.cu -- complex128 / float64 instantiation of the path (plans created with B200_DOUBLE).
//
// The reference computes in the precision of the sample array (`dtype = samples.dtype`,
// src/mrinufft/operators/base.py:934), i.e. finufft runs in double for float64 trajectories; its GPU
// peer casts to single (cufinufft.py:338-340).  The b200 backend is single-precision native -- every
// tuned kernel (row kernels, fused FFT passes) is float -- and offers double as an opt-in,
// correctness-first path: the same algorithm (fold, exponential-of-semicircle kernel of width
// w = ceil(log10(10 / eps)) <= 16, sigma = 2, deapodisation by quadrature) with the kernel evaluated
// directly (exp / sqrt in double, no polynomial), tile-owned spreading (double_rows.cu; point-driven
// with native double atomics for the geometries that file does not take, and in 1-D), point-driven
// interpolation, cuFFT Z2Z and separate pad / crop passes.  Same C entry points, data
// pointers are complex128 / float64.  Not supported in double: B200_SPREAD_ONLY plans, the sort
// read-back and the Toeplitz entry point.
#include <cmath>

#include <cstdlib>

#include "common.cuh"

namespace {

struct DblState {
  int32_t* d_org[3] = {nullptr, nullptr, nullptr};
  double* d_x1[3] = {nullptr, nullptr, nullptr};
  double* d_deapod[3] = {nullptr, nullptr, nullptr};
  double2* d_fw = nullptr;
  double2* d_res = nullptr;  // k-space residual of data_consistency
  size_t res_cap = 0;
  long long Mcap = 0;
  bool rows = false;  // the current points have a visit stream (double_rows.cu)
  cufftHandle fft = 0;
  bool fft_ok = false;
};

DblState* dstate(b200_plan* p) { return (DblState*)p->dbl; }

__device__ __forceinline__ double es_phi_d(double x, double hw, double beta) {
  const double r = x / hw;
  const double a = 1.0 - r * r;
  return a < 0.0 ? 0.0 : exp(beta * (sqrt(a) - 1.0));
}

// same fold as setpts.cu (one IEEE double operation per step)
__global__ void __launch_bounds__(256)
kd_fold(const double* __restrict__ xyz, long long M, Geom g, int32_t* __restrict__ o0,
        int32_t* __restrict__ o1, int32_t* __restrict__ o2, double* __restrict__ f0,
        double* __restrict__ f1, double* __restrict__ f2) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int32_t* op[3] = {o0, o1, o2};
  double* fp[3] = {f0, f1, f2};
  const double INV_2PI = 0.15915494309189535;
  for (int a = 0; a < g.dim; ++a) {
    const int nf = g.nf[a];
    double t = __dmul_rn(xyz[j * g.dim + a], INV_2PI);
    t = __dsub_rn(t, floor(__dadd_rn(t, 0.5)));
    double gg = __dmul_rn(__dadd_rn(t, 0.5), (double)nf);
    if (gg >= (double)nf) gg = __dsub_rn(gg, (double)nf);
    if (!(gg >= 0.0)) gg = 0.0;
    const double i1 = ceil(__dsub_rn(gg, 0.5 * (double)g.w));
    int o = (int)i1;
    if (o < 0) o += nf;
    if (o >= nf) o -= nf;
    op[a][j] = o;
    fp[a][j] = __dsub_rn(i1, gg);
  }
}

template <int DIM>
__device__ __forceinline__ void point_weights(const Geom& g, double beta, long long j,
                                              const int32_t* const* org, const double* const* x1,
                                              int (&o)[3], double (&wt)[3][B200_MAX_W]) {
  const double hw = 0.5 * g.w;
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    o[a] = org[a][j];
    const double x = x1[a][j];
    for (int i = 0; i < g.w; ++i) wt[a][i] = es_phi_d(x + i, hw, beta);
  }
}

struct PtArgs {
  const int32_t* org[3];
  const double* x1[3];
};

template <int DIM>
__global__ void __launch_bounds__(128)
kd_spread(Geom g, double beta, long long M, int T, PtArgs P, const double2* __restrict__ ksp,
          const double* __restrict__ density, double2* __restrict__ fw) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int o[3] = {0, 0, 0};
  double wt[3][B200_MAX_W];
  point_weights<DIM>(g, beta, j, P.org, P.x1, o, wt);
  const double dens = density ? density[j] : 1.0;
  const int w = g.w;
  for (int t = 0; t < T; ++t) {
    double2 c = ksp[(long long)t * M + j];
    c.x *= dens;
    c.y *= dens;
    double* fwt = reinterpret_cast<double*>(fw + (long long)t * g.nftot);
    for (int i0 = 0; i0 < w; ++i0) {
      int a0 = o[0] + i0;
      if (a0 >= g.nf[0]) a0 -= g.nf[0];
      const double w0 = wt[0][i0];
      if (DIM == 1) {
        atomicAdd(fwt + 2 * (long long)a0, c.x * w0);
        atomicAdd(fwt + 2 * (long long)a0 + 1, c.y * w0);
        continue;
      }
      for (int i1 = 0; i1 < w; ++i1) {
        int a1 = o[1] + i1;
        if (a1 >= g.nf[1]) a1 -= g.nf[1];
        const double w01 = w0 * wt[1][i1];
        const long long r1 = (long long)a0 * g.nf[1] + a1;
        if (DIM == 2) {
          atomicAdd(fwt + 2 * r1, c.x * w01);
          atomicAdd(fwt + 2 * r1 + 1, c.y * w01);
          continue;
        }
        for (int i2 = 0; i2 < w; ++i2) {
          int a2 = o[2] + i2;
          if (a2 >= g.nf[2]) a2 -= g.nf[2];
          const double ww = w01 * wt[2][i2];
          const long long r2 = r1 * g.nf[2] + a2;
          atomicAdd(fwt + 2 * r2, c.x * ww);
          atomicAdd(fwt + 2 * r2 + 1, c.y * ww);
        }
      }
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(128)
kd_interp(Geom g, double beta, long long M, int T, PtArgs P, const double2* __restrict__ fw,
          double2* __restrict__ ksp, double scale, const double2* __restrict__ obs,
          const int32_t* __restrict__ perm) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  // with a bin sort at hand (double_rows.cu) the threads of a warp take neighbouring points: their loads
  // fall on the same few lines
  const long long j = perm ? perm[s] : s;
  int o[3] = {0, 0, 0};
  double wt[3][B200_MAX_W];
  point_weights<DIM>(g, beta, j, P.org, P.x1, o, wt);
  const int w = g.w;
  for (int t = 0; t < T; ++t) {
    const double2* fwt = fw + (long long)t * g.nftot;
    double ax = 0.0, ay = 0.0;
    for (int i0 = 0; i0 < w; ++i0) {
      int a0 = o[0] + i0;
      if (a0 >= g.nf[0]) a0 -= g.nf[0];
      const double w0 = wt[0][i0];
      if (DIM == 1) {
        const double2 v = fwt[a0];
        ax += v.x * w0;
        ay += v.y * w0;
        continue;
      }
      for (int i1 = 0; i1 < w; ++i1) {
        int a1 = o[1] + i1;
        if (a1 >= g.nf[1]) a1 -= g.nf[1];
        const double w01 = w0 * wt[1][i1];
        const long long r1 = (long long)a0 * g.nf[1] + a1;
        if (DIM == 2) {
          const double2 v = fwt[r1];
          ax += v.x * w01;
          ay += v.y * w01;
          continue;
        }
        for (int i2 = 0; i2 < w; ++i2) {
          int a2 = o[2] + i2;
          if (a2 >= g.nf[2]) a2 -= g.nf[2];
          const double2 v = fwt[r1 * g.nf[2] + a2];
          const double ww = w01 * wt[2][i2];
          ax += v.x * ww;
          ay += v.y * ww;
        }
      }
    }
    double2 r = make_double2(ax * scale, ay * scale);
    const long long oi = (long long)t * M + j;
    if (obs) {
      r.x -= obs[oi].x;
      r.y -= obs[oi].y;
    }
    ksp[oi] = r;
  }
}

__device__ __forceinline__ int mode_to_fine_d(int n, int N, int nf) {
  const int k = n - N / 2;
  return k < 0 ? k + nf : k;
}

// image element (t, n) -> its place in the (pre-zeroed) oversampled grid
__global__ void __launch_bounds__(256)
kd_pad(Geom g, int T, const double2* __restrict__ img, const double2* __restrict__ smaps,
       const double* __restrict__ d0, const double* __restrict__ d1, const double* __restrict__ d2,
       double2* __restrict__ fw, int conj_smaps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.Ntot) return;
  int n[3] = {0, 0, 0};
  long long r = i;
  for (int a = g.dim - 1; a >= 0; --a) {
    n[a] = (int)(r % g.N[a]);
    r /= g.N[a];
  }
  double d = d0[n[0]];
  long long fi = mode_to_fine_d(n[0], g.N[0], g.nf[0]);
  if (g.dim > 1) {
    d *= d1[n[1]];
    fi = fi * g.nf[1] + mode_to_fine_d(n[1], g.N[1], g.nf[1]);
  }
  if (g.dim > 2) {
    d *= d2[n[2]];
    fi = fi * g.nf[2] + mode_to_fine_d(n[2], g.N[2], g.nf[2]);
  }
  for (int t = 0; t < T; ++t) {
    double2 v = smaps ? img[i] : img[(long long)t * g.Ntot + i];
    v.x *= d;
    v.y *= d;
    if (smaps) {
      double2 s = smaps[(long long)t * g.Ntot + i];
      if (conj_smaps) s.y = -s.y;
      v = make_double2(v.x * s.x - v.y * s.y, v.x * s.y + v.y * s.x);
    }
    fw[(long long)t * g.nftot + fi] = v;
  }
}

__global__ void __launch_bounds__(256)
kd_crop(Geom g, int T, const double2* __restrict__ fw, const double2* __restrict__ smaps,
        const double* __restrict__ d0, const double* __restrict__ d1, const double* __restrict__ d2,
        double2* __restrict__ img, int accumulate, double scale, int conj_smaps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.Ntot) return;
  int n[3] = {0, 0, 0};
  long long r = i;
  for (int a = g.dim - 1; a >= 0; --a) {
    n[a] = (int)(r % g.N[a]);
    r /= g.N[a];
  }
  double d = d0[n[0]] * scale;
  long long fi = mode_to_fine_d(n[0], g.N[0], g.nf[0]);
  if (g.dim > 1) {
    d *= d1[n[1]];
    fi = fi * g.nf[1] + mode_to_fine_d(n[1], g.N[1], g.nf[1]);
  }
  if (g.dim > 2) {
    d *= d2[n[2]];
    fi = fi * g.nf[2] + mode_to_fine_d(n[2], g.N[2], g.nf[2]);
  }
  double2 acc = make_double2(0.0, 0.0);
  for (int t = 0; t < T; ++t) {
    double2 v = fw[(long long)t * g.nftot + fi];
    if (smaps) {
      double2 s = smaps[(long long)t * g.Ntot + i];
      if (!conj_smaps) s.y = -s.y;  // adjoint: multiply by conj(smaps)
      acc.x += v.x * s.x - v.y * s.y;
      acc.y += v.x * s.y + v.y * s.x;
    } else {
      double2 o = make_double2(v.x * d, v.y * d);
      double2* dst = img + (long long)t * g.Ntot + i;
      if (accumulate) {
        o.x += dst->x;
        o.y += dst->y;
      }
      *dst = o;
    }
  }
  if (smaps) {
    double2 o = make_double2(acc.x * d, acc.y * d);
    if (accumulate) {
      o.x += img[i].x;
      o.y += img[i].y;
    }
    img[i] = o;
  }
}

// (-1)^k / phihat(k), double precision (cf. es_deapod_vector in es_kernel_host.cpp), Gauss-Legendre
// on theta with x = (w/2) sin(theta)
void deapod_d(int n, int nf, int w, double beta, std::vector<double>* out) {
  const int nq = 200;
  std::vector<double> x(nq), wt(nq);
  for (int i = 0; i < nq; ++i) {  // Gauss-Legendre nodes on [-1, 1] by Newton iteration
    double z = cos(M_PI * (i + 0.75) / (nq + 0.5)), pp = 0, p1 = 0;
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0;
      p1 = z;
      for (int k = 2; k <= nq; ++k) {
        const double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      pp = nq * (z * p1 - p0) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    double p0 = 1.0;
    p1 = z;
    for (int k = 2; k <= nq; ++k) {
      const double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
      p0 = p1;
      p1 = p2;
    }
    pp = nq * (z * p1 - p0) / (z * z - 1.0);
    x[i] = z;
    wt[i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  const double h = 0.5 * w;
  std::vector<double> f(nq), sx(nq);
  for (int q = 0; q < nq; ++q) {
    const double th = (x[q] + 1.0) * (M_PI / 4.0);
    f[q] = exp(beta * (cos(th) - 1.0)) * cos(th) * h * wt[q] * (M_PI / 4.0);
    sx[q] = h * sin(th);
  }
  out->resize(n);
  for (int i = 0; i < n; ++i) {
    const int k = i - n / 2;
    double s = 0;
    for (int q = 0; q < nq; ++q) s += f[q] * cos(2.0 * M_PI * k * sx[q] / nf);
    (*out)[i] = ((k % 2 == 0) ? 1.0 : -1.0) / (2.0 * s);
  }
}

PtArgs pt_args(DblState* ds) {
  PtArgs P;
  for (int a = 0; a < 3; ++a) {
    P.org[a] = ds->d_org[a];
    P.x1[a] = ds->d_x1[a];
  }
  return P;
}

int fft_d(b200_plan* p, DblState* ds, int T, int sign, cudaStream_t st) {
  if (p->fft_method == 4) return fft_any_c128(ds->d_fw, T, p->g, sign, st);  // own any-length passes (A/B, see api.cu)
  CUFFT_TRY(cufftSetStream(ds->fft, st));
  for (int t = 0; t < T; ++t) {
    cufftDoubleComplex* q = (cufftDoubleComplex*)(ds->d_fw + (long long)t * p->g.nftot);
    CUFFT_TRY(cufftExecZ2Z(ds->fft, q, q, sign < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    ++g_fft_execs;
  }
  return B200_OK;
}

int type2_d(b200_plan* p, const double2* img, const double2* smaps, double2* ksp, int T, int isign,
            double scale, int conj_smaps, const double2* obs, cudaStream_t st) {
  DblState* ds = dstate(p);
  const Geom& g = p->g;
  CUDA_TRY(cudaMemsetAsync(ds->d_fw, 0, (size_t)T * g.nftot * sizeof(double2), st));
  kd_pad<<<ceil_div(g.Ntot, 256), 256, 0, st>>>(g, T, img, smaps, ds->d_deapod[0], ds->d_deapod[1],
                                                ds->d_deapod[2], ds->d_fw, conj_smaps);
  CHECK_LAUNCH();
  B200_TRY(fft_d(p, ds, T, isign, st));
  if (p->M == 0) return B200_OK;
  const PtArgs P = pt_args(ds);
  const int nb = ceil_div(p->M, 128);
  const int32_t* perm = (ds->rows && !(p->rows_dbg & 64)) ? p->d_perm : nullptr;
  if (g.dim == 1) kd_interp<1><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ds->d_fw, ksp, scale, obs, perm);
  else if (g.dim == 2) kd_interp<2><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ds->d_fw, ksp, scale, obs, perm);
  else kd_interp<3><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ds->d_fw, ksp, scale, obs, perm);
  CHECK_LAUNCH();
  return B200_OK;
}

int type1_d(b200_plan* p, const double2* ksp, const double* density, const double2* smaps, double2* img,
            int T, int accumulate, int isign, double scale, int conj_smaps, cudaStream_t st) {
  DblState* ds = dstate(p);
  const Geom& g = p->g;
  // the row kernels write every cell of the grid themselves
  bool rows = false;
  if (ds->rows && p->M > 0) {
    const int rc = drows_spread(p, ksp, density, ds->d_fw, T, st);
    if (rc < 0) return rc;
    rows = rc == B200_OK;
  }
  if (!rows) CUDA_TRY(cudaMemsetAsync(ds->d_fw, 0, (size_t)T * g.nftot * sizeof(double2), st));
  if (p->M > 0 && !rows) {
    const PtArgs P = pt_args(ds);
    const int nb = ceil_div(p->M, 128);
    if (g.dim == 1) kd_spread<1><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ksp, density, ds->d_fw);
    else if (g.dim == 2) kd_spread<2><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ksp, density, ds->d_fw);
    else kd_spread<3><<<nb, 128, 0, st>>>(g, p->beta, p->M, T, P, ksp, density, ds->d_fw);
    CHECK_LAUNCH();
  }
  B200_TRY(fft_d(p, ds, T, isign, st));
  kd_crop<<<ceil_div(g.Ntot, 256), 256, 0, st>>>(g, T, ds->d_fw, smaps, ds->d_deapod[0], ds->d_deapod[1],
                                                 ds->d_deapod[2], img, accumulate, scale, conj_smaps);
  CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace

// ------------------------------------------------------------------ called from api.cu
int dbl_init(b200_plan* p) {
  DblState* ds = new DblState();
  p->dbl = ds;
  const Geom& g = p->g;
  for (int a = 0; a < g.dim; ++a) {
    std::vector<double> dv;
    deapod_d(g.N[a], g.nf[a], g.w, p->beta, &dv);
    CUDA_TRY(cudaMalloc(&ds->d_deapod[a], dv.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(ds->d_deapod[a], dv.data(), dv.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (g.dim >= 2) {
    CUDA_TRY(cudaMalloc(&p->d_bin_start, (size_t)(p->nbins_tot + 1) * sizeof(int32_t)));
    p->ws_bytes += (size_t)(p->nbins_tot + 1) * sizeof(int32_t);
  }
  const size_t fwb = (size_t)p->ntrans_max * g.nftot * sizeof(double2);
  if (cudaMalloc(&ds->d_fw, fwb) != cudaSuccess) {
    cudaGetLastError();
    b200_set_error("double path: cannot allocate %zu bytes of workspace", fwb);
    return B200_ENOMEM;
  }
  p->ws_bytes += fwb;
  long long n[3] = {g.nf[0], g.nf[1], g.nf[2]};
  size_t wsz = 0;
  cufftResult r = cufftCreate(&ds->fft);
  if (r == CUFFT_SUCCESS)
    r = cufftMakePlanMany64(ds->fft, g.dim, n, nullptr, 1, g.nftot, nullptr, 1, g.nftot, CUFFT_Z2Z, 1, &wsz);
  if (r != CUFFT_SUCCESS) {
    b200_set_error("cufftMakePlanMany64(Z2Z) failed: %d", (int)r);
    return B200_ECUFFT;
  }
  ds->fft_ok = true;
  p->ws_bytes += wsz;
  return B200_OK;
}

void dbl_free(b200_plan* p) {
  drows_free(p);
  DblState* ds = dstate(p);
  if (!ds) return;
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  for (int a = 0; a < 3; ++a) {
    fr(ds->d_org[a]);
    fr(ds->d_x1[a]);
    fr(ds->d_deapod[a]);
  }
  fr(ds->d_fw);
  fr(ds->d_res);
  if (ds->fft_ok) cufftDestroy(ds->fft);
  delete ds;
  p->dbl = nullptr;
}

int dbl_setpts(b200_plan* p, const double* xyz, cudaStream_t st) {
  DblState* ds = dstate(p);
  const long long M = p->M;
  if (M > ds->Mcap) {
    for (int a = 0; a < 3; ++a) {
      if (ds->d_org[a]) cudaFree(ds->d_org[a]);
      if (ds->d_x1[a]) cudaFree(ds->d_x1[a]);
      ds->d_org[a] = nullptr;
      ds->d_x1[a] = nullptr;
    }
    for (int a = 0; a < p->g.dim; ++a) {
      CUDA_TRY(cudaMalloc(&ds->d_org[a], (size_t)M * 4));
      CUDA_TRY(cudaMalloc(&ds->d_x1[a], (size_t)M * 8));
    }
    ds->Mcap = M;
  }
  ds->rows = false;
  if (M == 0) return B200_OK;
  kd_fold<<<ceil_div(M, 256), 256, 0, st>>>(xyz, M, p->g, ds->d_org[0], ds->d_org[1], ds->d_org[2],
                                           ds->d_x1[0], ds->d_x1[1], ds->d_x1[2]);
  CHECK_LAUNCH();
  if (drows_supported(p)) {
    // bin sort on the origins + visit stream of the tile-owned spreader
    B200_TRY(k1_reserve_points(p, M));
    for (int a = 0; a < p->g.dim; ++a)
      CUDA_TRY(cudaMemcpyAsync(p->d_org_u[a], ds->d_org[a], (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
    const int rc = drows_setpts(p, ds->d_x1, st);
    if (rc < 0) return rc;  // rc > 0: this trajectory is left to the point-driven kernel
    ds->rows = rc == B200_OK;
  }
  return B200_OK;
}

int dbl_type2(b200_plan* p, const void* img, const void* smaps, void* ksp, int T, int isign, double scale,
              int conj_smaps, cudaStream_t st) {
  return type2_d(p, (const double2*)img, (const double2*)smaps, (double2*)ksp, T, isign, scale, conj_smaps,
                 nullptr, st);
}

int dbl_type1(b200_plan* p, const void* ksp, const void* density, const void* smaps, void* img, int T,
              int accumulate, int isign, double scale, int conj_smaps, cudaStream_t st) {
  return type1_d(p, (const double2*)ksp, (const double*)density, (const double2*)smaps, (double2*)img, T,
                 accumulate, isign, scale, conj_smaps, st);
}

int dbl_data_consistency(b200_plan* p, const void* img, const void* smaps, const void* obs,
                         const void* density, void* grad, int T, int accumulate, double scale,
                         cudaStream_t st) {
  DblState* ds = dstate(p);
  const size_t need = (size_t)p->ntrans_max * (size_t)(p->M > 0 ? p->M : 1) * sizeof(double2);
  if (need > ds->res_cap) {
    if (ds->d_res) cudaFree(ds->d_res);
    ds->d_res = nullptr;
    CUDA_TRY(cudaMalloc(&ds->d_res, need));
    ds->res_cap = need;
  }
  B200_TRY(type2_d(p, (const double2*)img, (const double2*)smaps, ds->d_res, T, -1, scale, 0,
                   (const double2*)obs, st));
  return type1_d(p, ds->d_res, (const double*)density, (const double2*)smaps, (double2*)grad, T, accumulate,
                 +1, scale, 0, st);
}
