// api.cu -- the C ABI of libb200nufft.so (declared in include/b200nufft.h).
#include <cmath>
#include <cstring>

#include <cstdlib>

#include "common.cuh"

long long g_kernel_launches = 0;
long long g_fft_execs = 0;

int spread_point_driven(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                        cudaStream_t st);
int interp_point_driven(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                        const float2* obs, cudaStream_t st);
int spread_tiled(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                 cudaStream_t st);
int interp_tiled(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                 const float2* obs, const uint32_t* unread, cudaStream_t st);
void tiled_class_info(b200_plan* p, int T, int64_t out[4]);
bool tiled_supported(const b200_plan* p, int T);
void tiled_free(b200_plan* p);
void tiled_invalidate(b200_plan* p);
const uint32_t* tiled_empty_bits(b200_plan* p, int T, cudaStream_t st);

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

enum { EV_SPREAD = 0, EV_INTERP = 1, EV_FFT = 2, EV_GRID = 3 };

struct Timed {
  b200_plan* p;
  int slot;
  cudaStream_t st;
  Timed(b200_plan* p_, int slot_, cudaStream_t st_) : p(p_), slot(slot_), st(st_) {
    if (p->timing && p->ev_ok) cudaEventRecord(p->ev[2 * slot], st);
  }
  ~Timed() {
    if (p->timing && p->ev_ok) {
      cudaEventRecord(p->ev[2 * slot + 1], st);
      p->ev_used[slot] = 1;
    }
  }
};

int exec_fft(b200_plan* p, float2* fw, int T, int sign, cudaStream_t st) {
  Timed tm(p, EV_FFT, st);
  // option 2 = 4: the library's own any-length passes (fft_any.cu) instead of cuFFT -- correct for every grid,
  // measured at 1.5 - 1.8x cuFFT's time on grids with factors 3 / 5, so cuFFT stays the default here
  if (p->fft_method == 4) return fft_any_c64(fw, T, p->g, sign, st);
  CUFFT_TRY(cufftSetStream(p->fft, st));
  const int dir = sign < 0 ? CUFFT_FORWARD : CUFFT_INVERSE;
  int done = 0;
  // the cuFFT plan is batched over fft_batch grids; run ceil(T / fft_batch) executions
  while (done < T) {
    // a batched plan always transforms fft_batch grids; the workspace holds ntrans_max >=
    // fft_batch grids so transforming a few stale ones at the tail is harmless only if they
    // exist; use the single-grid plan for the remainder instead.
    if (T - done >= p->fft_batch) {
      CUFFT_TRY(cufftExecC2C(p->fft, (cufftComplex*)(fw + (long long)done * p->g.nftot),
                             (cufftComplex*)(fw + (long long)done * p->g.nftot), dir));
      done += p->fft_batch;
    } else {
      CUFFT_TRY(cufftSetStream(p->fft1, st));
      CUFFT_TRY(cufftExecC2C(p->fft1, (cufftComplex*)(fw + (long long)done * p->g.nftot),
                             (cufftComplex*)(fw + (long long)done * p->g.nftot), dir));
      done += 1;
    }
    ++g_fft_execs;
  }
  return B200_OK;
}

int do_spread(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
              cudaStream_t st) {
  Timed tm(p, EV_SPREAD, st);
  int method = p->spread_method;
  p->spread_empty = nullptr;
  if (method == 0) method = tiled_supported(p, T) ? 2 : 1;
  if (method == 2 && !tiled_supported(p, T)) method = 1;
  if (method == 2) {
    const int rc = spread_tiled(p, ksp, density, fw, T, st);
    if (rc != 1) return rc;  // 1: the row kernels cannot serve this trajectory
  }
  CUDA_TRY(cudaMemsetAsync(fw, 0, (size_t)T * p->g.nftot * sizeof(float2), st));
  return spread_point_driven(p, ksp, density, fw, T, st);
}

int do_interp(b200_plan* p, const float2* fw, float2* ksp, int T, float scale, const float2* obs,
              cudaStream_t st) {
  Timed tm(p, EV_INTERP, st);
  int method = p->interp_method;
  if (method == 0) method = tiled_supported(p, T) ? 2 : 1;
  if (method == 2 && !tiled_supported(p, T)) method = 1;
  const uint32_t* unread = p->interp_unread;
  p->interp_unread = nullptr;
  if (method == 2) {
    const int rc = interp_tiled(p, fw, ksp, T, scale, obs, unread, st);
    if (rc != 1) return rc;
  }
  if (unread) {
    b200_set_error("internal: the grid was produced for the row interpolator, which cannot serve this call");
    return B200_ESTATE;
  }
  return interp_point_driven(p, fw, ksp, T, scale, obs, st);
}

bool use_fftp(const b200_plan* p) {
  return p->fft_method != 1 && p->fft_method != 4 && fftp_supported(p);  // (2: forced, 3: with bulk tensor loads)
}

// image(s) -> transformed oversampled grid (K4a + FFT)
// `for_tiled_interp`: the grid is consumed by the row interpolator next, which only reads tiles that
// some point visits -- the last FFT pass may leave the others unwritten
int image_to_grid(b200_plan* p, const float2* img, const float2* smaps, int T, int isign,
                  int conj_smaps, cudaStream_t st, bool for_interp = false) {
  if (use_fftp(p)) {
    const uint32_t* unread = nullptr;
    if (for_interp && p->pts_set && p->M > 0 && p->g.dim == 3) {
      int method = p->interp_method;
      if (method == 0) method = tiled_supported(p, T) ? 2 : 1;
      if (method == 2 && tiled_supported(p, T)) unread = tiled_empty_bits(p, T, st);
    }
    p->interp_unread = unread;
    Timed tm(p, EV_FFT, st);
    return fftp_type2(p, img, smaps, p->d_fw, T, isign, conj_smaps, st, nullptr, unread);
  }
  {
    Timed tm(p, EV_GRID, st);
    B200_TRY(k4a_pad(p, img, smaps, p->d_fw, T, conj_smaps, st));
  }
  return exec_fft(p, p->d_fw, T, isign, st);
}

// spread oversampled grid -> image(s) (FFT + K4b)
int grid_to_image(b200_plan* p, const float2* smaps, float2* img, int T, int accumulate, int isign,
                  float scale, int conj_smaps, cudaStream_t st) {
  if (use_fftp(p)) {
    Timed tm(p, EV_FFT, st);
    return fftp_type1(p, p->d_fw, smaps, img, T, accumulate, isign, scale, conj_smaps, st, p->spread_empty);
  }
  B200_TRY(exec_fft(p, p->d_fw, T, isign, st));
  Timed tm(p, EV_GRID, st);
  return k4b_crop(p, p->d_fw, smaps, img, T, accumulate, scale, conj_smaps, st);
}

// residual buffer of data_consistency / pipe: [ntrans_max][M], grown whenever the current M needs more
// than it holds (setpts only frees it when M exceeds the point capacity)
int ensure_ksp_tmp(b200_plan* p) {
  const size_t need = (size_t)p->ntrans_max * (size_t)(p->M > 0 ? p->M : 1) * sizeof(float2);
  if (p->d_ksp_tmp && p->ksp_tmp_bytes >= need) return B200_OK;
  if (p->d_ksp_tmp) CUDA_TRY(cudaFree(p->d_ksp_tmp));
  p->d_ksp_tmp = nullptr;
  p->ksp_tmp_bytes = 0;
  CUDA_TRY(cudaMalloc(&p->d_ksp_tmp, need));
  p->ksp_tmp_bytes = need;
  return B200_OK;
}

int check_exec(b200_plan* p, int T, bool need_fft) {
  if (!p) {
    b200_set_error("null plan");
    return B200_EINVAL;
  }
  if (T < 1 || T > p->ntrans_max) {
    b200_set_error("T=%d outside [1, n_trans_max=%d]", T, p->ntrans_max);
    return B200_EINVAL;
  }
  if (p->M < 0 || !p->pts_set) {
    b200_set_error("execute called before b200_plan_setpts");
    return B200_ESTATE;
  }
  if (need_fft && (p->flags & B200_SPREAD_ONLY)) {
    b200_set_error("plan was created with B200_SPREAD_ONLY");
    return B200_ESTATE;
  }
  if (!need_fft && !(p->flags & B200_SPREAD_ONLY)) {
    b200_set_error("b200_spread/b200_interp need a B200_SPREAD_ONLY plan");
    return B200_ESTATE;
  }
  return B200_OK;
}

}  // namespace

extern "C" {

int b200_abi_version(void) { return 1; }

int b200_launch_count(int64_t* kernels, int64_t* ffts, int reset) {
  if (kernels) *kernels = g_kernel_launches;
  if (ffts) *ffts = g_fft_execs;
  if (reset) g_kernel_launches = g_fft_execs = 0;
  return B200_OK;
}

int b200_plan_create(b200_plan** out, int dim, const int64_t* n_modes, int n_trans_max,
                     double eps, double upsampfac, int flags, int device) {
  if (!out || !n_modes || dim < 1 || dim > 3 || n_trans_max < 1 || !(eps > 0)) {
    b200_set_error("b200_plan_create: bad argument (dim=%d, n_trans_max=%d, eps=%g)", dim,
                   n_trans_max, eps);
    return B200_EINVAL;
  }
  for (int a = 0; a < dim; ++a)
    if (n_modes[a] < 1 || n_modes[a] > (1 << 20)) {
      b200_set_error("b200_plan_create: n_modes[%d]=%lld out of range", a, (long long)n_modes[a]);
      return B200_EINVAL;
    }
  DeviceGuard guard(device);
  if (!guard.ok) {
    b200_set_error("b200_plan_create: cannot select CUDA device %d (%s)", device,
                   cudaGetErrorString(cudaGetLastError()));
    return B200_ECUDA;
  }
  b200_plan* p = new b200_plan();
  p->flags = flags;
  p->device = device;
  p->ntrans_max = n_trans_max;
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
      p->num_sms = sms;
  }
  p->eps = eps;
  p->sigma = upsampfac > 0 ? upsampfac : 2.0;
  int w;
  double beta;
  es_kernel_params(eps, p->sigma, &w, &beta);
  p->beta = beta;
  p->cpar = 4.0 / ((double)w * w);
  Geom& g = p->g;
  g.dim = dim;
  g.w = w;
  g.nftot = 1;
  g.Ntot = 1;
  for (int a = 0; a < 3; ++a) {
    g.N[a] = a < dim ? (int)n_modes[a] : 1;
    if (a < dim) {
      if (flags & B200_SPREAD_ONLY) {
        g.nf[a] = g.N[a];
      } else {
        int target = (int)std::ceil(p->sigma * g.N[a]);
        if (target < 2 * w) target = 2 * w;
        g.nf[a] = next235even(target);
        // fastest axis: the tiled spread / interp kernels want a last 16-cell tile that is either whole or at
        // least w - 1 cells wide (a footprint may touch at most two tiles, spread_rows.cu `tiled_supported`);
        // a grid like 450 = 28 x 16 + 2 would send every transform to the point-driven kernels (224^3 x 8 coils:
        // 28 / 43 ms per op / adj_op against ~9 ms).  Take the next admissible size instead (450 -> 480).
        if (a == dim - 1 && dim >= 2 && !(flags & B200_EXACT_GRID))
          while (g.nf[a] % 16 != 0 && g.nf[a] % 16 < w - 1) g.nf[a] = next235even(g.nf[a] + 2);
      }
    } else {
      g.nf[a] = 1;
    }
    g.nftot *= g.nf[a];
    g.Ntot *= g.N[a];
  }
  // 3-D grids with factors 3 / 5 (N = 96, 192, 200, 224, 240 ...): when the next power of two is at most a third
  // larger on every axis the plan takes it.  The kernel (designed for sigma = 2) is only more accurate on the
  // finer grid, the zero-padding-aware FFT passes and the tile skipping of fft_pruned.cu / the row kernels apply
  // (they need powers of two), and that more than pays for the larger grid: 192^3 x 8 coils 10.2 / 9.5 ->
  // 8.2 / 7.8 ms per op / adj_op, 224^3 (next235even = 450, not a multiple of the tile width) 28 / 43 -> 8.7 / 8.5
  // ms.  Not in 2-D (measured slower) and not for B200_EXACT_GRID plans (Toeplitz needs exactly 2 N).
  if (dim == 3 && !(flags & (B200_SPREAD_ONLY | B200_DOUBLE | B200_EXACT_GRID))) {
    int p2[3];
    bool take = true, changes = false;
    for (int a = 0; a < 3; ++a) {
      p2[a] = 32;
      while (p2[a] < g.nf[a] && p2[a] < (1 << 20)) p2[a] <<= 1;
      // (next235even(target) <= next power of two >= target, so this is the power of two above the target too)
      take = take && (double)p2[a] <= 1.34 * (double)g.nf[a];
      changes = changes || p2[a] != g.nf[a];
    }
    if (take && changes) {
      g.nftot = 1;
      for (int a = 0; a < 3; ++a) {
        g.nf[a] = p2[a];
        g.nftot *= g.nf[a];
      }
    }
  }
  for (int a = 0; a < dim; ++a)
    if (g.nf[a] < w) {
      b200_set_error("grid size %d along axis %d is smaller than the kernel width %d", g.nf[a], a,
                     w);
      delete p;
      return B200_EINVAL;
    }
  // pencil bins: 1 cell along slow axes, BX cells along the fastest axis
  const int BX = 16;  // pencil-bin width = tile width of the row kernels (spread_rows.cu)
  p->nbins_tot = 1;
  for (int a = 0; a < 3; ++a) {
    g.bin[a] = 1;
    g.nbins[a] = a < dim ? g.nf[a] : 1;
  }
  g.bin[dim - 1] = BX;
  g.nbins[dim - 1] = (g.nf[dim - 1] + BX - 1) / BX;
  for (int a = 0; a < dim; ++a) p->nbins_tot *= g.nbins[a];
  if (dim >= 2) p->nbins_tot *= 2;  // interior / crossing sub-bins (setpts.cu)
  if (p->nbins_tot >= (1LL << 31)) {
    b200_set_error("too many bins (%lld)", p->nbins_tot);
    delete p;
    return B200_EINVAL;
  }

  if (flags & B200_DOUBLE) {
    // complex128 path (double_path.cu): own workspace and kernels, same entry points
    if (flags & B200_SPREAD_ONLY) {
      b200_set_error("B200_DOUBLE plans do not support B200_SPREAD_ONLY");
      delete p;
      return B200_EINVAL;
    }
    const int rc = dbl_init(p);
    if (rc != B200_OK) {
      b200_plan_destroy(p);
      return rc;
    }
    *out = p;
    return B200_OK;
  }

  KernelTables kt;
  es_fit_polynomial(w, beta, eps, &kt);
  g.deg = kt.deg;

  int st = B200_OK;
  auto fail = [&](int code) {
    b200_plan_destroy(p);
    return code;
  };
#define CT(expr)                                                                   \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                     cudaGetErrorString(_e));                                      \
      return fail(_e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA);     \
    }                                                                              \
  } while (0)
  CT(cudaMalloc(&p->d_poly, kt.poly.size() * sizeof(float)));
  CT(cudaMemcpy(p->d_poly, kt.poly.data(), kt.poly.size() * sizeof(float),
                cudaMemcpyHostToDevice));
  if (!(flags & B200_SPREAD_ONLY)) {
    for (int a = 0; a < dim; ++a) {
      std::vector<float> dv;
      es_deapod_vector(g.N[a], g.nf[a], w, beta, &dv);
      CT(cudaMalloc(&p->d_deapod[a], dv.size() * sizeof(float)));
      CT(cudaMemcpy(p->d_deapod[a], dv.data(), dv.size() * sizeof(float),
                    cudaMemcpyHostToDevice));
    }
  }
  CT(cudaMalloc(&p->d_bin_start, (size_t)(p->nbins_tot + 1) * sizeof(int32_t)));
  p->ws_bytes += (size_t)(p->nbins_tot + 1) * sizeof(int32_t);
  if (!(flags & B200_SPREAD_ONLY)) {
    const size_t fwb = (size_t)n_trans_max * g.nftot * sizeof(float2);
    CT(cudaMalloc(&p->d_fw, fwb));
    p->ws_bytes += fwb;
    // cuFFT: one batched plan (fft_batch grids per execution) + a single-grid plan for tails
    long long n[3] = {g.nf[0], g.nf[1], g.nf[2]};
    p->fft_batch = n_trans_max;
    size_t wsz = 0;
    cufftResult r = cufftCreate(&p->fft);
    if (r == CUFFT_SUCCESS)
      r = cufftMakePlanMany64(p->fft, dim, n, nullptr, 1, g.nftot, nullptr, 1, g.nftot, CUFFT_C2C,
                              p->fft_batch, &wsz);
    if (r != CUFFT_SUCCESS) {
      b200_set_error("cufftMakePlanMany(batch=%d) failed: %d", p->fft_batch, (int)r);
      return fail(B200_ECUFFT);
    }
    p->fft_ok = true;
    p->ws_bytes += wsz;
    if (p->fft_batch > 1) {
      size_t wsz1 = 0;
      r = cufftCreate(&p->fft1);
      if (r == CUFFT_SUCCESS)
        r = cufftMakePlanMany64(p->fft1, dim, n, nullptr, 1, g.nftot, nullptr, 1, g.nftot,
                                CUFFT_C2C, 1, &wsz1);
      if (r != CUFFT_SUCCESS) {
        b200_set_error("cufftMakePlanMany(batch=1) failed: %d", (int)r);
        return fail(B200_ECUFFT);
      }
      p->fft1_ok = true;
      p->ws_bytes += wsz1;
    } else {
      p->fft1 = p->fft;
    }
  }
  for (int i = 0; i < 10; ++i) CT(cudaEventCreate(&p->ev[i]));
  p->ev_ok = true;
#undef CT
  (void)st;
  *out = p;
  return B200_OK;
}

int b200_plan_destroy(b200_plan* p) {
  if (!p) return B200_OK;
  DeviceGuard guard(p->device);
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  tiled_free(p);
  dbl_free(p);
  fftp_free(p);
  fr(p->d_poly);
  for (int a = 0; a < 3; ++a) {
    fr(p->d_deapod[a]);
    fr(p->d_ones[a]);
    fr(p->d_tw[a]);
    fr(p->d_org_u[a]);
    fr(p->d_x1_u[a]);
    fr(p->d_org_s[a]);
    fr(p->d_x1_s[a]);
  }
  fr(p->d_key_u);
  fr(p->d_key_s);
  fr(p->d_perm);
  fr(p->d_iota);
  fr(p->d_bin_start);
  fr(p->d_sort_tmp);
  fr(p->d_fw);
  fr(p->d_ksp_tmp);
  fr(p->d_pipe_tmp);
  if (p->fft1_ok) cufftDestroy(p->fft1);
  if (p->fft_ok) cufftDestroy(p->fft);
  if (p->ev_ok)
    for (int i = 0; i < 10; ++i) cudaEventDestroy(p->ev[i]);
  delete p;
  return B200_OK;
}

int b200_plan_info(const b200_plan* p, int64_t info[16]) {
  if (!p || !info) {
    b200_set_error("b200_plan_info: null argument");
    return B200_EINVAL;
  }
  memset(info, 0, 16 * sizeof(int64_t));
  for (int a = 0; a < 3; ++a) {
    info[a] = p->g.nf[a];
    info[4 + a] = p->g.bin[a];
    info[7 + a] = p->g.nbins[a];
  }
  info[3] = p->g.w;
  info[10] = p->g.deg;
  info[11] = p->pts_set ? p->M : 0;
  info[12] = (int64_t)p->ws_bytes;
  return B200_OK;
}

int b200_plan_rows_class(b200_plan* p, int T, int64_t out[4]) {
  if (!p || !out) {
    b200_set_error("b200_plan_rows_class: null argument");
    return B200_EINVAL;
  }
  out[0] = out[1] = out[2] = out[3] = 0;
  if (p->dbl) {
    drows_info(p, out);
    return B200_OK;
  }
  if (!tiled_supported(p, T)) return B200_OK;
  tiled_class_info(p, T, out);
  return B200_OK;
}

int b200_plan_kernel_params(const b200_plan* p, double out[4]) {
  if (!p || !out) {
    b200_set_error("b200_plan_kernel_params: null argument");
    return B200_EINVAL;
  }
  out[0] = p->beta;
  out[1] = p->cpar;
  out[2] = p->sigma;
  out[3] = p->eps;
  return B200_OK;
}

int b200_plan_set_option(b200_plan* p, int key, int64_t value) {
  if (!p) {
    b200_set_error("null plan");
    return B200_EINVAL;
  }
  switch (key) {
    case 0: p->spread_method = (int)value; break;
    case 1: p->interp_method = (int)value; break;
    case 2: p->fft_method = (int)value; break;
    case 3: p->rows_dbg = (int)value; break;
    case 4:
      if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16 && value != 32) {
        b200_set_error("option 4 (smallest coil class of the row kernels) takes 0, 1, 2, 4, 8, 16 or 32");
        return B200_EINVAL;
      }
      p->rows_class = (int)value;
      break;
    case 5:
      p->fft_lookahead = value > 0 ? (int)value : 0;
      break;
    default:
      b200_set_error("unknown option key %d", key);
      return B200_EINVAL;
  }
  return B200_OK;
}

int b200_plan_enable_timing(b200_plan* p, int on) {
  if (!p) {
    b200_set_error("null plan");
    return B200_EINVAL;
  }
  p->timing = on != 0;
  for (int i = 0; i < 5; ++i) p->ev_used[i] = 0;
  return B200_OK;
}

int b200_plan_last_timings(b200_plan* p, float out[8]) {
  if (!p || !out) {
    b200_set_error("null argument");
    return B200_EINVAL;
  }
  DeviceGuard guard(p->device);
  for (int i = 0; i < 8; ++i) out[i] = 0.f;
  for (int s = 0; s < 5; ++s) {
    if (!p->ev_used[s]) continue;
    CUDA_TRY(cudaEventSynchronize(p->ev[2 * s + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev[2 * s], p->ev[2 * s + 1]));
    out[s] = ms;
  }
  return B200_OK;
}

int b200_plan_setpts(b200_plan* p, int64_t M, const float* xyz, void* stream) {
  if (!p || M < 0 || (M > 0 && !xyz) || M >= (1LL << 31)) {
    b200_set_error("b200_plan_setpts: bad argument (M=%lld)", (long long)M);
    return B200_EINVAL;
  }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  p->M = M;
  p->pts_set = false;
  if (p->dbl) {
    B200_TRY(dbl_setpts(p, (const double*)xyz, st));
    p->pts_set = true;
    return B200_OK;
  }
  tiled_invalidate(p);
  B200_TRY(k1_setpts(p, xyz, st));
  p->pts_set = true;
  return B200_OK;
}

int b200_plan_get_sort(b200_plan* p, int32_t* origin, float* x1, int32_t* key, int32_t* perm,
                       void* stream) {
  if (!p || !p->pts_set) {
    b200_set_error("b200_plan_get_sort: setpts has not been called");
    return B200_ESTATE;
  }
  if (p->dbl) {
    b200_set_error("b200_plan_get_sort: B200_DOUBLE plans do not sort their points");
    return B200_ESTATE;
  }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)p->M * 4;
  if (n == 0) return B200_OK;
  for (int a = 0; a < p->g.dim; ++a) {
    if (origin)
      CUDA_TRY(cudaMemcpyAsync(origin + (size_t)a * p->M, p->d_org_u[a], n,
                               cudaMemcpyDeviceToDevice, st));
    if (x1)
      CUDA_TRY(cudaMemcpyAsync(x1 + (size_t)a * p->M, p->d_x1_u[a], n, cudaMemcpyDeviceToDevice,
                               st));
  }
  if (key) CUDA_TRY(cudaMemcpyAsync(key, p->d_key_u, n, cudaMemcpyDeviceToDevice, st));
  if (perm) CUDA_TRY(cudaMemcpyAsync(perm, p->d_perm, n, cudaMemcpyDeviceToDevice, st));
  return B200_OK;
}

int b200_type2(b200_plan* p, const void* img, const void* smaps, void* ksp, int T, int isign,
               float scale, int conj_smaps, void* stream) {
  B200_TRY(check_exec(p, T, true));
  if (!img || !ksp) {
    b200_set_error("b200_type2: null buffer");
    return B200_EINVAL;
  }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dbl) return dbl_type2(p, img, smaps, ksp, T, isign, (double)scale, conj_smaps, st);
  B200_TRY(image_to_grid(p, (const float2*)img, (const float2*)smaps, T, isign, conj_smaps, st, true));
  return do_interp(p, p->d_fw, (float2*)ksp, T, scale, nullptr, st);
}

int b200_type1(b200_plan* p, const void* ksp, const float* density, const void* smaps, void* img,
               int T, int accumulate, int isign, float scale, int conj_smaps, void* stream) {
  B200_TRY(check_exec(p, T, true));
  if (!img || !ksp) {
    b200_set_error("b200_type1: null buffer");
    return B200_EINVAL;
  }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dbl)
    return dbl_type1(p, ksp, density, smaps, img, T, accumulate, isign, (double)scale, conj_smaps, st);
  p->spread_may_skip_empty = use_fftp(p) && p->g.dim == 3;
  const int rc = do_spread(p, (const float2*)ksp, density, p->d_fw, T, st);
  p->spread_may_skip_empty = false;
  B200_TRY(rc);
  return grid_to_image(p, (const float2*)smaps, (float2*)img, T, accumulate, isign, scale, conj_smaps,
                       st);
}

int b200_data_consistency(b200_plan* p, const void* img, const void* smaps, const void* obs,
                          const float* density, void* grad, int T, int accumulate, float scale,
                          void* stream) {
  B200_TRY(check_exec(p, T, true));
  if (!img || !obs || !grad) {
    b200_set_error("b200_data_consistency: null buffer");
    return B200_EINVAL;
  }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dbl)
    return dbl_data_consistency(p, img, smaps, obs, density, grad, T, accumulate, (double)scale, st);
  B200_TRY(ensure_ksp_tmp(p));
  B200_TRY(image_to_grid(p, (const float2*)img, (const float2*)smaps, T, -1, 0, st, true));
  // K5: residual fused into the interpolation epilogue
  B200_TRY(do_interp(p, p->d_fw, p->d_ksp_tmp, T, scale, (const float2*)obs, st));
  p->spread_may_skip_empty = use_fftp(p) && p->g.dim == 3;
  const int rc = do_spread(p, p->d_ksp_tmp, density, p->d_fw, T, st);
  p->spread_may_skip_empty = false;
  B200_TRY(rc);
  return grid_to_image(p, (const float2*)smaps, (float2*)grad, T, accumulate, +1, scale, 0, st);
}

int b200_toeplitz_apply(b200_plan* p, const void* img, const void* smaps, const float* kern,
                        void* out, int T, int accumulate, float scale, void* stream) {
  if (!p || !img || !kern || !out) {
    b200_set_error("b200_toeplitz_apply: null argument");
    return B200_EINVAL;
  }
  if (T < 1 || T > p->ntrans_max) {
    b200_set_error("T=%d outside [1, n_trans_max=%d]", T, p->ntrans_max);
    return B200_EINVAL;
  }
  if ((p->flags & B200_SPREAD_ONLY) || p->dbl) {
    b200_set_error("b200_toeplitz_apply needs a full single-precision plan");
    return B200_ESTATE;
  }
  for (int a = 0; a < p->g.dim; ++a)
    if (p->g.nf[a] != 2 * p->g.N[a]) {
      b200_set_error("b200_toeplitz_apply: the oversampled grid must be exactly 2 N (axis %d: N=%d, nf=%d)",
                     a, p->g.N[a], p->g.nf[a]);
      return B200_EINVAL;
    }
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int a = 0; a < p->g.dim; ++a) {
    if (p->d_ones[a]) continue;
    std::vector<float> ones(p->g.N[a], 1.f);
    CUDA_TRY(cudaMalloc(&p->d_ones[a], ones.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpy(p->d_ones[a], ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // zero-pad -> FFT -> multiply by the (real) spectrum of the Toeplitz embedding -> inverse FFT ->
  // crop, with the sensitivity-map multiplies and the coil sum fused like in type 2 / type 1.
  // Circular convolution commutes with shifts, so padding / cropping at the mode-centred
  // position of the NUFFT grid instead of the top-left corner gives the same values.
  p->unit_deapod = true;
  int rc = B200_OK;
  if (use_fftp(p)) {
    Timed tm(p, EV_FFT, st);
    rc = fftp_type2(p, (const float2*)img, (const float2*)smaps, p->d_fw, T, -1, 0, st, kern);
    if (rc == B200_OK)
      rc = fftp_type1(p, p->d_fw, (const float2*)smaps, (float2*)out, T, accumulate, +1, scale, 0, st);
  } else {
    rc = k4a_pad(p, (const float2*)img, (const float2*)smaps, p->d_fw, T, 0, st);
    if (rc == B200_OK) rc = exec_fft(p, p->d_fw, T, -1, st);
    if (rc == B200_OK) rc = k_mul_real(p, p->d_fw, kern, T, st);
    if (rc == B200_OK) rc = exec_fft(p, p->d_fw, T, +1, st);
    if (rc == B200_OK) rc = k4b_crop(p, p->d_fw, (const float2*)smaps, (float2*)out, T, accumulate, scale, 0, st);
  }
  p->unit_deapod = false;
  return rc;
}

int b200_spread(b200_plan* p, const void* ksp, void* grid, int T, void* stream) {
  B200_TRY(check_exec(p, T, false));
  DeviceGuard guard(p->device);
  return do_spread(p, (const float2*)ksp, nullptr, (float2*)grid, T, (cudaStream_t)stream);
}

int b200_interp(b200_plan* p, const void* grid, void* ksp, int T, void* stream) {
  B200_TRY(check_exec(p, T, false));
  DeviceGuard guard(p->device);
  return do_interp(p, (const float2*)grid, (float2*)ksp, T, 1.f, nullptr, (cudaStream_t)stream);
}

int b200_pipe_iteration(b200_plan* p, float* d, void* stream) {
  B200_TRY(check_exec(p, 1, false));
  DeviceGuard guard(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  B200_TRY(ensure_ksp_tmp(p));
  if (!p->d_fw) {
    CUDA_TRY(cudaMalloc(&p->d_fw, (size_t)p->g.nftot * sizeof(float2)));
    p->ws_bytes += (size_t)p->g.nftot * sizeof(float2);
  }
  B200_TRY(k_real_to_cpx(p, d, p->d_ksp_tmp, st));
  B200_TRY(do_spread(p, p->d_ksp_tmp, nullptr, p->d_fw, 1, st));
  B200_TRY(do_interp(p, p->d_fw, p->d_ksp_tmp, 1, 1.f, nullptr, st));
  return k_pipe_update(p, d, p->d_ksp_tmp, st);
}

}  // extern "C"
