// common.cuh -- plan structure, error plumbing and launch helpers shared by every
// translation unit of libb200nufft.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/b200nufft.h"

#define B200_MAX_W 16
#define B200_MAX_DEG 15
#define B200_NUM_SMS 148

// ---------------------------------------------------------------- error plumbing
void b200_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      b200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                     cudaGetErrorString(_e));                                       \
      return B200_ECUDA;                                                            \
    }                                                                               \
  } while (0)

#define CUFFT_TRY(expr)                                                             \
  do {                                                                              \
    cufftResult _r = (expr);                                                        \
    if (_r != CUFFT_SUCCESS) {                                                      \
      b200_set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr,      \
                     (int)_r);                                                      \
      return B200_ECUFFT;                                                           \
    }                                                                               \
  } while (0)

#define B200_TRY(expr)                                                              \
  do {                                                                              \
    int _s = (expr);                                                                \
    if (_s != B200_OK) return _s;                                                   \
  } while (0)

// process-wide launch counters (bench.py's `gpu_launches`)
extern long long g_kernel_launches;
extern long long g_fft_execs;
#define COUNT_LAUNCH() (++g_kernel_launches)

#define CHECK_LAUNCH()                                                              \
  do {                                                                              \
    COUNT_LAUNCH();                                                                 \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      b200_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,              \
                     cudaGetErrorString(_e));                                       \
      return B200_ECUDA;                                                            \
    }                                                                               \
  } while (0)

// ---------------------------------------------------------------- geometry passed to kernels by value
struct Geom {
  int dim;
  int w;           // kernel width
  int deg;         // polynomial degree
  int nf[3];       // fine grid size per axis (axis dim-1 fastest); unused axes = 1
  int N[3];        // image size per axis
  int bin[3];      // bin size per axis (cells)
  int nbins[3];    // bins per axis
  long long nftot; // prod nf
  long long Ntot;  // prod N
};

// fft_any.cu: unnormalised in-place FFT (sign < 0: exp(-i ...)) of T oversampled grids, any axis lengths
int fft_any_c64(float2* fw, int T, const Geom& g, int sign, cudaStream_t st);
int fft_any_c128(double2* fw, int T, const Geom& g, int sign, cudaStream_t st);

// ---------------------------------------------------------------- the plan
struct b200_plan {
  Geom g;
  int flags = 0;
  int device = 0;
  int ntrans_max = 1;
  int num_sms = B200_NUM_SMS;
  double eps = 1e-6, sigma = 2.0, beta = 0, cpar = 0;

  // device tables
  float* d_poly = nullptr;                       // [(deg+1)][w]: coefficient k of tap i at k*w+i
  float* d_deapod[3] = {nullptr, nullptr, nullptr};  // N[a] floats: (-1)^k / phihat(k)
  float2* d_tw[3] = {nullptr, nullptr, nullptr};     // nf[a] roots of unity (fft_pruned.cu)
  float* d_ones[3] = {nullptr, nullptr, nullptr};    // N[a] ones: "no deapodisation" (Toeplitz apply)
  bool unit_deapod = false;                          // grid passes use d_ones instead of d_deapod
  const float* dvec(int a) const { return unit_deapod ? d_ones[a] : d_deapod[a]; }

  // points
  long long M = 0, Mcap = 0;
  int32_t* d_org_u[3] = {nullptr, nullptr, nullptr};  // footprint origin, unsorted
  float* d_x1_u[3] = {nullptr, nullptr, nullptr};     // first-tap offset, unsorted
  int32_t* d_key_u = nullptr;                         // bin key, unsorted
  int32_t* d_key_s = nullptr;                         // bin key, sorted
  int32_t* d_perm = nullptr;                          // sorted position -> point index
  int32_t* d_iota = nullptr;
  int32_t* d_org_s[3] = {nullptr, nullptr, nullptr};  // sorted copies
  float* d_x1_s[3] = {nullptr, nullptr, nullptr};
  int32_t* d_bin_start = nullptr;                     // nbins_tot + 1
  long long nbins_tot = 0;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;

  // workspace
  float2* d_fw = nullptr;       // [ntrans_max][nftot] oversampled grids
  float2* d_ksp_tmp = nullptr;  // [ntrans_max][M] residual of data_consistency
  size_t ksp_tmp_bytes = 0;     // its capacity: M may grow again after it was sized for a smaller M
  float* d_pipe_tmp = nullptr;
  size_t ws_bytes = 0;

  cufftHandle fft = 0, fft1 = 0;  // batched plan and single-grid plan (tails)
  bool fft_ok = false, fft1_ok = false;
  int fft_batch = 0;
  bool pts_set = false;

  // state of the tiled spread/interp kernels (spread_tiled.cu)
  void* tiled = nullptr;
  // complex128 path (double_path.cu): non-null for plans created with B200_DOUBLE
  void* dbl = nullptr;
  void* drows = nullptr;  // state of its tile-owned spreader (double_rows.cu)
  // tensor maps of the TMA variant of the FFT passes (fft_pruned.cu)
  void* tma = nullptr;
  // spreading into a grid that the fused FFT passes consume next: tiles without visitors are neither
  // zero-filled by the spreader nor read by the first FFT pass (flag byte per tile, see spread_rows.cu)
  bool spread_may_skip_empty = false;       // set by the caller of do_spread
  const uint32_t* spread_empty = nullptr;   // set by the spreader: valid for the grid it just wrote
  const uint32_t* interp_unread = nullptr;  // set by the type-2 FFT: tiles of the grid it left unwritten
  int empty_nyh = 0, empty_nbx = 0;

  // options
  int spread_method = 0, interp_method = 0, fft_method = 0;
  int rows_dbg = 0;  // option 3: timing-experiment switches of the spreading row kernel (see b200nufft.h)
  int fft_lookahead = 0;  // option 5: look-ahead (in CTAs) of the L2 prefetch of the strided FFT passes (0: off)
  int rows_class = 0;  // option 4: smallest coil class the row kernels may pick (0: by the call's coil count)

  // timing
  bool timing = false;
  cudaEvent_t ev[10] = {};
  bool ev_ok = false;
  int ev_used[5] = {0, 0, 0, 0, 0};
};

// ---------------------------------------------------------------- host-side kernel math (es_kernel_host.cpp)
struct KernelTables {
  int w, deg;
  double beta;
  std::vector<float> poly;  // (deg+1) * w
};
void es_kernel_params(double eps, double sigma, int* w, double* beta);
int next235even(int n);
void es_fit_polynomial(int w, double beta, double eps, KernelTables* out);
void es_deapod_vector(int n, int nf, int w, double beta, std::vector<float>* out);

// ---------------------------------------------------------------- kernels (implemented in the .cu files)
int k1_setpts(b200_plan* p, const float* xyz, cudaStream_t st);
int k1_reserve_points(b200_plan* p, long long M);       // point arrays of the plan for M points
int k1_sort_origins(b200_plan* p, cudaStream_t st);     // sort by bin key from p->d_org_u (complex128 path)
int k2_spread(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
              cudaStream_t st);
int k3_interp(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
              const float2* obs, const float* density, cudaStream_t st);
int k4a_pad(b200_plan* p, const float2* img, const float2* smaps, float2* fw, int T,
            int conj_smaps, cudaStream_t st);
int k4b_crop(b200_plan* p, const float2* fw, const float2* smaps, float2* img, int T,
             int accumulate, float scale, int conj_smaps, cudaStream_t st);
int k_pipe_update(b200_plan* p, float* d, const float2* ksp, cudaStream_t st);
int k_real_to_cpx(b200_plan* p, const float* d, float2* out, cudaStream_t st);

// fused pad/crop + zero-padding-aware FFT passes (fft_pruned.cu)
bool fftp_supported(const b200_plan* p);
int fftp_type2(b200_plan* p, const float2* img, const float2* smaps, float2* fw, int T, int isign,
               int conj_smaps, cudaStream_t st, const float* mul = nullptr,
               const uint32_t* unread = nullptr);
int k_mul_real(b200_plan* p, float2* fw, const float* kern, int T, cudaStream_t st);
int fftp_type1(b200_plan* p, float2* fw, const float2* smaps, float2* img, int T, int accumulate,
               int isign, float scale, int conj_smaps, cudaStream_t st, const uint32_t* empty = nullptr);
void fftp_free(b200_plan* p);

// complex128 path (double_path.cu)
int dbl_init(b200_plan* p);
void dbl_free(b200_plan* p);
int dbl_setpts(b200_plan* p, const double* xyz, cudaStream_t st);
int dbl_type2(b200_plan* p, const void* img, const void* smaps, void* ksp, int T, int isign, double scale,
              int conj_smaps, cudaStream_t st);
int dbl_type1(b200_plan* p, const void* ksp, const void* density, const void* smaps, void* img, int T,
              int accumulate, int isign, double scale, int conj_smaps, cudaStream_t st);
int dbl_data_consistency(b200_plan* p, const void* img, const void* smaps, const void* obs,
                         const void* density, void* grad, int T, int accumulate, double scale,
                         cudaStream_t st);
// tile-owned spread / interp in double (double_rows.cu); both return 1 when they cannot serve the plan
bool drows_supported(const b200_plan* p);
int drows_setpts(b200_plan* p, const double* const* x1_unsorted, cudaStream_t st);
int drows_spread(b200_plan* p, const double2* ksp, const double* density, double2* fw, int T, cudaStream_t st);
void drows_free(b200_plan* p);
void drows_info(const b200_plan* p, int64_t out[4]);  // b200_plan_rows_class of a complex128 plan

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Function attributes (dynamic shared-memory limit) are per device: a launcher sets them on the first launch
// on every device of the process, not once per process.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    const bool f = !done[d];
    done[d] = true;
    return f;
  }
};
