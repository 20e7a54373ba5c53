// vecops.cu -- the vector updates of the iterative solvers (cg, lsqr, lsmr), one pass over memory each.
//
// The reference's solvers (src/mrinufft/extras/optim.py: lsqr 249-495, lsmr 498-798, cg 801-902) update
// their iterates with array expressions: every `x += t1 * w`, `w = v + t2 * w`, `norm(w)` is a pass (or
// several, with temporaries) over the whole image or k-space batch.  Here each group of updates that reads the
// same vectors is one kernel: complex vectors of shape (B, n), per-batch scalars passed by value, the squared
// norms / dot products the iteration needs accumulated in double on the way (block reduction + one atomicAdd
// per block and quantity).  HBM-bound by construction: 8-byte (complex64) or 16-byte (complex128) coalesced
// accesses, grid sized to a few waves of 148 SMs.
#include "common.cuh"

namespace {

constexpr int VT = 256;    // threads per CTA
constexpr int VB = 32;     // batches per launch (their scalars travel in the kernel parameters)

struct cd {
  double x, y;
};
__device__ __forceinline__ cd to_cd(float2 a) { return {(double)a.x, (double)a.y}; }
__device__ __forceinline__ cd to_cd(double2 a) { return {a.x, a.y}; }

template <class V>
struct Real;
template <>
struct Real<float2> {
  typedef float type;
};
template <>
struct Real<double2> {
  typedef double type;
};

template <class V>
__device__ __forceinline__ V mk(typename Real<V>::type x, typename Real<V>::type y) {
  V v;
  v.x = x;
  v.y = y;
  return v;
}
// a + s b with a complex scalar s of the vector's precision
template <class V>
__device__ __forceinline__ V cfma(V s, V b, V a) {
  return mk<V>(a.x + s.x * b.x - s.y * b.y, a.y + s.x * b.y + s.y * b.x);
}
template <class V>
__device__ __forceinline__ V cmul(V s, V b) {
  return mk<V>(s.x * b.x - s.y * b.y, s.x * b.y + s.y * b.x);
}

struct Scalars {
  double2 s[3][VB];   // up to three complex scalars per batch
};

// ---- the update functors: operator()(batch in launch, element index, accumulators) ----------------------
// out = a x + b y  (y may be null: out = a x);  acc[0] += |out|^2
template <class V>
struct Axpby {
  static constexpr int NRED = 1;
  V* out;
  const V* x;
  const V* y;
  __device__ __forceinline__ void operator()(const Scalars& sc, int b, long long i, double (&acc)[NRED]) const {
    typedef typename Real<V>::type R;
    const V a = mk<V>((R)sc.s[0][b].x, (R)sc.s[0][b].y);
    V r = cmul(a, x[i]);
    if (y) r = cfma(mk<V>((R)sc.s[1][b].x, (R)sc.s[1][b].y), y[i], r);
    out[i] = r;
    acc[0] += (double)r.x * r.x + (double)r.y * r.y;
  }
};

// cg (optim.py:866-880): acc = { |x|^2, Re / Im sum x (x - y), Re / Im sum y y }  (un-conjugated products)
template <class V>
struct Dots {
  static constexpr int NRED = 5;
  const V* x;
  const V* y;
  __device__ __forceinline__ void operator()(const Scalars&, int, long long i, double (&acc)[NRED]) const {
    const cd a = to_cd(x[i]), g = to_cd(y[i]);
    const cd d = {a.x - g.x, a.y - g.y};
    acc[0] += a.x * a.x + a.y * a.y;
    acc[1] += a.x * d.x - a.y * d.y;
    acc[2] += a.x * d.y + a.y * d.x;
    acc[3] += g.x * g.x - g.y * g.y;
    acc[4] += 2.0 * g.x * g.y;
  }
};

// cg (optim.py:881-883): v = g + beta v ;  x = x - v / L        scalars: beta (complex), -1 / L
template <class V>
struct CgStep {
  static constexpr int NRED = 1;
  V* x;
  V* v;
  const V* g;
  __device__ __forceinline__ void operator()(const Scalars& sc, int b, long long i, double (&acc)[NRED]) const {
    typedef typename Real<V>::type R;
    const V vel = cfma(mk<V>((R)sc.s[0][b].x, (R)sc.s[0][b].y), v[i], g[i]);
    v[i] = vel;
    const R m = (R)sc.s[1][b].x;
    const V xi = x[i];
    x[i] = mk<V>(xi.x + m * vel.x, xi.y + m * vel.y);
  }
};

// lsqr (optim.py:441-446): acc[0] += |w|^2 ;  x += t1 w ;  w = v + t2 w        (real scalars t1, t2)
template <class V>
struct LsqrStep {
  static constexpr int NRED = 1;
  V* x;
  V* w;
  const V* v;
  __device__ __forceinline__ void operator()(const Scalars& sc, int b, long long i, double (&acc)[NRED]) const {
    typedef typename Real<V>::type R;
    const R t1 = (R)sc.s[0][b].x, t2 = (R)sc.s[1][b].x;
    const V wi = w[i], xi = x[i], vi = v[i];
    acc[0] += (double)wi.x * wi.x + (double)wi.y * wi.y;
    x[i] = mk<V>(xi.x + t1 * wi.x, xi.y + t1 * wi.y);
    w[i] = mk<V>(vi.x + t2 * wi.x, vi.y + t2 * wi.y);
  }
};

// lsmr (optim.py:716-724): hbar = h + a hbar ;  x += b hbar ;  h = v + c h ;  acc[0] += |x|^2
template <class V>
struct LsmrStep {
  static constexpr int NRED = 1;
  V* x;
  V* hbar;
  V* h;
  const V* v;
  __device__ __forceinline__ void operator()(const Scalars& sc, int b, long long i, double (&acc)[NRED]) const {
    typedef typename Real<V>::type R;
    const R a = (R)sc.s[0][b].x, bb = (R)sc.s[1][b].x, c = (R)sc.s[2][b].x;
    const V hi = h[i], hb = hbar[i], xi = x[i], vi = v[i];
    const V hbn = mk<V>(hi.x + a * hb.x, hi.y + a * hb.y);
    const V xn = mk<V>(xi.x + bb * hbn.x, xi.y + bb * hbn.y);
    hbar[i] = hbn;
    x[i] = xn;
    h[i] = mk<V>(vi.x + c * hi.x, vi.y + c * hi.y);
    acc[0] += (double)xn.x * xn.x + (double)xn.y * xn.y;
  }
};

// grid.x = CTAs per batch, grid.y = batches of this launch; `red` [batch][NRED] doubles or null
template <class Op>
__global__ void __launch_bounds__(VT) k_vec(const Op op, const Scalars sc, long long n, long long first, double* red) {
  constexpr int NRED = Op::NRED;
  const int b = blockIdx.y;
  const long long base = (first + b) * n;
  double acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
  const long long stride = (long long)gridDim.x * VT;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += stride) op(sc, b, base + i, acc);
  if (!red) return;
  __shared__ double part[VT / 32][NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    double a = acc[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = a;
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < VT / 32; ++w) a += part[w][threadIdx.x];
    atomicAdd(red + (first + b) * NRED + threadIdx.x, a);
  }
}

// scalars: host arrays of B complex doubles (re, im interleaved), or null
template <class Op>
int run(const Op& op, int B, long long n, const double* const (&sc)[3], double* red, cudaStream_t st) {
  if (B < 1 || n < 1) {
    b200_set_error("b200_vec_*: empty vectors (B=%d, n=%lld)", B, n);
    return B200_EINVAL;
  }
  if (red) CUDA_TRY(cudaMemsetAsync(red, 0, (size_t)B * Op::NRED * sizeof(double), st));
  for (int first = 0; first < B; first += VB) {
    const int nb = B - first < VB ? B - first : VB;
    Scalars s{};
    for (int k = 0; k < 3; ++k)
      if (sc[k])
        for (int b = 0; b < nb; ++b) s.s[k][b] = make_double2(sc[k][2 * (first + b)], sc[k][2 * (first + b) + 1]);
    long long per = (n + VT - 1) / VT;
    const long long cap = (8LL * B200_NUM_SMS * 4 + nb - 1) / nb;   // ~ four waves of 8 CTAs per SM over the batches
    if (per > cap) per = cap;
    k_vec<Op><<<dim3((unsigned)per, nb), VT, 0, st>>>(op, s, n, first, red);
    CHECK_LAUNCH();
  }
  return B200_OK;
}

template <class V>
int dispatch(int what, void* const* p, int B, long long n, const double* const (&sc)[3], double* red, cudaStream_t st) {
  switch (what) {
    case 0: return run(Axpby<V>{(V*)p[0], (const V*)p[1], (const V*)p[2]}, B, n, sc, red, st);
    case 1: return run(Dots<V>{(const V*)p[0], (const V*)p[1]}, B, n, sc, red, st);
    case 2: return run(CgStep<V>{(V*)p[0], (V*)p[1], (const V*)p[2]}, B, n, sc, nullptr, st);
    case 3: return run(LsqrStep<V>{(V*)p[0], (V*)p[1], (const V*)p[2]}, B, n, sc, red, st);
    case 4: return run(LsmrStep<V>{(V*)p[0], (V*)p[1], (V*)p[2], (const V*)p[3]}, B, n, sc, red, st);
  }
  b200_set_error("b200_vec_*: unknown operation %d", what);
  return B200_EINVAL;
}

int entry(int what, void* const* p, int np, int B, int64_t n, const double* s0, const double* s1, const double* s2,
          double* red, int dbl, void* stream) {
  for (int i = 0; i < np; ++i)
    if (!p[i]) {
      b200_set_error("b200_vec_*: null vector");
      return B200_EINVAL;
    }
  const double* const sc[3] = {s0, s1, s2};
  return dbl ? dispatch<double2>(what, p, B, n, sc, red, (cudaStream_t)stream)
             : dispatch<float2>(what, p, B, n, sc, red, (cudaStream_t)stream);
}

// ---- temporal weights of the off-resonance-corrected operator (virtual coil index l C + c) ----------------
// combine: y[b, c, k] = sum_l kv[b, l, c, k] * Bw[k % NK, l];  expand: kv[b, l, c, k] = conj(Bw[k % NK, l]) * y[b, c, k]
template <bool EXPAND>
__global__ void __launch_bounds__(VT) k_orc(float2* __restrict__ kv, const float2* __restrict__ bw, float2* __restrict__ y,
                                            int L, int C, long long K, int NK) {
  const long long k = (long long)blockIdx.x * VT + threadIdx.x;
  if (k >= K) return;
  const int c = blockIdx.y, b = blockIdx.z;
  const float2* w = bw + (size_t)(k % NK) * L;
  float2* v = kv + ((size_t)b * L * C + c) * K + k;
  float2* yy = y + ((size_t)b * C + c) * K + k;
  if (EXPAND) {
    const float2 a = *yy;
    for (int l = 0; l < L; ++l) {
      const float2 t = w[l];
      v[(size_t)l * C * K] = make_float2(t.x * a.x + t.y * a.y, t.x * a.y - t.y * a.x);
    }
  } else {
    float2 acc = make_float2(0.f, 0.f);
    for (int l = 0; l < L; ++l) {
      const float2 t = w[l], a = v[(size_t)l * C * K];
      acc.x += t.x * a.x - t.y * a.y;
      acc.y += t.x * a.y + t.y * a.x;
    }
    *yy = acc;
  }
}

}  // namespace

extern "C" int b200_orc_weights(void* kv, const void* bw, void* y, int B, int L, int C, int64_t K, int NK, int expand,
                                void* stream) {
  if (!kv || !bw || !y || B < 1 || L < 1 || C < 1 || K < 1 || NK < 1 || C > 65535 || B > 65535) {
    b200_set_error("b200_orc_weights: bad arguments");
    return B200_EINVAL;
  }
  dim3 grid((unsigned)((K + VT - 1) / VT), C, B);
  if (expand) k_orc<true><<<grid, VT, 0, (cudaStream_t)stream>>>((float2*)kv, (const float2*)bw, (float2*)y, L, C, K, NK);
  else k_orc<false><<<grid, VT, 0, (cudaStream_t)stream>>>((float2*)kv, (const float2*)bw, (float2*)y, L, C, K, NK);
  CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_vec_axpby(void* out, const void* x, const void* y, const double* a, const double* b, int B,
                              int64_t n, double* sumsq, int dbl, void* stream) {
  if (!a || (y && !b)) {
    b200_set_error("b200_vec_axpby: null scalars");
    return B200_EINVAL;
  }
  void* p[3] = {out, const_cast<void*>(x), const_cast<void*>(y)};
  return entry(0, p, 2, B, n, a, b, nullptr, sumsq, dbl, stream);
}

extern "C" int b200_vec_cg_dots(const void* gnew, const void* gold, int B, int64_t n, double* out5, int dbl,
                                void* stream) {
  if (!out5) {
    b200_set_error("b200_vec_cg_dots: null output");
    return B200_EINVAL;
  }
  void* p[2] = {const_cast<void*>(gnew), const_cast<void*>(gold)};
  return entry(1, p, 2, B, n, nullptr, nullptr, nullptr, out5, dbl, stream);
}

extern "C" int b200_vec_cg_step(void* x, void* v, const void* g, const double* beta, const double* minus_inv_l,
                                int B, int64_t n, int dbl, void* stream) {
  if (!beta || !minus_inv_l) {
    b200_set_error("b200_vec_cg_step: null scalars");
    return B200_EINVAL;
  }
  void* p[3] = {x, v, const_cast<void*>(g)};
  return entry(2, p, 3, B, n, beta, minus_inv_l, nullptr, nullptr, dbl, stream);
}

extern "C" int b200_vec_lsqr_step(void* x, void* w, const void* v, const double* t1, const double* t2, int B,
                                  int64_t n, double* sumsq_w, int dbl, void* stream) {
  if (!t1 || !t2) {
    b200_set_error("b200_vec_lsqr_step: null scalars");
    return B200_EINVAL;
  }
  void* p[3] = {x, w, const_cast<void*>(v)};
  return entry(3, p, 3, B, n, t1, t2, nullptr, sumsq_w, dbl, stream);
}

extern "C" int b200_vec_lsmr_step(void* x, void* hbar, void* h, const void* v, const double* a, const double* b,
                                  const double* c, int B, int64_t n, double* sumsq_x, int dbl, void* stream) {
  if (!a || !b || !c) {
    b200_set_error("b200_vec_lsmr_step: null scalars");
    return B200_EINVAL;
  }
  void* p[4] = {x, hbar, h, const_cast<void*>(v)};
  return entry(4, p, 4, B, n, a, b, c, sumsq_x, dbl, stream);
}
