// fft_smem.cuh -- Stockham autosort FFT of ANY length on rows held in shared memory (complex64 or complex128).
// Shared by stack_fftz.cu (z transform of the stacked operator) and fft_any.cu (grids that are not powers of
// two, complex128 grids).  Radix 4 / 2 / 3 / 5 / 7 butterflies in registers, any other prime factor p as a direct
// p-point DFT (O(p^2) per butterfly: only what the factorisation leaves); twiddles from one table
// exp(-2 pi i t / L) computed in double on the host, once per (device, L, precision).
#pragma once
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace fftsm {

constexpr int ZT = 256;      // threads per CTA
constexpr int MAXRAD = 24;   // radices of a length < 2^24

struct Radices {
  int n;
  int r[MAXRAD];
};

// what the stage loop needs to know about the rows
struct Rows {
  int Z;      // length
  int rows;   // rows in the tile (a power of two <= 16)
  int zp;     // row stride in elements (odd: row-fastest accesses spread over the banks)
  int rshift; // log2(rows)
  Radices rad;
  unsigned zmagic;          // fdiv magic of Z (tile loads / stores: element index -> (row, position))
  unsigned smagic[MAXRAD];  // per stage: floor(2^32 / s) + 1 for the stage's stride s (0 for s = 1), see fdiv
};

// n / d for n d < 2^32 through the precomputed magic = floor(2^32 / d) + 1 (0 stands for d = 1): one IMAD.HI
// instead of the ~25 instructions of a run-time 32-bit division
__host__ __device__ inline unsigned fdiv_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(0x100000000ull / d) + 1u; }
__device__ __forceinline__ int fdiv(int n, unsigned magic) { return magic ? (int)__umulhi((unsigned)n, magic) : n; }

template <class V>
struct RealOf;
template <>
struct RealOf<float2> {
  typedef float type;
};
template <>
struct RealOf<double2> {
  typedef double type;
};
template <class V>
__device__ __forceinline__ V mkc(typename RealOf<V>::type x, typename RealOf<V>::type y) {
  V v;
  v.x = x;
  v.y = y;
  return v;
}
template <class V>
__device__ __forceinline__ V cmulf(V a, V b) {
  return mkc<V>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class V>
__device__ __forceinline__ V caddf(V a, V b) {
  return mkc<V>(a.x + b.x, a.y + b.y);
}
template <class V>
__device__ __forceinline__ V csubf(V a, V b) {
  return mkc<V>(a.x - b.x, a.y - b.y);
}
// multiply by exp(-+ i pi / 2): -i forward, +i inverse
template <bool INV, class V>
__device__ __forceinline__ V rot90(V a) {
  return INV ? mkc<V>(-a.y, a.x) : mkc<V>(a.y, -a.x);
}

// One Stockham stage of radix R on one row: butterfly (p, q), sub-length n = R m, stride s (s n = Z).
//   y[q + s (R p + k)] = w_n^(p k) sum_j x[q + s (p + j m)] w_R^(j k),   w_L = tw[Z / L]
template <int R, bool INV, class V>
__device__ __forceinline__ void butterfly(const V* __restrict__ src, V* __restrict__ dst, const V* __restrict__ tw,
                                          int p, int q, int s, int m, int Z) {
  V a[R];
#pragma unroll
  for (int j = 0; j < R; ++j) a[j] = src[q + s * (p + j * m)];
  V o[R];
  if constexpr (R == 2) {
    o[0] = caddf(a[0], a[1]);
    o[1] = csubf(a[0], a[1]);
  } else if constexpr (R == 4) {
    const V t0 = caddf(a[0], a[2]), t1 = csubf(a[0], a[2]);
    const V t2 = caddf(a[1], a[3]), t3 = rot90<INV>(csubf(a[1], a[3]));
    o[0] = caddf(t0, t2);
    o[1] = caddf(t1, t3);
    o[2] = csubf(t0, t2);
    o[3] = csubf(t1, t3);
  } else if constexpr (R == 8) {
    // two 4-point transforms of the even / odd inputs, joined by the eighth roots of unity
    typedef typename RealOf<V>::type T;
    const T h = (T)0.70710678118654752440;
    V e[4], d[4];
    {
      const V t0 = caddf(a[0], a[4]), t1 = csubf(a[0], a[4]);
      const V t2 = caddf(a[2], a[6]), t3 = rot90<INV>(csubf(a[2], a[6]));
      e[0] = caddf(t0, t2); e[1] = caddf(t1, t3); e[2] = csubf(t0, t2); e[3] = csubf(t1, t3);
    }
    {
      const V t0 = caddf(a[1], a[5]), t1 = csubf(a[1], a[5]);
      const V t2 = caddf(a[3], a[7]), t3 = rot90<INV>(csubf(a[3], a[7]));
      d[0] = caddf(t0, t2); d[1] = caddf(t1, t3); d[2] = csubf(t0, t2); d[3] = csubf(t1, t3);
    }
    // d[k] *= w8^k: w8 = (1 -+ i) / sqrt 2, w8^2 = -+ i, w8^3 = (-1 -+ i) / sqrt 2
    d[1] = INV ? mkc<V>((d[1].x - d[1].y) * h, (d[1].x + d[1].y) * h) : mkc<V>((d[1].x + d[1].y) * h, (d[1].y - d[1].x) * h);
    d[2] = rot90<INV>(d[2]);
    d[3] = INV ? mkc<V>((-d[3].x - d[3].y) * h, (d[3].x - d[3].y) * h) : mkc<V>((d[3].y - d[3].x) * h, (-d[3].x - d[3].y) * h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[k] = caddf(e[k], d[k]);
      o[k + 4] = csubf(e[k], d[k]);
    }
  } else {
    V w[R];
#pragma unroll
    for (int j = 1; j < R; ++j) w[j] = tw[j * (Z / R)];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      V acc = a[0];
#pragma unroll
      for (int j = 1; j < R; ++j) {
        const int e = (j * k) % R;  // compile time
        if (e == 0) acc = caddf(acc, a[j]);
        else acc = caddf(acc, cmulf(a[j], w[e]));
      }
      o[k] = acc;
    }
  }
  const int t1 = p * s;  // p k s < Z for every k < R
  V* d = dst + q + s * (R * p);
  d[0] = o[0];
#pragma unroll
  for (int k = 1; k < R; ++k) d[s * k] = t1 ? cmulf(o[k], tw[t1 * k]) : o[k];
}

// any other (prime) radix: direct r-point DFT, inputs re-read from shared memory
template <class V>
__device__ __forceinline__ void butterfly_any(const V* __restrict__ src, V* __restrict__ dst, const V* __restrict__ tw,
                                              int r, int p, int q, int s, int m, int Z) {
  const int zr = Z / r;
  for (int k = 0; k < r; ++k) {
    V acc = src[q + s * p];
    int e = 0;
    for (int j = 1; j < r; ++j) {
      e += k;
      if (e >= r) e -= r;
      acc = caddf(acc, cmulf(src[q + s * (p + j * m)], tw[e * zr]));
    }
    dst[q + s * (r * p + k)] = cmulf(acc, tw[p * k * s]);
  }
}

// all stages on the rows of the tile (`tw` in shared memory, already conjugated for the inverse); returns the
// buffer that holds the result (natural order).  Ends with a __syncthreads().
template <bool INV, class V>
__device__ __forceinline__ V* fft_rows(V* a, V* b, const V* tw, const Rows& g) {
  int s = 1, n = g.Z;
  for (int st = 0; st < g.rad.n; ++st) {
    const int r = g.rad.r[st], m = n / r, nb = g.Z / r;
    const unsigned magic = g.smagic[st];
    // lane = row (rows vary fastest: odd row stride -> distinct banks), the butterflies of a row are strided
    // over the rest of the CTA
    const int row = threadIdx.x & (g.rows - 1);
    const V* x = a + row * g.zp;
    V* y = b + row * g.zp;
    for (int bf = threadIdx.x >> g.rshift; bf < nb; bf += ZT >> g.rshift) {
      const int p = fdiv(bf, magic), q = bf - p * s;
      switch (r) {
        case 8: butterfly<8, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 2: butterfly<2, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 3: butterfly<3, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 4: butterfly<4, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 5: butterfly<5, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 7: butterfly<7, INV>(x, y, tw, p, q, s, m, g.Z); break;
        default: butterfly_any(x, y, tw, r, p, q, s, m, g.Z); break;
      }
    }
    __syncthreads();
    V* t = a;
    a = b;
    b = t;
    n = m;
    s *= r;
  }
  return a;
}

// ---------------------------------------------------------------- host side
// table exp(-2 pi i t / Z), t < Z, on the current device (cached per device, length and precision)
template <class V>
inline int twiddles(int Z, const V** out) {
  static std::mutex mu;
  static std::map<std::tuple<int, int>, V*> cache;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(std::make_tuple(dev, Z));
  if (it == cache.end()) {
    std::vector<V> h(Z);
    const double pi = 3.14159265358979323846;
    for (int t = 0; t < Z; ++t) {
      const double a = -2.0 * pi * (double)t / (double)Z;
      h[t].x = (typename RealOf<V>::type)cos(a);
      h[t].y = (typename RealOf<V>::type)sin(a);
    }
    V* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)Z * sizeof(V)));
    CUDA_TRY(cudaMemcpy(d, h.data(), (size_t)Z * sizeof(V), cudaMemcpyHostToDevice));
    it = cache.emplace(std::make_tuple(dev, Z), d).first;
  }
  *out = it->second;
  return B200_OK;
}

inline int factorise(int Z, Radices* rad) {
  rad->n = 0;
  auto push = [&](int r) {
    if (rad->n < MAXRAD) rad->r[rad->n] = r;
    ++rad->n;
  };
  while (Z % 8 == 0) { push(8); Z /= 8; }
  while (Z % 4 == 0) { push(4); Z /= 4; }
  for (int p : {2, 3, 5, 7})
    while (Z % p == 0) { push(p); Z /= p; }
  for (int p = 11; Z > 1; p += 2)
    while (Z % p == 0) { push(p); Z /= p; }
  if (rad->n > MAXRAD) {
    b200_set_error("FFT length has too many prime factors");
    return B200_EINVAL;
  }
  return B200_OK;
}

// after Z and rows are set: radices, row shift, per-stage division magics
inline int prepare(Rows* g) {
  B200_TRY(factorise(g->Z, &g->rad));
  g->rshift = 0;
  while ((1 << g->rshift) < g->rows) ++g->rshift;
  if ((1 << g->rshift) != g->rows || g->rows > ZT) {
    b200_set_error("internal: FFT tile rows must be a power of two");
    return B200_EINVAL;
  }
  g->zmagic = fdiv_magic((unsigned)g->Z);
  unsigned s = 1;
  for (int st = 0; st < g->rad.n; ++st) {
    g->smagic[st] = fdiv_magic(s);
    s *= (unsigned)g->rad.r[st];
  }
  return B200_OK;
}

}  // namespace fftsm
