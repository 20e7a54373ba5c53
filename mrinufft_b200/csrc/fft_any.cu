// fft_any.cu -- in-place multi-dimensional FFT of the oversampled grids for ANY axis lengths, complex64 and
// complex128: an own replacement of the cuFFT C2C / Z2Z executions that serve the plans where the
// zero-padding-aware power-of-two passes of fft_pruned.cu do not apply (grids with factors 3, 5, 7, ...; 1-D;
// complex128 plans).  Selected with option key 2 = 4.  Correct on every grid of the test suite, but measured at
// 1.5 - 1.8x cuFFT's time (three or four stages through shared memory per axis against cuFFT's two in
// registers), so cuFFT remains the default for those grids; the stacked operator's z transform, where the
// fusion with its neighbours pays, uses the same shared-memory FFT (stack_fftz.cu).
//
// One pass per axis.  The T grids (n0, n1, n2), C order, are seen as [outer][L][S] for the axis of length L
// (S = product of the faster dimensions).  A CTA takes a tile of `rows` lines of that axis into shared
// memory -- for the fastest axis (S = 1) `rows` consecutive lines, read along the line; for a strided axis
// `rows` neighbouring columns, so that every global access is a `rows * sizeof(element)`-byte segment --, runs
// the Stockham autosort FFT of fft_smem.cuh on them and writes them back in natural order.  HBM-bound: one
// read and one write of the grid per axis.
#include "fft_smem.cuh"

namespace {

using namespace fftsm;

template <class V>
struct AxisArgs {
  V* data;
  const V* tw;
  long long outer;   // lines' slow index count (T x slower dimensions)
  long long S;       // element stride of the axis = number of columns per outer index
  long long tiles_per_outer;
  Rows r;
};

template <class V, bool INV>
__global__ void __launch_bounds__(ZT) k_fft_axis(const AxisArgs<V> g) {
  extern __shared__ __align__(16) unsigned char fsm_raw[];
  V* tw = reinterpret_cast<V*>(fsm_raw);
  const int L = g.r.Z, rows = g.r.rows, zp = g.r.zp;
  V* bufa = tw + L;
  V* bufb = bufa + rows * zp;
  for (int t = threadIdx.x; t < L; t += ZT) {
    const V w = g.tw[t];
    tw[t] = INV ? mkc<V>(w.x, -w.y) : w;
  }
  const long long tile = blockIdx.x;
  if (g.S == 1) {
    // fastest axis: `rows` consecutive lines, contiguous in memory
    const long long o0 = tile * rows;
    const int nrow = (int)min((long long)rows, g.outer - o0);
    V* base = g.data + o0 * L;
    for (int i = threadIdx.x; i < nrow * L; i += ZT) {
      const int row = fdiv(i, g.r.zmagic), k = i - row * L;
      bufa[row * zp + k] = base[i];
    }
    if (nrow < rows)
      for (int i = threadIdx.x + nrow * L; i < rows * L; i += ZT) {
        const int row = fdiv(i, g.r.zmagic), k = i - row * L;
        bufa[row * zp + k] = mkc<V>(0, 0);
      }
    __syncthreads();
    const V* res = fft_rows<INV>(bufa, bufb, tw, g.r);
    for (int i = threadIdx.x; i < nrow * L; i += ZT) {
      const int row = fdiv(i, g.r.zmagic), k = i - row * L;
      base[i] = res[row * zp + k];
    }
  } else {
    // strided axis: `rows` neighbouring columns of one outer index
    const long long o = tile / g.tiles_per_outer, c0 = (tile - o * g.tiles_per_outer) * rows;
    const int ncol = (int)min((long long)rows, g.S - c0);
    V* base = g.data + o * L * g.S + c0;
    for (int i = threadIdx.x; i < rows * L; i += ZT) {
      const int k = i >> g.r.rshift, c = i & (rows - 1);
      bufa[c * zp + k] = c < ncol ? base[(long long)k * g.S + c] : mkc<V>(0, 0);
    }
    __syncthreads();
    const V* res = fft_rows<INV>(bufa, bufb, tw, g.r);
    for (int i = threadIdx.x; i < rows * L; i += ZT) {
      const int k = i >> g.r.rshift, c = i & (rows - 1);
      if (c < ncol) base[(long long)k * g.S + c] = res[c * zp + k];
    }
  }
}

template <class V>
int axis_pass(V* data, long long outer, int L, long long S, int sign, cudaStream_t st) {
  if (L == 1) return B200_OK;
  AxisArgs<V> g{};
  g.data = data;
  g.outer = outer;
  g.S = S;
  g.r.Z = L;
  g.r.zp = L | 1;
  B200_TRY(twiddles<V>(L, &g.tw));
  auto bytes = [&](int r) { return ((size_t)2 * r * g.r.zp + L) * sizeof(V); };
  // 16 lines per tile when two CTAs still fit an SM, fewer for long lines
  int rows = 16;
  while (rows > 4 && bytes(rows) > 110 * 1024) rows >>= 1;
  while (rows > 1 && bytes(rows) > 216 * 1024) rows >>= 1;
  if (bytes(rows) > 216 * 1024) {
    b200_set_error("FFT axis of length %d does not fit one CTA's shared memory", L);
    return B200_EINVAL;
  }
  g.r.rows = rows;
  B200_TRY(prepare(&g.r));
  long long tiles;
  if (S == 1) {
    g.tiles_per_outer = 1;
    tiles = (outer + rows - 1) / rows;
  } else {
    g.tiles_per_outer = (S + rows - 1) / rows;
    tiles = outer * g.tiles_per_outer;
  }
  if (tiles > 0x7fffffffLL) {
    b200_set_error("FFT grid too large for one launch");
    return B200_EINVAL;
  }
  const size_t smem = bytes(rows);
  if (sign < 0) {
    auto k = k_fft_axis<V, false>;
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)tiles, ZT, smem, st>>>(g);
  } else {
    auto k = k_fft_axis<V, true>;
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)tiles, ZT, smem, st>>>(g);
  }
  CHECK_LAUNCH();
  return B200_OK;
}

template <class V>
int fft_grids(V* fw, int T, const Geom& g, int sign, cudaStream_t st) {
  // fastest axis first (its lines are contiguous), then the strided ones
  long long S = 1;
  for (int a = g.dim - 1; a >= 0; --a) {
    const int L = g.nf[a];
    const long long outer = (long long)T * (g.nftot / ((long long)L * S));
    B200_TRY(axis_pass<V>(fw, outer, L, S, sign, st));
    S *= L;
  }
  return B200_OK;
}

}  // namespace

// Unnormalised FFT (sign < 0: exp(-i ...), else exp(+i ...)) of T grids of the plan's oversampled size, in place.
int fft_any_c64(float2* fw, int T, const Geom& g, int sign, cudaStream_t st) { return fft_grids<float2>(fw, T, g, sign, st); }
int fft_any_c128(double2* fw, int T, const Geom& g, int sign, cudaStream_t st) { return fft_grids<double2>(fw, T, g, sign, st); }

// Plan-less entry point: T contiguous C-order arrays of `dim` axes n[0..dim), complex64 (dbl = 0) or complex128,
// transformed in place along every axis (unnormalised; sign < 0: exp(-i ...)).
extern "C" int b200_fft_c2c(void* data, int T, int dim, const int64_t* n, int sign, int dbl, void* stream) {
  if (!data || !n || T < 1 || dim < 1 || dim > 3) {
    b200_set_error("b200_fft_c2c: bad arguments (T=%d, dim=%d)", T, dim);
    return B200_EINVAL;
  }
  Geom g{};
  g.dim = dim;
  g.nftot = 1;
  for (int a = 0; a < 3; ++a) {
    g.nf[a] = a < dim ? (int)n[a] : 1;
    if (g.nf[a] < 1 || (a < dim && n[a] > (1 << 20))) {
      b200_set_error("b200_fft_c2c: axis %d has length %lld", a, (long long)n[a]);
      return B200_EINVAL;
    }
    g.nftot *= g.nf[a];
  }
  return dbl ? fft_grids<double2>((double2*)data, T, g, sign, (cudaStream_t)stream)
             : fft_grids<float2>((float2*)data, T, g, sign, (cudaStream_t)stream);
}
