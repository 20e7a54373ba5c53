// fft_pruned.cu -- zero-padding-aware FFT passes on the oversampled grid, fused with the
// pad / crop + deapodisation (+ sensitivity-map) passes (kernels K4a / K4b).
//
// The oversampled grid of a type-2 transform is the image zero-padded from N to nf = 2N modes per
// axis, and a type-1 transform keeps only N of the nf output modes per axis.  A library 3-D FFT
// moves the full grid three times (~6.4 GB per 512^3 coil) after/before a separate pad/crop pass.
// Here each axis is one pass that touches only what is non-zero / needed:
//
//   type 2:  x-pass  image (N0 N1 rows of N2)  -> grid rows  (N0 N1 rows of nf2)    [pad+deapod+smaps fused]
//            y-pass  N0 planes, N1 of nf1 rows non-zero -> all nf1 rows               [in place]
//            z-pass  N0 of nf0 planes non-zero -> all nf0 planes                      [in place]
//   type 1:  the mirror image: z-pass keeps N0 planes, y-pass keeps N1 rows of those, the x-pass
//            reads N0 N1 rows and writes the image                  [crop+deapod+conj(smaps)+coil sum fused]
//
// 2.9 GB instead of 7.6 GB per coil at 256^3 -> 512^3.  Every pass is a two-step (R1 x R2)
// Cooley-Tukey on a tile of 16 columns (or rows): step A loads R1 <= 32 points per thread straight
// from global memory (coalesced 128-byte segments, up to 32 independent loads in flight per
// thread), runs an in-register FFT (fft_reg.cuh, packed FFMA2 arithmetic), applies the twiddles and
// parks the result in shared memory; step B reads R2 points per thread, runs the second in-register
// FFT and stores straight to global memory.  One shared-memory round trip per pass, no bank
// conflicts, no global transposes.
//
// Power-of-two grid sizes 32..1024 per axis (sigma = 2 of a power-of-two image); other sizes use
// cuFFT + k_pad / k_crop (api.cu).  Replaces finufft's FFTW call + deconvolve step reached through
// `Plan.execute` / `Plan.execute_adjoint` (src/mrinufft/operators/interfaces/finufft.py:69,76) and
// the smaps / coil-combine passes of src/mrinufft/operators/base.py:988-993, 1045-1051.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "device_utils.cuh"
#include "fft_reg.cuh"

namespace {

using fftreg::brev;
using fftreg::sfor;

constexpr int TX = 16;   // columns (strided passes) or rows (contiguous passes) per CTA
constexpr int FT = 256;  // threads per CTA
constexpr int TXR1 = 8;  // rows per CTA of the type-1 x-pass (two CTAs per SM at 128 registers)

template <int L> struct Split;
template <> struct Split<32>   { static constexpr int R1 = 8,  R2 = 4;  };
template <> struct Split<64>   { static constexpr int R1 = 8,  R2 = 8;  };
template <> struct Split<128>  { static constexpr int R1 = 16, R2 = 8;  };
template <> struct Split<256>  { static constexpr int R1 = 16, R2 = 16; };
template <> struct Split<512>  { static constexpr int R1 = 32, R2 = 16; };
template <> struct Split<1024> { static constexpr int R1 = 32, R2 = 32; };

// kept index set of an axis of length L: n < np || n >= L - nm   (all: np = L, nm = 0)
struct Keep {
  int np, nm;
};
__device__ __forceinline__ bool kept(int n, int L, Keep k) { return n < k.np || n >= L - k.nm; }
// j-th kept index in ascending order
__device__ __forceinline__ int kept_index(int j, int L, Keep k) { return j < k.np ? j : L - k.nm + (j - k.np); }

Keep keep_all(int L) { return Keep{L, 0}; }
// image index i (mode i - N/2) lives at fine index (i - N/2) mod L: N - N/2 non-negative modes, N/2 negative
Keep keep_modes(int N) { return Keep{N - N / 2, N / 2}; }

// ------------------------------------------------------------------------------ strided pass
struct StridedArgs {
  float2* base;            // coil 0 of the grid
  long long coil_stride;   // elements between coils
  long long stride_n;      // stride of the transformed axis
  long long outer_stride;  // stride of the outer (slower or faster non-contiguous) axis
  int outer_L;             // its length
  Keep outer_keep;         // which outer indices are processed (blockIdx.y enumerates them)
  Keep in, out;            // non-zero inputs / wanted outputs along the transformed axis
  const float* mul;        // nullable: real factor per grid point applied to the outputs (Toeplitz)
  // nullable: bit strings of the spreader's empty tiles (spread_rows.cu, k_mark_empty) -- tiles it
  // left unwritten because no point visits them; read as zeros.  Only for passes along axis 0.
  const uint32_t* empty;
  int empty_3d;            // 1: transformed axis = z, outer index = y; 0 (2-D): transformed axis = y
  int empty_out;           // 0: applies to the inputs (type 1), 1: to the outputs (type 2)
  int nyh, nbx;
  int tma_z;               // TMA variant: 1 = the transformed axis is dimension 2 of the tensor map (3-D z-pass),
                           // 0 = dimension 1 (3-D y-pass, 2-D pass)
  int lookahead;           // > 0: every CTA pulls the input lines of the tile `lookahead` CTAs behind it in launch
                           // order into L2 (prefetch.global.L2), so that tile's loads are L2 hits when it runs
};

// ---- TMA (cp.async.bulk.tensor) + mbarrier plumbing of the bulk-load variant of the strided pass
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// one box of the 4-D tensor (x, y, z, coil) -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
          "r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(bar))
      : "memory");
}

// EMODE 0: plain; 1: the spreader's empty tiles are not read (type 1, first pass); 2: tiles that no
// point visits are not written (type 2, last pass: the row interpolator never reads them)
// KIN / KOUT = 1: the non-zero inputs / wanted outputs are the modes of an N = L / 2 image (sigma = 2), a
// pattern known at compile time: inputs n1 < R1/4 or n1 >= 3 R1/4, outputs k2 < R2/4 or k2 >= 3 R2/4 --
// no per-element predicates.  2: all of them, no predicates either.  0: generic (run-time `Keep`).
// TMA = true: the tile is brought into shared memory with bulk tensor copies (cp.async.bulk.tensor, boxes of
// 16 columns x BZ = 32 grid rows / planes = 4 KB, issued by one thread, completion on an mbarrier) instead of R1
// global loads per thread; step A then works in place on the tile.  Boxes that hold only zero padding, or
// only tiles the spreader left unwritten, are not issued.
template <int L, int DIR, bool MUL, int EMODE, int KIN, int KOUT, bool TMA>
__global__ void __launch_bounds__(FT, L <= 512 ? 3 : 1)
k_fft_strided(StridedArgs A, const float2* __restrict__ tw, const __grid_constant__ CUtensorMap tmap) {
  constexpr bool EMPTY = EMODE != 0;
  constexpr int R1 = Split<L>::R1, R2 = Split<L>::R2;
  extern __shared__ __align__(128) float2 S[];  // [L][TX]
  const int lo = kept_index(blockIdx.y, A.outer_L, A.outer_keep);
  // with the Toeplitz multiply the coil index varies fastest over the grid of CTAs: the 32 CTAs that
  // need the same tile of the (coil-independent) factor run together and share it through L2
  const int bt = MUL ? blockIdx.x : blockIdx.z, bx = MUL ? blockIdx.z : blockIdx.x;
  const long long goff = (long long)lo * A.outer_stride + (long long)bx * TX;
  float2* g = A.base + (long long)bt * A.coil_stride + goff;
  // the bit string of this CTA's column of spreader tiles (16 columns = one tile width)
  __shared__ uint32_t ebits[EMPTY ? L / 32 : 1];
  // the same bits regrouped by thread: tmask[j] bit i = plane i * NJ + j, where a thread of step A (EMODE 1:
  // j = n2, i = n1, NJ = R2) or step B (EMODE 2: j = k1, i = k2, NJ = R1) owns one j -- one word per thread
  // and compile-time bit positions instead of a shared-memory load and a run-time shift per element
  constexpr int NJ = EMODE == 1 ? R2 : R1, NI = L / NJ;
  __shared__ uint32_t tmask[EMPTY ? NJ : 1];
  if (EMPTY) {
    constexpr int wpc = L / 32;  // 3-D only: one bit per grid plane
    const long long col = (long long)(lo >> 1) * A.nbx + bx;
    if (threadIdx.x < wpc) ebits[threadIdx.x] = __ldg(A.empty + col * wpc + threadIdx.x);
    if (threadIdx.x >= 32 && threadIdx.x < 32 + NJ) {
      const int j = threadIdx.x - 32;
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int n = i * NJ + j;
        m |= ((__ldg(A.empty + col * wpc + (n >> 5)) >> (n & 31)) & 1u) << i;
      }
      tmask[j] = m;
    }
    __syncthreads();
    // a column without any visited tile (outside the trajectory's support): the transform of zeros
    bool all_empty = true;
    for (int w = 0; w < wpc; ++w) all_empty = all_empty && ebits[w] == 0xffffffffu;
    if (all_empty && EMODE == 2) return;  // nobody will read this column
    if (all_empty) {
      for (int item = threadIdx.x; item < L * TX; item += FT) {
        const int k = item / TX, tx = item % TX;
        if (KOUT == 2 || kept(k, L, A.out)) __stcs(g + (long long)k * A.stride_n + tx, make_float2(0.f, 0.f));
      }
      return;
    }
  }
  // Look-ahead prefetch: the strided passes are bound by the latency of their (independent, but per CTA
  // serialised) loads; the tile of a CTA that starts a wave later is requested now, without holding
  // registers or warps for it.  One 128-byte line per thread and prefetched row.
  if (!MUL && A.lookahead > 0) {
    const long long nx = gridDim.x, ny = gridDim.y, nz = gridDim.z;
    const long long id = blockIdx.x + nx * (blockIdx.y + ny * (long long)blockIdx.z) + A.lookahead;
    if (id < nx * ny * nz) {
      const int pbx = (int)(id % nx), pby = (int)((id / nx) % ny), pbt = (int)(id / (nx * ny));
      const int plo = kept_index(pby, A.outer_L, A.outer_keep);
      const char* pg = reinterpret_cast<const char*>(A.base + (long long)pbt * A.coil_stride +
                                                     (long long)plo * A.outer_stride + (long long)pbx * TX);
      for (int n = threadIdx.x; n < L; n += FT) {
        bool live = KIN == 1 ? (n < L / 4 || n >= 3 * L / 4) : (KIN == 2 || kept(n, L, A.in));
        if (EMODE == 1 && live) {
          const long long col = (long long)(plo >> 1) * A.nbx + pbx;
          live = !((__ldg(A.empty + col * (L / 32) + (n >> 5)) >> (n & 31)) & 1u);
        }
        if (live) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(pg + (long long)n * A.stride_n * (long long)sizeof(float2)));
      }
    }
  }
  if constexpr (TMA) {
    constexpr int BZ = L / 4 < 32 ? L / 4 : 32, NB = L / BZ;  // rows / planes per box, boxes per tile
    static_assert(EMODE != 1 || BZ == 32, "one word of tile flags per box");
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned need = 0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        bool want = KIN == 1 ? (b * BZ < L / 4 || b * BZ >= 3 * L / 4)
                             : (KIN == 2 || b * BZ < A.in.np || (b + 1) * BZ > L - A.in.nm);
        if (EMODE == 1) want = want && ebits[b] != 0xffffffffu;
        need |= want ? (1u << b) : 0u;
      }
      mbar_arrive_expect_tx(&bar, (unsigned)__popc(need) * (unsigned)(BZ * TX * sizeof(float2)));
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if ((need >> b) & 1u)
          tma_load_4d(S + b * BZ * TX, &tmap, bx * TX, A.tma_z ? lo : b * BZ, A.tma_z ? b * BZ : lo, bt, &bar);
    }
    mbar_wait(&bar, 0);
  }
  // step A: R1-point FFTs over n1 (n = n1 R2 + n2), twiddle W_L^(n2 k1)
  for (int item = threadIdx.x; item < R2 * TX; item += FT) {
    const int n2 = item / TX, tx = item % TX;
    float2 a[R1];
    const uint32_t im = EMODE == 1 ? tmask[EMODE == 1 ? n2 : 0] : 0u;
    const float2* gin = TMA ? S + n2 * TX + tx : g + (long long)n2 * A.stride_n + tx;
    // (32-bit element strides: L stride_n < 2^31 for every grid the plan accepts, so an address is one
    // IMAD.WIDE of a compile-time multiple of the stride)
    const unsigned sR2 = TMA ? (unsigned)(TX * R2) : (unsigned)(A.stride_n * R2);
    sfor<0, R1>([&](auto I) {
      constexpr int n1 = decltype(I)::value;
      constexpr bool static_zero = KIN == 1 && n1 >= R1 / 4 && n1 < 3 * R1 / 4;
      if constexpr (!static_zero) {  // (the zero middle of a sigma = 2 pad is never materialised)
        const int n = n1 * R2 + n2;
        bool live = KIN != 0 ? true : kept(n, L, A.in);
        if (EMODE == 1) live = live && !((im >> n1) & 1u);  // the tile of grid plane n (3-D only)
        a[n1] = live ? gin[(unsigned)n1 * sR2] : make_float2(0.f, 0.f);
      }
    });
    if constexpr (KIN == 1) fftreg::fft_zero_middle<R1, DIR>(a);
    else fftreg::fft<R1, DIR>(a);
    sfor<0, R1>([&](auto I) {
      constexpr int k1 = decltype(I)::value;
      float2 v = a[brev(k1, R1)];
      if (k1 > 0) {
        float2 w = __ldg(tw + n2 * k1);
        if (DIR < 0) w.y = -w.y;
        v = fftreg::cmul(v, w);
      }
      S[(k1 * R2 + n2) * TX + tx] = v;
    });
  }
  __syncthreads();
  // step B: R2-point FFTs over n2, output k = k1 + R1 k2
  for (int item = threadIdx.x; item < R1 * TX; item += FT) {
    const int k1 = item / TX, tx = item % TX;
    float2 b[R2];
    sfor<0, R2>([&](auto I) {
      constexpr int n2 = decltype(I)::value;
      b[n2] = S[(k1 * R2 + n2) * TX + tx];
    });
    // Toeplitz factor of this thread's outputs: requested before the FFT so that the loads are in
    // flight while it runs
    float mf[MUL ? R2 : 1];
    const uint32_t om = EMODE == 2 ? tmask[EMODE == 2 ? k1 : 0] : 0u;
    float2* gout = g + (long long)k1 * A.stride_n + tx;
    const unsigned sR1 = (unsigned)(A.stride_n * R1);
    if (MUL) {
      sfor<0, R2>([&](auto I) {
        constexpr int k2 = decltype(I)::value;
        const int k = k1 + R1 * k2;
        mf[MUL ? k2 : 0] = kept(k, L, A.out) ? __ldg(A.mul + goff + (long long)k * A.stride_n + tx) : 0.f;
      });
    }
    fftreg::fft<R2, DIR>(b);
    sfor<0, R2>([&](auto I) {
      constexpr int k2 = decltype(I)::value;
      constexpr bool static_drop = KOUT == 1 && k2 >= R2 / 4 && k2 < 3 * R2 / 4;
      if constexpr (!static_drop) {
        const int k = k1 + R1 * k2;
        bool wanted = KOUT != 0 ? true : kept(k, L, A.out);
        if (EMODE == 2) wanted = wanted && !((om >> k2) & 1u);
        if (wanted) {
          float2 v = b[brev(k2, R2)];
          if (MUL) v = cscale(v, mf[MUL ? k2 : 0]);
          __stcs(gout + (unsigned)k2 * sR1, v);
        }
      }
    });
  }
}

// ------------------------------------------------------------------------------ contiguous passes
struct RowArgs {
  Geom g;
  const float2* img_in;   // type 2: image(s)
  float2* img_out;        // type 1: image(s)
  const float2* smaps;    // nullable
  float2* fw;
  const float* d_slow0;   // deapodisation vectors: slowest axis, middle axis (3-D only), fastest axis
  const float* d_slow1;
  const float* d_fast;
  int T;
  int conj_smaps;
  int accumulate;
  float scale;
};

struct RowInfo {
  long long fw_off;   // offset of the row in one coil's grid
  long long img_off;  // offset of the row in one coil's image
  float dsl;          // product of the slow-axis deapodisation factors
};

// r-th image row -> where it sits in the grid
__device__ __forceinline__ RowInfo row_info(const RowArgs& A, int r) {
  const Geom& g = A.g;
  RowInfo ri;
  if (g.dim == 3) {
    const int i0 = r / g.N[1], i1 = r % g.N[1];
    ri.fw_off = ((long long)mode_to_fine(i0, g.N[0], g.nf[0]) * g.nf[1] + mode_to_fine(i1, g.N[1], g.nf[1])) * g.nf[2];
    ri.img_off = (long long)r * g.N[2];
    ri.dsl = A.d_slow0[i0] * A.d_slow1[i1];
  } else {
    ri.fw_off = (long long)mode_to_fine(r, g.N[0], g.nf[0]) * g.nf[1];
    ri.img_off = (long long)r * g.N[1];
    ri.dsl = A.d_slow0[r];
  }
  return ri;
}

// type 2 x-pass: image row -> grid row.  grid (ceil(nrows / TX) * T): the coil index varies fastest, so that
// the T CTAs that expand the same 16 image rows run together and share them through L2 (with the coil as the
// slowest grid index the 134 MB image of cfg-C was re-read from HBM once per coil: 4.2 of 17 GB)
// HALF: Nx == L / 2 (sigma = 2, Nx even): the non-zero inputs of a thread are its first and last
// R1 / 4 points, at offsets known at compile time -- no per-element index arithmetic or predicates
// (the generic path spends 70 % of its instructions there, ncu).
template <int L, int DIR, bool HALF>
__global__ void __launch_bounds__(FT, (HALF && L <= 512) ? 3 : 1)
k_fft_rows_t2(RowArgs A, int nrows, const float2* __restrict__ tw) {
  constexpr int R1 = Split<L>::R1, R2 = Split<L>::R2, RS = R1 * (R2 + 1);
  extern __shared__ float2 S[];  // [TX][R1][R2 + 1]
  const int Nx = A.g.N[A.g.dim - 1];
  const int t = blockIdx.x % A.T, tile = blockIdx.x / A.T;
  const float2* img = A.smaps ? A.img_in : A.img_in + (long long)t * A.g.Ntot;
  const float2* sm = A.smaps ? A.smaps + (long long)t * A.g.Ntot : nullptr;
  for (int item = threadIdx.x; item < R2 * TX; item += FT) {
    const int row = item / R2, n2 = item % R2;
    const int r = tile * TX + row;
    float2 a[R1];
    if (HALF && r < nrows) {
      const RowInfo ri = row_info(A, r);
      constexpr int Q = R1 / 4;  // non-zero points per end
      // fine index n = n1 R2 + n2: modes n >= 0 (n1 < Q) sit at image index n + Nx/2, modes
      // n - L < 0 (n1 >= 3Q) at image index n - L + Nx/2 = (n1 - 3Q) R2 + n2
      const float2* ip = img + ri.img_off + n2;
      const float* dp = A.d_fast + n2;
      constexpr int hi_off = L / 4;  // Nx / 2
      float2 v[2 * Q];
      float dw[2 * Q];
      sfor<0, Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        v[q] = __ldg(ip + hi_off + q * R2);
        v[Q + q] = __ldg(ip + q * R2);
      });
      sfor<0, Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        dw[q] = __ldg(dp + hi_off + q * R2);
        dw[Q + q] = __ldg(dp + q * R2);
      });
      if (sm) {
        const float2* sp = sm + ri.img_off + n2;
        float2 sv[2 * Q];
        sfor<0, Q>([&](auto I) {
          constexpr int q = decltype(I)::value;
          sv[q] = __ldg(sp + hi_off + q * R2);
          sv[Q + q] = __ldg(sp + q * R2);
        });
        sfor<0, 2 * Q>([&](auto I) {
          constexpr int q = decltype(I)::value;
          const float2 x = cscale(v[q], ri.dsl * dw[q]);
          v[q] = A.conj_smaps ? cmul_conj(x, sv[q]) : cmul(x, sv[q]);
        });
      } else {
        sfor<0, 2 * Q>([&](auto I) {
          constexpr int q = decltype(I)::value;
          v[q] = cscale(v[q], ri.dsl * dw[q]);
        });
      }
      sfor<0, R1>([&](auto I) {
        constexpr int n1 = decltype(I)::value;
        a[n1] = n1 < Q ? v[n1 < Q ? n1 : 0]
                       : (n1 >= 3 * Q ? v[n1 >= 3 * Q ? Q + (n1 - 3 * Q) : 0] : make_float2(0.f, 0.f));
      });
    } else if (r < nrows) {
      const RowInfo ri = row_info(A, r);
      // predicated loads, issued in two batches of R1/2 taps so that all real loads of a batch
      // (image, deapodisation factor, sensitivity map) are in flight together
      sfor<0, 2>([&](auto HB) {
        constexpr int h0 = decltype(HB)::value * (R1 / 2);
        float dw[R1 / 2];
        float2 sv[R1 / 2];
        sfor<0, R1 / 2>([&](auto I) {
          constexpr int q = decltype(I)::value;
          const int ix = fine_to_mode((h0 + q) * R2 + n2, Nx, L);
          a[h0 + q] = ix >= 0 ? __ldg(img + ri.img_off + ix) : make_float2(0.f, 0.f);
        });
        sfor<0, R1 / 2>([&](auto I) {
          constexpr int q = decltype(I)::value;
          const int ix = fine_to_mode((h0 + q) * R2 + n2, Nx, L);
          dw[q] = ix >= 0 ? __ldg(A.d_fast + ix) : 0.f;
        });
        if (sm) {
          sfor<0, R1 / 2>([&](auto I) {
            constexpr int q = decltype(I)::value;
            const int ix = fine_to_mode((h0 + q) * R2 + n2, Nx, L);
            sv[q] = ix >= 0 ? __ldg(sm + ri.img_off + ix) : make_float2(0.f, 0.f);
          });
          sfor<0, R1 / 2>([&](auto I) {
            constexpr int q = decltype(I)::value;
            const float2 v = cscale(a[h0 + q], ri.dsl * dw[q]);
            a[h0 + q] = A.conj_smaps ? cmul_conj(v, sv[q]) : cmul(v, sv[q]);
          });
        } else {
          sfor<0, R1 / 2>([&](auto I) {
            constexpr int q = decltype(I)::value;
            a[h0 + q] = cscale(a[h0 + q], ri.dsl * dw[q]);
          });
        }
      });
    } else {
      sfor<0, R1>([&](auto I) { a[decltype(I)::value] = make_float2(0.f, 0.f); });
    }
    if constexpr (HALF) fftreg::fft_zero_middle<R1, DIR>(a);
    else fftreg::fft<R1, DIR>(a);
    sfor<0, R1>([&](auto I) {
      constexpr int k1 = decltype(I)::value;
      float2 v = a[brev(k1, R1)];
      if (k1 > 0) {
        float2 w = __ldg(tw + n2 * k1);
        if (DIR < 0) w.y = -w.y;
        v = fftreg::cmul(v, w);
      }
      S[row * RS + k1 * (R2 + 1) + n2] = v;
    });
  }
  __syncthreads();
  for (int item = threadIdx.x; item < R1 * TX; item += FT) {
    const int row = item / R1, k1 = item % R1;
    const int r = tile * TX + row;
    if (r >= nrows) continue;
    float2 b[R2];
    sfor<0, R2>([&](auto I) {
      constexpr int n2 = decltype(I)::value;
      b[n2] = S[row * RS + k1 * (R2 + 1) + n2];
    });
    fftreg::fft<R2, DIR>(b);
    float2* out = A.fw + (long long)t * A.g.nftot + row_info(A, r).fw_off;
    sfor<0, R2>([&](auto I) {
      constexpr int k2 = decltype(I)::value;
      __stcs(out + k1 + R1 * k2, b[brev(k2, R2)]);
    });
  }
}

// type 1 x-pass: grid row -> image row, crop + deapodise (+ conj(smaps) multiply + coil sum).
// grid (ceil(nrows / TX), 1, smaps ? 1 : T); with smaps the CTA loops over the T coils and keeps
// the coil sum of its outputs in registers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// With smaps the CTA loops over the T coils; the raw rows of coil t+1 stream into a second
// shared-memory buffer with cp.async while coil t is transformed, and the sensitivity-map values of
// coil t are requested before the barriers of the iteration, so no DRAM latency is exposed.
// HALF: Nx == L / 2 (sigma = 2): a thread keeps outputs k2 < R2/4 and k2 >= 3 R2/4 only, so the coil
// accumulators and map values are compact arrays of R2 / 2.
template <int L, int DIR, bool HALF>
__global__ void __launch_bounds__(Split<L>::R1 * TXR1, (L <= 512 ? 2 : 1))
k_fft_rows_t1(RowArgs A, int nrows, const float2* __restrict__ tw) {
  constexpr int R1 = Split<L>::R1, R2 = Split<L>::R2, RS = R1 * (R2 + 1);
  constexpr int NT = R1 * TXR1;
  constexpr int KQ = HALF ? R2 / 2 : R2;  // outputs a thread may keep
  extern __shared__ __align__(16) float2 S[];   // [TXR1][R1][R2 + 1] | raw[2][TXR1][L]
  float2* raw = S + TXR1 * RS;
  const int Nx = A.g.N[A.g.dim - 1];
  const bool sense = A.smaps != nullptr;
  const int t_begin = sense ? 0 : blockIdx.z;
  const int t_end = sense ? A.T : blockIdx.z + 1;
  // step-A role of this thread (threads >= R2 * TXR1 idle in step A), step-B role (all threads)
  const int rowA = threadIdx.x / R2, n2A = threadIdx.x % R2;
  const bool doA = threadIdx.x < R2 * TXR1;
  const int rowB = threadIdx.x / R1, k1 = threadIdx.x % R1;
  const int rB = blockIdx.x * TXR1 + rowB;
  const bool doB = rB < nrows;
  RowInfo riB{};
  if (doB) riB = row_info(A, rB);
  float2 acc[KQ];
#pragma unroll
  for (int q = 0; q < KQ; ++q) acc[q] = make_float2(0.f, 0.f);
  // q-th kept output of this thread: k = k1 + R1 * k2(q)
  auto k2_of = [](int q) { return HALF ? (q < R2 / 4 ? q : q + R2 / 2) : q; };

  // cp.async roles: 16-byte chunks, TXR1 * L / 2 of them per coil
  constexpr int CHUNKS = TXR1 * L / 2;
  auto prefetch = [&](int t, int buf) {
    const float2* fwt = A.fw + (long long)t * A.g.nftot;
    for (int c = threadIdx.x; c < CHUNKS; c += NT) {
      const int row = c / (L / 2), e = (c % (L / 2)) * 2;
      const int r = blockIdx.x * TXR1 + row;
      if (r < nrows) cp_async16(raw + (buf * TXR1 + row) * L + e, fwt + row_info(A, r).fw_off + e);
    }
    cp_async_commit();
  };
  prefetch(t_begin, 0);
  for (int t = t_begin; t < t_end; ++t) {
    const int buf = (t - t_begin) & 1;
    if (t + 1 < t_end) {
      prefetch(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    // this coil's sensitivity-map values: issued now, consumed after both barriers
    float2 sv[KQ];
    if (sense && doB) {
      const float2* sm = A.smaps + (long long)t * A.g.Ntot + riB.img_off;
      sfor<0, KQ>([&](auto I) {
        constexpr int q = decltype(I)::value;
        const int ix = fine_to_mode(k1 + R1 * k2_of(q), Nx, L);
        sv[q] = ix >= 0 ? __ldg(sm + ix) : make_float2(0.f, 0.f);
      });
    }
    __syncthreads();
    if (doA) {
      float2 a[R1];
      const float2* in = raw + (buf * TXR1 + rowA) * L;
      const bool live = blockIdx.x * TXR1 + rowA < nrows;
      sfor<0, R1>([&](auto I) {
        constexpr int n1 = decltype(I)::value;
        a[n1] = live ? in[n1 * R2 + n2A] : make_float2(0.f, 0.f);
      });
      fftreg::fft<R1, DIR>(a);
      sfor<0, R1>([&](auto I) {
        constexpr int kk = decltype(I)::value;
        float2 v = a[brev(kk, R1)];
        if (kk > 0) {
          float2 w = __ldg(tw + n2A * kk);
          if (DIR < 0) w.y = -w.y;
          v = fftreg::cmul(v, w);
        }
        S[rowA * RS + kk * (R2 + 1) + n2A] = v;
      });
    }
    __syncthreads();
    if (doB) {
      float2 b[R2];
      sfor<0, R2>([&](auto I) {
        constexpr int n2 = decltype(I)::value;
        b[n2] = S[rowB * RS + k1 * (R2 + 1) + n2];
      });
      fftreg::fft<R2, DIR>(b);
      sfor<0, KQ>([&](auto I) {
        constexpr int q = decltype(I)::value;
        constexpr int k2 = HALF ? (q < R2 / 4 ? q : q + R2 / 2) : q;
        const float2 v = b[brev(k2, R2)];
        if (sense) {
          const float2 pr = A.conj_smaps ? cmul(v, sv[q]) : cmul_conj(v, sv[q]);
          acc[q].x += pr.x;
          acc[q].y += pr.y;
        } else {
          acc[q] = v;
        }
      });
    }
    // the next iteration's first __syncthreads orders these reads of S / raw[buf] before their reuse
  }
  // epilogue: deapodise, scale, (accumulate,) store
  if (!doB) return;
  float2* outimg = sense ? A.img_out : A.img_out + (long long)blockIdx.z * A.g.Ntot;
  sfor<0, KQ>([&](auto I) {
    constexpr int q = decltype(I)::value;
    const int ix = fine_to_mode(k1 + R1 * k2_of(q), Nx, L);
    if (ix >= 0) {
      float2 v = cscale(acc[q], riB.dsl * A.d_fast[ix] * A.scale);
      float2* o = outimg + riB.img_off + ix;
      if (A.accumulate) {
        const float2 old = *o;
        v.x += old.x;
        v.y += old.y;
      }
      *o = v;
    }
  });
}

// type 1 x-pass for L = 512, sigma = 2: three radix-8 steps (512 = 8 x 8 x 8) instead of 32 x 16.
// An 8-point register FFT needs 16 registers instead of 64, so the kernel runs at 64 registers and four
// CTAs (32 warps) per SM, all 64 threads of a row work in every step, and the inputs are read
// straight from global memory (8 coalesced loads per thread, 64 KB in flight per SM) -- the 32 x 16
// version was latency bound at 128 registers and 2.65 TB/s.
//   n = 64 n1 + 8 n2 + n3,  k = k1 + 8 k2 + 64 k3
//   step 1 (n2, n3): FFT over n1, twiddle W^(8 n2 k1)          -> S1[k1][8 n2 + n3]
//   step 2 (k1, n3): FFT over n2, twiddle W^(n3 (k1 + 8 k2))    -> S2[k2][9 k1 + n3]
//   step 3 (k1, k2): FFT over n3; kept outputs k3 in {6, 7, 0, 1} are image columns j, 64+j, 128+j,
//                    192+j with j = k1 + 8 k2 = the thread's index in its row: coalesced map reads
//                    and image writes.
template <int DIR>
__global__ void __launch_bounds__(256, 4)
k_fft_rows_t1_512(RowArgs A, int nrows, const float2* __restrict__ tw) {
  constexpr int L = 512, ROWS = 4, PS = 72;  // PS: padded slab stride (conflict-free exchanges)
  __shared__ float2 S1[ROWS][8 * PS];
  __shared__ float2 S2[ROWS][8 * PS];
  __shared__ float2 stw[L];
  for (int i = threadIdx.x; i < L; i += 256) {
    float2 w = __ldg(tw + i);
    if (DIR < 0) w.y = -w.y;
    stw[i] = w;
  }
  const int row = threadIdx.x >> 6, j = threadIdx.x & 63;
  const int hi = j >> 3, lo = j & 7;
  const int r = blockIdx.x * ROWS + row;
  const bool live = r < nrows;
  RowInfo ri{};
  if (live) ri = row_info(A, r);
  const bool sense = A.smaps != nullptr;
  const int t_begin = sense ? 0 : blockIdx.z;
  const int t_end = sense ? A.T : blockIdx.z + 1;
  float2 acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
  __syncthreads();
  for (int t = t_begin; t < t_end; ++t) {
    float2 a[8];
    {
      const float2* in = A.fw + (long long)t * A.g.nftot + ri.fw_off + j;
      sfor<0, 8>([&](auto I) {
        constexpr int n1 = decltype(I)::value;
        a[n1] = live ? __ldg(in + n1 * 64) : make_float2(0.f, 0.f);
      });
    }
    float2 sv[4];
    if (sense && live) {
      const float2* sm = A.smaps + (long long)t * A.g.Ntot + ri.img_off + j;
      sfor<0, 4>([&](auto I) { sv[decltype(I)::value] = __ldg(sm + 64 * decltype(I)::value); });
    }
    // step 1: thread (n2 = hi, n3 = lo)
    fftreg::fft<8, DIR>(a);
    sfor<0, 8>([&](auto I) {
      constexpr int k1 = decltype(I)::value;
      float2 v = a[brev(k1, 8)];
      if (k1 > 0) v = fftreg::cmul(v, stw[8 * hi * k1]);
      S1[row][k1 * PS + j] = v;
    });
    __syncthreads();
    // step 2: thread (k1 = hi, n3 = lo)
    sfor<0, 8>([&](auto I) {
      constexpr int n2 = decltype(I)::value;
      a[n2] = S1[row][hi * PS + n2 * 8 + lo];
    });
    fftreg::fft<8, DIR>(a);
    sfor<0, 8>([&](auto I) {
      constexpr int k2 = decltype(I)::value;
      float2 v = a[brev(k2, 8)];
      v = fftreg::cmul(v, stw[lo * (hi + 8 * k2)]);
      S2[row][k2 * PS + hi * 9 + lo] = v;
    });
    __syncthreads();
    // step 3: thread (k1 = lo, k2 = hi)
    sfor<0, 8>([&](auto I) {
      constexpr int n3 = decltype(I)::value;
      a[n3] = S2[row][hi * PS + lo * 9 + n3];
    });
    fftreg::fft<8, DIR>(a);
    sfor<0, 4>([&](auto I) {
      constexpr int q = decltype(I)::value;
      constexpr int k3 = (q + 6) & 7;  // q = 0..3 -> k3 = 6, 7, 0, 1 -> image column 64 q + j
      const float2 v = a[brev(k3, 8)];
      if (sense) {
        const float2 pr = A.conj_smaps ? cmul(v, sv[q]) : cmul_conj(v, sv[q]);
        acc[q].x += pr.x;
        acc[q].y += pr.y;
      } else {
        acc[q] = v;
      }
    });
    // the barriers of the next iteration order these reads of S2 (and the step-2 reads of S1) before
    // their buffers are written again
  }
  if (!live) return;
  float2* outimg = (sense ? A.img_out : A.img_out + (long long)blockIdx.z * A.g.Ntot) + ri.img_off + j;
  sfor<0, 4>([&](auto I) {
    constexpr int q = decltype(I)::value;
    float2 v = cscale(acc[q], ri.dsl * __ldg(A.d_fast + 64 * q + j) * A.scale);
    float2* o = outimg + 64 * q;
    if (A.accumulate) {
      const float2 old = *o;
      v.x += old.x;
      v.y += old.y;
    }
    *o = v;
  });
}

// ------------------------------------------------------------------------------ host side
int ensure_twiddles(b200_plan* p) {
  for (int a = 0; a < p->g.dim; ++a) {
    if (p->d_tw[a]) continue;
    const int L = p->g.nf[a];
    std::vector<float2> h(L);
    for (int j = 0; j < L; ++j) {
      const double ang = 2.0 * 3.14159265358979323846 * (double)j / (double)L;
      h[j] = make_float2((float)cos(ang), (float)sin(ang));
    }
    CUDA_TRY(cudaMalloc(&p->d_tw[a], (size_t)L * sizeof(float2)));
    CUDA_TRY(cudaMemcpy(p->d_tw[a], h.data(), (size_t)L * sizeof(float2), cudaMemcpyHostToDevice));
  }
  return B200_OK;
}

template <class K>
int set_smem(K kern, size_t bytes) {
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return B200_OK;
}

#define DISPATCH_L(L_, DIR_, CALL)                                   \
  do {                                                               \
    if (DIR_ < 0) {                                                  \
      switch (L_) {                                                  \
        case 32:   { constexpr int LL = 32,   DD = -1; CALL; } break; \
        case 64:   { constexpr int LL = 64,   DD = -1; CALL; } break; \
        case 128:  { constexpr int LL = 128,  DD = -1; CALL; } break; \
        case 256:  { constexpr int LL = 256,  DD = -1; CALL; } break; \
        case 512:  { constexpr int LL = 512,  DD = -1; CALL; } break; \
        default:   { constexpr int LL = 1024, DD = -1; CALL; } break; \
      }                                                              \
    } else {                                                         \
      switch (L_) {                                                  \
        case 32:   { constexpr int LL = 32,   DD = 1; CALL; } break;  \
        case 64:   { constexpr int LL = 64,   DD = 1; CALL; } break;  \
        case 128:  { constexpr int LL = 128,  DD = 1; CALL; } break;  \
        case 256:  { constexpr int LL = 256,  DD = 1; CALL; } break;  \
        case 512:  { constexpr int LL = 512,  DD = 1; CALL; } break;  \
        default:   { constexpr int LL = 1024, DD = 1; CALL; } break;  \
      }                                                              \
    }                                                                \
  } while (0)

// tensor maps of the TMA variant, cached per plan: [0] boxes along dimension 1, [1] along dimension 2
struct TmaMaps {
  CUtensorMap map[2];
  bool ok[2] = {false, false};
  const void* base = nullptr;
  int T = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The grid as a 4-D tensor (x, y, z, coil) of 8-byte elements; box = 16 columns x BZ rows (kind 0) or planes
// (kind 1).  Returns nullptr when the driver has no encoder (the caller then takes the plain-load variant).
const CUtensorMap* tma_map(b200_plan* p, float2* fw, int kind, int L) {
  if (!p->tma) p->tma = new TmaMaps();
  TmaMaps* tm = (TmaMaps*)p->tma;
  if (tm->base != fw || tm->T != p->ntrans_max) {
    tm->ok[0] = tm->ok[1] = false;
    tm->base = fw;
    tm->T = p->ntrans_max;
  }
  if (tm->ok[kind]) return &tm->map[kind];
  static EncodeTiledFn encode = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (EncodeTiledFn)fn;
    else
      cudaGetLastError();
  }
  if (!encode) return nullptr;
  const Geom& g = p->g;
  const cuuint64_t d0 = g.nf[g.dim - 1], d1 = g.nf[g.dim - 2], d2 = g.dim == 3 ? g.nf[0] : 1;
  const cuuint64_t dims[4] = {d0, d1, d2, (cuuint64_t)p->ntrans_max};
  const cuuint64_t strides[3] = {d0 * 8, d0 * d1 * 8, (cuuint64_t)g.nftot * 8};
  const cuuint32_t bz = (cuuint32_t)(L / 4 < 32 ? L / 4 : 32);
  const cuuint32_t box[4] = {(cuuint32_t)TX, kind == 0 ? bz : 1u, kind == 1 ? bz : 1u, 1u};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(&tm->map[kind], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, fw, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nullptr;
  tm->ok[kind] = true;
  return &tm->map[kind];
}

template <int L, int DIR, bool MUL, int EMODE, int KIN, int KOUT>
int launch_strided_m(const StridedArgs& A, int ntx, int nouter, int T, const float2* tw, cudaStream_t st,
                     const CUtensorMap* tmap) {
  const size_t smem = (size_t)L * TX * sizeof(float2);
  const dim3 grid = MUL ? dim3(T, nouter, ntx) : dim3(ntx, nouter, T);
  if constexpr (L >= 128 && !MUL) {
    if (tmap) {
      auto kern = k_fft_strided<L, DIR, MUL, EMODE, KIN, KOUT, true>;
      static PerDeviceOnce once;
      if (once.first()) B200_TRY(set_smem(kern, smem));
      kern<<<grid, FT, smem, st>>>(A, tw, *tmap);
      CHECK_LAUNCH();
      return B200_OK;
    }
  }
  auto kern = k_fft_strided<L, DIR, MUL, EMODE, KIN, KOUT, false>;
  static PerDeviceOnce once;
  if (once.first()) B200_TRY(set_smem(kern, smem));
  static const CUtensorMap none{};
  kern<<<grid, FT, smem, st>>>(A, tw, none);
  CHECK_LAUNCH();
  return B200_OK;
}

// (32-column tiles -- 256-byte segments per grid row, 512 threads, one CTA per SM -- were measured at
// cfg-C: 50.0 ms for the six passes against 42.4 ms with 16-column tiles; not kept.  Neither were, in round 2
// (six passes at 30.1 ms):
//  - a three-step radix-8 version of the pass for L = 512 (8 x 8 x 8 like k_fft_rows_t1_512: 40 registers, 512
//    threads, 48 warps per SM instead of 24, two more shared-memory trips and 14 instead of 31/4 twiddles per 8
//    points): correct to 1.5e-7 and 8 % slower on z- and y-passes alike -- the passes are bound by instruction
//    issue, not by latency that more warps would hide;
//  - persistent CTAs that fetch the empty-tile strings of their next tile while they work (the string ->
//    barrier -> loads round trip at the start of a CTA is 21 % of the z-passes' stall samples): 34.0 ms, and
//    32.8 ms for the same code with one tile per CTA -- the tile loop and its index arithmetic cost more than
//    the round trip;
//  - per-row strings that let the 3-D y-passes skip the rows of columns the z-pass skips (21 % of their
//    traffic at cfg-C): 30.3 ms -- not bound by bytes;
//  - type-2 z-pass: the loads of step A issued before the strings are waited for (the strings concern its
//    outputs only): 30.2 against 29.9 ms.)
template <int L, int DIR>
int launch_strided(const StridedArgs& A, int ntx, int nouter, int T, const float2* tw, cudaStream_t st,
                   const CUtensorMap* tmap) {
  // sigma = 2 patterns: modes of an N = L / 2 image <-> Keep{L / 4, L / 4} (KIN / KOUT 1); everything <-> Keep{L, 0} (2)
  const bool in_half = A.in.np == L / 4 && A.in.nm == L / 4, in_all = A.in.np == L && A.in.nm == 0;
  const bool out_half = A.out.np == L / 4 && A.out.nm == L / 4, out_all = A.out.np == L && A.out.nm == 0;
  if (A.mul) return launch_strided_m<L, DIR, true, 0, 0, 0>(A, ntx, nouter, T, tw, st, nullptr);
  if (in_half && out_all) {  // type 2
    if (A.empty && A.empty_out) return launch_strided_m<L, DIR, false, 2, 1, 2>(A, ntx, nouter, T, tw, st, tmap);
    if (!A.empty) return launch_strided_m<L, DIR, false, 0, 1, 2>(A, ntx, nouter, T, tw, st, tmap);
  }
  if (in_all && out_half) {  // type 1
    if (A.empty && !A.empty_out) return launch_strided_m<L, DIR, false, 1, 2, 1>(A, ntx, nouter, T, tw, st, tmap);
    if (!A.empty) return launch_strided_m<L, DIR, false, 0, 2, 1>(A, ntx, nouter, T, tw, st, tmap);
  }
  if (A.empty && A.empty_out) return launch_strided_m<L, DIR, false, 2, 0, 0>(A, ntx, nouter, T, tw, st, tmap);
  if (A.empty) return launch_strided_m<L, DIR, false, 1, 0, 0>(A, ntx, nouter, T, tw, st, tmap);
  return launch_strided_m<L, DIR, false, 0, 0, 0>(A, ntx, nouter, T, tw, st, tmap);
}

template <int L, int DIR, bool HALF>
int launch_rows_t2_h(const RowArgs& A, int nrows, const float2* tw, cudaStream_t st) {
  auto kern = k_fft_rows_t2<L, DIR, HALF>;
  const size_t smem = (size_t)TX * Split<L>::R1 * (Split<L>::R2 + 1) * sizeof(float2);
  static PerDeviceOnce once;
  if (once.first()) B200_TRY(set_smem(kern, smem));
  kern<<<dim3((unsigned)ceil_div(nrows, TX) * (unsigned)A.T, 1, 1), FT, smem, st>>>(A, nrows, tw);
  CHECK_LAUNCH();
  return B200_OK;
}

template <int L, int DIR>
int launch_rows_t2(const RowArgs& A, int nrows, const float2* tw, cudaStream_t st) {
  const int Nx = A.g.N[A.g.dim - 1];
  if (2 * Nx == L && Nx % 2 == 0) return launch_rows_t2_h<L, DIR, true>(A, nrows, tw, st);
  return launch_rows_t2_h<L, DIR, false>(A, nrows, tw, st);
}

template <int L, int DIR, bool HALF>
int launch_rows_t1_h(const RowArgs& A, int nrows, const float2* tw, cudaStream_t st) {
  auto kern = k_fft_rows_t1<L, DIR, HALF>;
  const size_t smem = ((size_t)TXR1 * Split<L>::R1 * (Split<L>::R2 + 1) + 2 * (size_t)TXR1 * L) * sizeof(float2);
  static PerDeviceOnce once;
  if (once.first()) B200_TRY(set_smem(kern, smem));
  kern<<<dim3(ceil_div(nrows, TXR1), 1, A.smaps ? 1 : A.T), Split<L>::R1 * TXR1, smem, st>>>(A, nrows, tw);
  CHECK_LAUNCH();
  return B200_OK;
}

template <int L, int DIR>
int launch_rows_t1(const RowArgs& A, int nrows, const float2* tw, cudaStream_t st) {
  const int Nx = A.g.N[A.g.dim - 1];
  if (L == 512 && Nx == 256) {
    k_fft_rows_t1_512<DIR><<<dim3(ceil_div(nrows, 4), 1, A.smaps ? 1 : A.T), 256, 0, st>>>(A, nrows, tw);
    CHECK_LAUNCH();
    return B200_OK;
  }
  if (2 * Nx == L && Nx % 2 == 0) return launch_rows_t1_h<L, DIR, true>(A, nrows, tw, st);
  return launch_rows_t1_h<L, DIR, false>(A, nrows, tw, st);
}

// pass along axis `a` (not the fastest one) of the [nf0][nf1][nf2] (or [nf0][nf1]) grid
int strided_pass(b200_plan* p, float2* fw, int T, int a, int dir, Keep in, Keep out, Keep outer_keep,
                 cudaStream_t st, const float* mul = nullptr, const uint32_t* empty = nullptr,
                 int empty_out = 0) {
  const Geom& g = p->g;
  StridedArgs A;
  A.base = fw;
  A.mul = mul;
  A.empty = (a == 0 && TX == 16 && g.dim == 3) ? empty : nullptr;  // (2-D grids are small: not worth it)
  A.empty_out = empty_out;
  A.empty_3d = g.dim == 3 ? 1 : 0;
  A.nyh = p->empty_nyh;
  A.nbx = p->empty_nbx;
  A.coil_stride = g.nftot;
  A.in = in;
  A.out = out;
  const int nfx = g.nf[g.dim - 1];
  int nouter = 1;
  if (g.dim == 3 && a == 0) {  // z-pass: outer = y
    A.stride_n = (long long)g.nf[1] * g.nf[2];
    A.outer_stride = g.nf[2];
    A.outer_L = g.nf[1];
    A.outer_keep = outer_keep;
    nouter = outer_keep.np + outer_keep.nm;
  } else if (g.dim == 3 && a == 1) {  // y-pass: outer = z
    A.stride_n = g.nf[2];
    A.outer_stride = (long long)g.nf[1] * g.nf[2];
    A.outer_L = g.nf[0];
    A.outer_keep = outer_keep;
    nouter = outer_keep.np + outer_keep.nm;
  } else {  // 2-D, axis 0
    A.stride_n = g.nf[1];
    A.outer_stride = 0;
    A.outer_L = 1;
    A.outer_keep = Keep{1, 0};
    nouter = 1;
  }
  const int L = g.nf[a];
  // option 2 = 3: bulk tensor loads (TMA) for the tiles of the strided passes; the tensor map describes the
  // plan's own workspace, which is what every caller passes
  const CUtensorMap* tmap = nullptr;
  A.tma_z = (g.dim == 3 && a == 0) ? 1 : 0;
  A.lookahead = p->fft_lookahead;
  // (default: the passes whose tile lies inside one plane -- 3-D y-pass, 2-D pass -- take the bulk loads:
  // 4.38 -> 4.13 ms and 4.47 -> 3.85 ms at 512^3 x 32 coils under ncu; the z-passes, 512 rows 2 MB apart,
  // are unchanged by them: profiles/r02_fft_passes_full.txt)
  const bool want_tma = p->fft_method == 3 || (p->fft_method == 0 && !A.tma_z);
  if (want_tma && L >= 128 && !mul && fw == p->d_fw) tmap = tma_map(p, fw, A.tma_z, L);
  DISPATCH_L(L, dir, return (launch_strided<LL, DD>(A, nfx / TX, nouter, T, p->d_tw[a], st, tmap)));
  return B200_OK;
}

bool pow2_ok(int L) { return L >= 32 && L <= 1024 && (L & (L - 1)) == 0; }

}  // namespace

void fftp_free(b200_plan* p) {
  if (p->tma) delete (TmaMaps*)p->tma;
  p->tma = nullptr;
}

bool fftp_supported(const b200_plan* p) {
  const Geom& g = p->g;
  if (p->flags & B200_SPREAD_ONLY) return false;
  if (g.dim < 2 || g.dim > 3) return false;
  for (int a = 0; a < g.dim; ++a)
    if (!pow2_ok(g.nf[a]) || g.N[a] > g.nf[a]) return false;
  return true;
}

// K4a + FFT:  image(s) -> oversampled grid, all T coils
int fftp_type2(b200_plan* p, const float2* img, const float2* smaps, float2* fw, int T, int isign,
               int conj_smaps, cudaStream_t st, const float* mul, const uint32_t* unread) {
  B200_TRY(ensure_twiddles(p));
  const Geom& g = p->g;
  const int dir = isign < 0 ? -1 : 1;
  RowArgs R{};
  R.g = g;
  R.img_in = img;
  R.smaps = smaps;
  R.fw = fw;
  R.d_slow0 = p->dvec(0);
  R.d_slow1 = g.dim == 3 ? p->dvec(1) : nullptr;
  R.d_fast = p->dvec(g.dim - 1);
  R.T = T;
  R.conj_smaps = conj_smaps;
  const int nrows = g.dim == 3 ? g.N[0] * g.N[1] : g.N[0];
  const int Lx = g.nf[g.dim - 1];
  DISPATCH_L(Lx, dir, B200_TRY((launch_rows_t2<LL, DD>(R, nrows, p->d_tw[g.dim - 1], st))));
  if (g.dim == 3) {
    // y-pass on the N0 non-zero planes, then z-pass everywhere
    B200_TRY(strided_pass(p, fw, T, 1, dir, keep_modes(g.N[1]), keep_all(g.nf[1]), keep_modes(g.N[0]), st));
    B200_TRY(strided_pass(p, fw, T, 0, dir, keep_modes(g.N[0]), keep_all(g.nf[0]), keep_all(g.nf[1]), st, mul,
                          mul ? nullptr : unread, 1));
  } else {
    B200_TRY(strided_pass(p, fw, T, 0, dir, keep_modes(g.N[0]), keep_all(g.nf[0]), Keep{1, 0}, st, mul,
                          mul ? nullptr : unread, 1));
  }
  return B200_OK;
}

// FFT + K4b:  oversampled grid -> image(s)
int fftp_type1(b200_plan* p, float2* fw, const float2* smaps, float2* img, int T, int accumulate,
               int isign, float scale, int conj_smaps, cudaStream_t st, const uint32_t* empty) {
  B200_TRY(ensure_twiddles(p));
  const Geom& g = p->g;
  const int dir = isign < 0 ? -1 : 1;
  if (g.dim == 3) {
    B200_TRY(strided_pass(p, fw, T, 0, dir, keep_all(g.nf[0]), keep_modes(g.N[0]), keep_all(g.nf[1]), st, nullptr,
                          empty));
    B200_TRY(strided_pass(p, fw, T, 1, dir, keep_all(g.nf[1]), keep_modes(g.N[1]), keep_modes(g.N[0]), st));
  } else {
    B200_TRY(strided_pass(p, fw, T, 0, dir, keep_all(g.nf[0]), keep_modes(g.N[0]), Keep{1, 0}, st, nullptr, empty));
  }
  RowArgs R{};
  R.g = g;
  R.img_out = img;
  R.smaps = smaps;
  R.fw = fw;
  R.d_slow0 = p->dvec(0);
  R.d_slow1 = g.dim == 3 ? p->dvec(1) : nullptr;
  R.d_fast = p->dvec(g.dim - 1);
  R.T = T;
  R.conj_smaps = conj_smaps;
  R.accumulate = accumulate;
  R.scale = scale;
  const int nrows = g.dim == 3 ? g.N[0] * g.N[1] : g.N[0];
  const int Lx = g.nf[g.dim - 1];
  DISPATCH_L(Lx, dir, B200_TRY((launch_rows_t1<LL, DD>(R, nrows, p->d_tw[g.dim - 1], st))));
  return B200_OK;
}
