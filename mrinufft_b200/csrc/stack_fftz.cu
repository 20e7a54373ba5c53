// stack_fftz.cu -- the z transform of the stacked (2.5-D) operator, fused with everything around it.
//
// A stack-of-trajectories acquisition is "FFT along z, then one 2-D NUFFT per kz plane"
// (src/mrinufft/operators/stacked.py:178-195 `_fftz` / `_ifftz`, :203-240 `_op_sense` / `_op_calibless`,
// :254-300 the adjoints).  The reference does this with six full-array steps per coil batch (smaps
// multiply, ifftshift, fft, fftshift, plane selection, moveaxis + copy).  Here ONE kernel per direction
// does all of it on a tile of `rows` neighbouring (x, y..y+rows) columns of the volume:
//
//   forward: planes[c NZ + j, x, y] = scale * FFTc_z(img[c | 0, x, y, :] * smaps[c, x, y, :])[zsel[j]]
//   adjoint: out[c | 0, x, y, :]    = (sum_c) conj(smaps[c]) * scale * IFFTc_z(scatter_j planes[c NZ + j, x, y])
//
// FFTc = fftshift . fft . ifftshift: both shifts are index rotations at the load and at the store, the plane
// selection is a gather at the store, the (C NZ, X, Y) layout of the 2-D operator's coil axis is the store's
// address pattern (`rows` consecutive y per plane: full 128-byte lines), the SENSE coil sum stays in shared
// memory.  The transform itself is a Stockham autosort FFT in shared memory for ANY length Z: radix 4 / 2 /
// 3 / 5 / 7 butterflies in registers, any other prime factor p as a direct p-point DFT (O(p^2) per butterfly,
// only what the factorisation leaves).  Twiddles come from one table exp(-2 pi i t / Z) computed in double
// on the host, once per (device, Z).
#include <cmath>
#include <map>
#include <mutex>

#include "common.cuh"

namespace {

constexpr int ZT = 256;      // threads per CTA
constexpr int MAXRAD = 24;   // radices of a length < 2^24

struct Radices {
  int n;
  int r[MAXRAD];
};

struct ZArgs {
  const float2* img;    // forward: (C | 1, X, Y, Z) input; adjoint: unused
  float2* out;          // adjoint: (C | 1, X, Y, Z) output
  const float2* smaps;  // nullptr or (C, X, Y, Z)
  float2* planes;       // (C * NZ, X, Y): coil axis of the 2-D operator
  const int* zsel;      // [NZ] kz plane of stack j
  const float2* tw;     // [Z] exp(-2 pi i t / Z)
  int C, X, Y, Z, NZ;
  int rows;             // (x, y) columns per CTA (consecutive y)
  int zp;               // shared-memory row stride in float2 (odd: the row-fastest accesses spread over the banks)
  float scale;
  Radices rad;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by exp(-+ i pi / 2): -i forward, +i inverse
template <bool INV>
__device__ __forceinline__ float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// One Stockham stage of radix R on one row: butterfly (p, q), sub-length n = R m, stride s (s n = Z).
//   y[q + s (R p + k)] = w_n^(p k) sum_j x[q + s (p + j m)] w_R^(j k),   w_L = tw[Z / L]
template <int R, bool INV>
__device__ __forceinline__ void butterfly(const float2* __restrict__ src, float2* __restrict__ dst,
                                          const float2* __restrict__ tw, int p, int q, int s, int m, int Z) {
  float2 a[R];
#pragma unroll
  for (int j = 0; j < R; ++j) a[j] = src[q + s * (p + j * m)];
  float2 o[R];
  if constexpr (R == 2) {
    o[0] = caddf(a[0], a[1]);
    o[1] = csubf(a[0], a[1]);
  } else if constexpr (R == 4) {
    const float2 t0 = caddf(a[0], a[2]), t1 = csubf(a[0], a[2]);
    const float2 t2 = caddf(a[1], a[3]), t3 = rot90<INV>(csubf(a[1], a[3]));
    o[0] = caddf(t0, t2);
    o[1] = caddf(t1, t3);
    o[2] = csubf(t0, t2);
    o[3] = csubf(t1, t3);
  } else {
    float2 w[R];
#pragma unroll
    for (int j = 1; j < R; ++j) w[j] = tw[j * (Z / R)];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      float2 acc = a[0];
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
  float2* d = dst + q + s * (R * p);
  d[0] = o[0];
#pragma unroll
  for (int k = 1; k < R; ++k) d[s * k] = t1 ? cmulf(o[k], tw[t1 * k]) : o[k];
}

// any other (prime) radix: direct r-point DFT, inputs re-read from shared memory
__device__ __forceinline__ void butterfly_any(const float2* __restrict__ src, float2* __restrict__ dst,
                                              const float2* __restrict__ tw, int r, int p, int q, int s, int m,
                                              int Z) {
  const int zr = Z / r;
  for (int k = 0; k < r; ++k) {
    float2 acc = src[q + s * p];
    int e = 0;
    for (int j = 1; j < r; ++j) {
      e += k;
      if (e >= r) e -= r;
      acc = caddf(acc, cmulf(src[q + s * (p + j * m)], tw[e * zr]));
    }
    dst[q + s * (r * p + k)] = cmulf(acc, tw[p * k * s]);
  }
}

// all stages on the `rows` rows of the tile; returns the buffer that holds the result (natural order)
template <bool INV>
__device__ __forceinline__ float2* fft_rows(float2* a, float2* b, const float2* tw, const ZArgs& g) {
  int s = 1, n = g.Z;
  for (int st = 0; st < g.rad.n; ++st) {
    const int r = g.rad.r[st], m = n / r, nb = g.Z / r;
    for (int i = threadIdx.x; i < g.rows * nb; i += ZT) {
      const int row = i / nb, bf = i - row * nb, p = bf / s, q = bf - p * s;
      const float2* x = a + row * g.zp;
      float2* y = b + row * g.zp;
      switch (r) {
        case 2: butterfly<2, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 3: butterfly<3, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 4: butterfly<4, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 5: butterfly<5, INV>(x, y, tw, p, q, s, m, g.Z); break;
        case 7: butterfly<7, INV>(x, y, tw, p, q, s, m, g.Z); break;
        default: butterfly_any(x, y, tw, r, p, q, s, m, g.Z); break;
      }
    }
    __syncthreads();
    float2* t = a;
    a = b;
    b = t;
    n = m;
    s *= r;
  }
  return a;
}

// grid.x = X * ceil(Y / rows) tiles, grid.y = coils (1 with smaps: the CTA walks the coils of its tile)
template <bool ADJ, bool SENSE>
__global__ void __launch_bounds__(ZT) k_stack_fftz(const ZArgs g) {
  extern __shared__ float2 zsm[];
  float2* tw = zsm;                       // [Z]
  float2* bufa = tw + g.Z;                // [rows][zp]
  float2* bufb = bufa + g.rows * g.zp;    // [rows][zp]
  float2* accb = bufb + g.rows * g.zp;    // [rows][zp]  (ADJ && SENSE only)
  const int Z = g.Z, h = Z / 2, rows = g.rows;
  const int ytiles = (g.Y + rows - 1) / rows;
  const int x = blockIdx.x / ytiles, y0 = (blockIdx.x - x * ytiles) * rows;
  const int nrow = min(rows, g.Y - y0);
  for (int t = threadIdx.x; t < Z; t += ZT) {
    const float2 w = g.tw[t];
    tw[t] = ADJ ? make_float2(w.x, -w.y) : w;
  }
  const int c0 = SENSE ? 0 : blockIdx.y, c1 = SENSE ? g.C : c0 + 1;
  const size_t colZ = ((size_t)x * g.Y + y0) * Z;      // first element of the tile inside one coil volume
  const size_t vol = (size_t)g.X * g.Y * Z;
  for (int c = c0; c < c1; ++c) {
    if constexpr (!ADJ) {
      // load (coalesced along z), sensitivity map, ifftshift: element z goes to slot (z - h) mod Z
      const float2* src = g.img + (SENSE ? 0 : (size_t)c * vol) + colZ;
      const float2* sm = SENSE ? g.smaps + (size_t)c * vol + colZ : nullptr;
      for (int i = threadIdx.x; i < nrow * Z; i += ZT) {
        const int row = i / Z, z = i - row * Z;
        float2 v = src[i];
        if (SENSE) v = cmulf(v, sm[i]);
        int n = z - h;
        if (n < 0) n += Z;
        bufa[row * g.zp + n] = v;
      }
      __syncthreads();
      const float2* res = fft_rows<false>(bufa, bufb, tw, g);
      // fftshift + plane selection + (coil, stack)-major planes: rows (= y) fastest
      for (int i = threadIdx.x; i < g.NZ * rows; i += ZT) {
        const int j = i / rows, row = i - j * rows;
        if (row < nrow) {
          int n = g.zsel[j] - h;
          if (n < 0) n += Z;
          const float2 v = res[row * g.zp + n];
          g.planes[(((size_t)c * g.NZ + j) * g.X + x) * g.Y + y0 + row] = make_float2(v.x * g.scale, v.y * g.scale);
        }
      }
      __syncthreads();
    } else {
      for (int i = threadIdx.x; i < rows * g.zp; i += ZT) bufa[i] = make_float2(0.f, 0.f);
      __syncthreads();
      for (int i = threadIdx.x; i < g.NZ * rows; i += ZT) {
        const int j = i / rows, row = i - j * rows;
        if (row < nrow) {
          int n = g.zsel[j] - h;
          if (n < 0) n += Z;
          bufa[row * g.zp + n] = g.planes[(((size_t)c * g.NZ + j) * g.X + x) * g.Y + y0 + row];
        }
      }
      __syncthreads();
      const float2* res = fft_rows<true>(bufa, bufb, tw, g);
      float2* dst = g.out + (SENSE ? 0 : (size_t)c * vol) + colZ;
      const float2* sm = SENSE ? g.smaps + (size_t)c * vol + colZ : nullptr;
      for (int i = threadIdx.x; i < nrow * Z; i += ZT) {
        const int row = i / Z, z = i - row * Z;
        int n = z - h;
        if (n < 0) n += Z;
        float2 v = res[row * g.zp + n];
        v = make_float2(v.x * g.scale, v.y * g.scale);
        if (SENSE) {
          const float2 s = sm[i];
          v = cmulf(v, make_float2(s.x, -s.y));
          float2* a = accb + row * g.zp + z;   // every thread owns its (row, z) slots across the coil loop
          if (c > c0) v = caddf(v, *a);
          if (c + 1 < c1) *a = v;
          else dst[i] = v;
        } else {
          dst[i] = v;
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------- host side
std::mutex g_tw_mutex;
std::map<std::pair<int, int>, float2*> g_tw;  // (device, Z) -> table

int twiddles(int Z, const float2** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tw_mutex);
  auto it = g_tw.find({dev, Z});
  if (it == g_tw.end()) {
    std::vector<float2> h(Z);
    const double pi = 3.14159265358979323846;
    for (int t = 0; t < Z; ++t) {
      const double a = -2.0 * pi * (double)t / (double)Z;
      h[t] = make_float2((float)cos(a), (float)sin(a));
    }
    float2* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)Z * sizeof(float2)));
    CUDA_TRY(cudaMemcpy(d, h.data(), (size_t)Z * sizeof(float2), cudaMemcpyHostToDevice));
    it = g_tw.emplace(std::make_pair(dev, Z), d).first;
  }
  *out = it->second;
  return B200_OK;
}

int factorise(int Z, Radices* rad) {
  rad->n = 0;
  auto push = [&](int r) {
    if (rad->n < MAXRAD) rad->r[rad->n] = r;
    ++rad->n;
  };
  while (Z % 4 == 0) { push(4); Z /= 4; }
  for (int p : {2, 3, 5, 7})
    while (Z % p == 0) { push(p); Z /= p; }
  for (int p = 11; Z > 1; p += 2)
    while (Z % p == 0) { push(p); Z /= p; }
  if (rad->n > MAXRAD) {
    b200_set_error("b200_stack_fftz: too many prime factors");
    return B200_EINVAL;
  }
  return B200_OK;
}

template <bool ADJ>
int launch(ZArgs& g, bool sense, cudaStream_t st) {
  B200_TRY(factorise(g.Z, &g.rad));
  B200_TRY(twiddles(g.Z, &g.tw));
  g.zp = g.Z | 1;
  const int nbuf = (ADJ && sense) ? 3 : 2;
  const size_t budget = 216 * 1024;
  int rows = 16;
  auto bytes = [&](int r) { return ((size_t)nbuf * r * g.zp + g.Z) * sizeof(float2); };
  while (rows > 1 && bytes(rows) > budget) rows >>= 1;
  if (bytes(rows) > budget) {
    b200_set_error("b200_stack_fftz: Z=%d does not fit one CTA's shared memory", g.Z);
    return B200_EINVAL;
  }
  g.rows = rows;
  const long long tiles = (long long)g.X * ((g.Y + rows - 1) / rows);
  if (tiles > 0x7fffffffLL || g.C > 65535) {
    b200_set_error("b200_stack_fftz: volume too large for one launch");
    return B200_EINVAL;
  }
  dim3 grid((unsigned)tiles, sense ? 1 : g.C);
  const size_t smem = bytes(rows);
  if (sense) {
    auto k = k_stack_fftz<ADJ, true>;
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, ZT, smem, st>>>(g);
  } else {
    auto k = k_stack_fftz<ADJ, false>;
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, ZT, smem, st>>>(g);
  }
  CHECK_LAUNCH();
  return B200_OK;
}

int check_args(const void* a, const void* b, const int32_t* zsel, int C, int X, int Y, int Z, int NZ) {
  if (!a || !b || !zsel) {
    b200_set_error("b200_stack_fftz: null argument");
    return B200_EINVAL;
  }
  if (C < 1 || X < 1 || Y < 1 || Z < 1 || NZ < 1) {
    b200_set_error("b200_stack_fftz: sizes must be positive (C=%d X=%d Y=%d Z=%d NZ=%d)", C, X, Y, Z, NZ);
    return B200_EINVAL;
  }
  return B200_OK;
}

}  // namespace

extern "C" int b200_stack_fftz_forward(const void* img, const void* smaps, void* planes, const int32_t* zsel,
                                       int C, int X, int Y, int Z, int NZ, float scale, void* stream) {
  B200_TRY(check_args(img, planes, zsel, C, X, Y, Z, NZ));
  ZArgs g{};
  g.img = (const float2*)img;
  g.smaps = (const float2*)smaps;
  g.planes = (float2*)planes;
  g.zsel = zsel;
  g.C = C; g.X = X; g.Y = Y; g.Z = Z; g.NZ = NZ;
  g.scale = scale;
  return launch<false>(g, smaps != nullptr, (cudaStream_t)stream);
}

extern "C" int b200_stack_fftz_adjoint(const void* planes, const void* smaps, void* out, const int32_t* zsel,
                                       int C, int X, int Y, int Z, int NZ, float scale, void* stream) {
  B200_TRY(check_args(planes, out, zsel, C, X, Y, Z, NZ));
  ZArgs g{};
  g.out = (float2*)out;
  g.smaps = (const float2*)smaps;
  g.planes = (float2*)const_cast<void*>(planes);
  g.zsel = zsel;
  g.C = C; g.X = X; g.Y = Y; g.Z = Z; g.NZ = NZ;
  g.scale = scale;
  return launch<true>(g, smaps != nullptr, (cudaStream_t)stream);
}
