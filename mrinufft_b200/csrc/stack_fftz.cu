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
#include "fft_smem.cuh"

namespace {

using namespace fftsm;


struct ZArgs {
  const float2* img;    // forward: (C | 1, X, Y, Z) input; adjoint: unused
  float2* out;          // adjoint: (C | 1, X, Y, Z) output
  const float2* smaps;  // nullptr or (C, X, Y, Z)
  float2* planes;       // (C * NZ, X, Y): coil axis of the 2-D operator
  const int* zsel;      // [NZ] kz plane of stack j
  const float2* tw;     // [Z] exp(-2 pi i t / Z)
  int C, X, Y, NZ;
  float scale;
  Rows r;               // Z, (x, y) columns per CTA (consecutive y), shared-memory row stride, radices
};

// grid.x = X * ceil(Y / rows) tiles, grid.y = coils (1 with smaps: the CTA walks the coils of its tile)
template <bool ADJ, bool SENSE>
__global__ void __launch_bounds__(ZT) k_stack_fftz(const ZArgs g) {
  extern __shared__ float2 zsm[];
  float2* tw = zsm;                       // [Z]
  const int Z = g.r.Z, h = Z / 2, rows = g.r.rows, zp = g.r.zp;
  float2* bufa = tw + Z;                  // [rows][zp]
  float2* bufb = bufa + rows * zp;        // [rows][zp]
  float2* accb = bufb + rows * zp;        // [rows][zp]  (ADJ && SENSE only)
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
        const int row = fdiv(i, g.r.zmagic), z = i - row * Z;
        float2 v = src[i];
        if (SENSE) v = cmulf(v, sm[i]);
        int n = z - h;
        if (n < 0) n += Z;
        bufa[row * zp + n] = v;
      }
      __syncthreads();
      const float2* res = fft_rows<false>(bufa, bufb, tw, g.r);
      // fftshift + plane selection + (coil, stack)-major planes: rows (= y) fastest
      for (int i = threadIdx.x; i < g.NZ * rows; i += ZT) {
        const int j = i >> g.r.rshift, row = i & (rows - 1);
        if (row < nrow) {
          int n = g.zsel[j] - h;
          if (n < 0) n += Z;
          const float2 v = res[row * zp + n];
          g.planes[(((size_t)c * g.NZ + j) * g.X + x) * g.Y + y0 + row] = make_float2(v.x * g.scale, v.y * g.scale);
        }
      }
      __syncthreads();
    } else {
      for (int i = threadIdx.x; i < rows * zp; i += ZT) bufa[i] = make_float2(0.f, 0.f);
      __syncthreads();
      for (int i = threadIdx.x; i < g.NZ * rows; i += ZT) {
        const int j = i >> g.r.rshift, row = i & (rows - 1);
        if (row < nrow) {
          int n = g.zsel[j] - h;
          if (n < 0) n += Z;
          bufa[row * zp + n] = g.planes[(((size_t)c * g.NZ + j) * g.X + x) * g.Y + y0 + row];
        }
      }
      __syncthreads();
      const float2* res = fft_rows<true>(bufa, bufb, tw, g.r);
      float2* dst = g.out + (SENSE ? 0 : (size_t)c * vol) + colZ;
      const float2* sm = SENSE ? g.smaps + (size_t)c * vol + colZ : nullptr;
      for (int i = threadIdx.x; i < nrow * Z; i += ZT) {
        const int row = fdiv(i, g.r.zmagic), z = i - row * Z;
        int n = z - h;
        if (n < 0) n += Z;
        float2 v = res[row * zp + n];
        v = make_float2(v.x * g.scale, v.y * g.scale);
        if (SENSE) {
          const float2 s = sm[i];
          v = cmulf(v, make_float2(s.x, -s.y));
          float2* a = accb + row * zp + z;   // every thread owns its (row, z) slots across the coil loop
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
template <bool ADJ>
int launch(ZArgs& g, bool sense, cudaStream_t st) {
  B200_TRY(twiddles<float2>(g.r.Z, &g.tw));
  g.r.zp = g.r.Z | 1;
  const int nbuf = (ADJ && sense) ? 3 : 2;
  const size_t budget = 216 * 1024;
  int rows = 16;
  auto bytes = [&](int r) { return ((size_t)nbuf * r * g.r.zp + g.r.Z) * sizeof(float2); };
  while (rows > 1 && bytes(rows) > budget) rows >>= 1;
  if (bytes(rows) > budget) {
    b200_set_error("b200_stack_fftz: Z=%d does not fit one CTA's shared memory", g.r.Z);
    return B200_EINVAL;
  }
  g.r.rows = rows;
  B200_TRY(prepare(&g.r));
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
  g.C = C; g.X = X; g.Y = Y; g.r.Z = Z; g.NZ = NZ;
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
  g.C = C; g.X = X; g.Y = Y; g.r.Z = Z; g.NZ = NZ;
  g.scale = scale;
  return launch<true>(g, smaps != nullptr, (cudaStream_t)stream);
}
