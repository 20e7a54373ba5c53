// fft_reg.cuh -- in-register power-of-two FFTs (R <= 32 complex points per thread) built from
// Blackwell's packed float2 arithmetic (__ffma2_rn / __fadd2_rn / __fmul2_rn -> SASS FFMA2 ...).
// Building block of the pruned FFT passes in fft_pruned.cu.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

namespace fftreg {

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
__device__ __forceinline__ void sfor(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    sfor<B + 1, E>(f);
  }
}

// cos / sin of 2 pi k / 32, k = 0..15
__host__ __device__ constexpr float cos32(int k) {
  constexpr float t[16] = {1.f,           0.98078528f,   0.923879533f,  0.831469612f,
                           0.707106781f,  0.555570233f,  0.382683432f,  0.195090322f,
                           0.f,           -0.195090322f, -0.382683432f, -0.555570233f,
                           -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
  return t[k];
}
__host__ __device__ constexpr float sin32(int k) {
  constexpr float t[16] = {0.f,          0.195090322f, 0.382683432f, 0.555570233f,
                           0.707106781f, 0.831469612f, 0.923879533f, 0.98078528f,
                           1.f,          0.98078528f,  0.923879533f, 0.831469612f,
                           0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f};
  return t[k];
}

__host__ __device__ constexpr int brev(int k, int R) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
}
// a * (c + i s)
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
  return __ffma2_rn(a, make_float2(c, c), __fmul2_rn(make_float2(a.y, a.x), make_float2(-s, s)));
}
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return cmul_cs(a, w.x, w.y); }

// d * exp(DIR * 2 pi i K / 32), K in [0, 16)
template <int K, int DIR>
__device__ __forceinline__ float2 mul_tw32(float2 d) {
  if constexpr (K == 0) {
    return d;
  } else if constexpr (K == 8) {
    return DIR > 0 ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  } else {
    return cmul_cs(d, cos32(K), DIR > 0 ? sin32(K) : -sin32(K));
  }
}

template <int R, int H, int DIR>
struct Stage {
  static __device__ __forceinline__ void run(float2 (&a)[R]) {
    sfor<0, R / 2>([&](auto I) {
      constexpr int i = decltype(I)::value;
      constexpr int g = (i / H) * 2 * H;
      constexpr int j = i % H;
      constexpr int k32 = j * (16 / H);
      const float2 u = a[g + j], v = a[g + j + H];
      a[g + j] = cadd(u, v);
      a[g + j + H] = mul_tw32<k32, DIR>(csub(u, v));
    });
    if constexpr (H > 1) Stage<R, H / 2, DIR>::run(a);
  }
};

// In-place decimation-in-frequency FFT: on return A[k] = a[brev(k, R)],
// A[k] = sum_n a[n] exp(DIR * 2 pi i n k / R).
template <int R, int DIR>
__device__ __forceinline__ void fft(float2 (&a)[R]) {
  static_assert(R >= 2 && R <= 32 && (R & (R - 1)) == 0, "R must be a power of two <= 32");
  Stage<R, R / 2, DIR>::run(a);
}

// -d * exp(DIR * 2 pi i K / 32), K in [0, 16)
template <int K, int DIR>
__device__ __forceinline__ float2 mul_tw32_neg(float2 d) {
  if constexpr (K == 0) {
    return make_float2(-d.x, -d.y);
  } else if constexpr (K == 8) {
    return DIR > 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
  } else {
    return cmul_cs(d, -cos32(K), DIR > 0 ? -sin32(K) : sin32(K));
  }
}

// The same transform for inputs whose middle half is zero, a[n] = 0 for R/4 <= n < 3R/4 (a zero-padded
// signal in FFT order, sigma = 2): every butterfly of the first stage has one zero input, so the stage is R/2
// twiddle multiplies and no additions, and the zeros are never materialised.  a[R/4 .. 3R/4) need not be
// initialised.
template <int R, int DIR>
__device__ __forceinline__ void fft_zero_middle(float2 (&a)[R]) {
  static_assert(R >= 4 && R <= 32 && (R & (R - 1)) == 0, "R must be a power of two in [4, 32]");
  constexpr int H = R / 2;
  sfor<0, H>([&](auto I) {
    constexpr int j = decltype(I)::value;
    constexpr int k32 = j * (16 / H);
    if constexpr (j < R / 4) {
      a[j + H] = mul_tw32<k32, DIR>(a[j]);  // u + 0, (u - 0) w
    } else {
      const float2 v = a[j + H];
      a[j] = v;  // 0 + v, (0 - v) w
      a[j + H] = mul_tw32_neg<k32, DIR>(v);
    }
  });
  Stage<R, H / 2, DIR>::run(a);
}

}  // namespace fftreg
