// spread_tiled.cu -- "method 2": output-owned, atomic-free spreading (K2) and its transpose for
// interpolation (K3), written for sm_100a.
//
// Why not the usual shared-memory sub-grid + atomicAdd design: on sm_100a a float atomicAdd on
// shared memory compiles to an ATOMS.CAST.SPIN compare-and-swap loop (checked with cuobjdump),
// and a 3-D width-7 kernel needs 343 complex accumulations per point per coil.  Instead every
// fine-grid cell is OWNED by exactly one warp and accumulated in REGISTERS:
//
//   * a warp owns one grid "row": CX = 32 consecutive cells along the fastest axis at fixed slow
//     coordinates (z, y), for up to 32 coils -- lane = coil, register i = cell i (64 accumulators);
//   * the points whose footprint covers the row are found through the pencil-bin sort of K1
//     (bins of 1 x 1 x 32 cells, key order (z0, bx, y0)): for each of the w slow-axis offsets dz the
//     candidates with y0 in [y-w+1, y] are ONE contiguous range of sorted points per x-bin;
//   * per visiting point: 2w FFMA into statically indexed registers selected by a warp-uniform
//     switch on the x offset; weights come from a per-point table computed once per trajectory,
//     the point's sample values from a (sorted point, coil) transposed copy of the k-space batch
//     (one coalesced 256-byte load per visit);
//   * a row is written to HBM exactly once, through a shared-memory transpose, as full 256-byte
//     coalesced stores per coil -- no memset of the oversampled grid, no halo flush, no atomics,
//     bit-reproducible results.
//
// Interpolation is the exact transpose: the warp loads its row into registers once (coalesced),
// every visiting point takes its 7-tap dot product from registers and adds the partial sum into
// a (sorted point, coil) accumulator with one vector `red.global.add.v2.f32` per lane.
//
// Replaces finufft's spread/interp stage (call sites
// src/mrinufft/operators/interfaces/finufft.py:69,76; algorithm docs/explanations/nufft.rst:253-309).
#include "common.cuh"
#include "device_utils.cuh"

namespace {

constexpr int CX = 32;        // cells per row segment (== pencil-bin width)
constexpr int WARPS = 8;      // rows per CTA
constexpr int THREADS = WARPS * 32;
constexpr int REC = 24;       // floats per point record: wx[7] x0 | wy[7] y0 | wz[7] -

struct TiledState {
  float* d_rec = nullptr;     // [M][REC] per sorted point
  float2* d_kt = nullptr;     // [M][32] transposed (sorted point, coil) k-space batch
  size_t kt_bytes = 0;
  int* d_counter = nullptr;   // persistent-CTA work counter
  long long M = 0;
  bool rec_valid = false;
};

TiledState* state(b200_plan* p) {
  if (!p->tiled) p->tiled = new TiledState();
  return (TiledState*)p->tiled;
}

// ---------------------------------------------------------------------------------- pre-passes
template <int W>
__global__ void __launch_bounds__(256)
k_point_records(Geom g, long long M, const float* __restrict__ poly,
                const int32_t* __restrict__ o0, const int32_t* __restrict__ o1,
                const int32_t* __restrict__ o2, const float* __restrict__ f0,
                const float* __restrict__ f1, const float* __restrict__ f2,
                float* __restrict__ rec) {
  __shared__ float spoly[(B200_MAX_DEG + 1) * B200_MAX_W];
  for (int i = threadIdx.x; i < (g.deg + 1) * W; i += blockDim.x) spoly[i] = poly[i];
  __syncthreads();
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  // axis roles: x = fastest axis (dim-1), y = dim-2, z = dim-3
  const int32_t* op[3] = {o0, o1, o2};
  const float* fp[3] = {f0, f1, f2};
  float out[REC];
#pragma unroll
  for (int i = 0; i < REC; ++i) out[i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // r = 0: x, 1: y, 2: z
    const int a = g.dim - 1 - r;
    if (a < 0) {
      out[r * 8] = 1.f;  // unused axis: single unit tap
      continue;
    }
    const float z = fmaf(2.f, fp[a][s], (float)(W - 1));
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float acc = spoly[g.deg * W + i];
      for (int k = g.deg - 1; k >= 0; --k) acc = fmaf(acc, z, spoly[k * W + i]);
      out[r * 8 + i] = acc;
    }
    out[r * 8 + 7] = __int_as_float(op[a][s]);
  }
  float4* dst = reinterpret_cast<float4*>(rec + s * REC);
#pragma unroll
  for (int q = 0; q < REC / 4; ++q)
    dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

// kt[s][t] = ksp[t][perm[s]] * density[perm[s]]   (t < T; lanes t >= T are zero-filled)
__global__ void __launch_bounds__(256)
k_gather_kspace(long long M, int T, const int32_t* __restrict__ perm,
                const float2* __restrict__ ksp, const float* __restrict__ density,
                float2* __restrict__ kt) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long s = idx >> 5;
  int t = (int)(idx & 31);
  if (s >= M) return;
  float2 v = make_float2(0.f, 0.f);
  if (t < T) {
    const int j = perm[s];
    v = ksp[(long long)t * M + j];
    if (density) {
      const float d = density[j];
      v.x *= d;
      v.y *= d;
    }
  }
  kt[s * 32 + t] = v;
}

// ksp[t][perm[s]] = scale * kt[s][t] (- obs[t][perm[s]])
__global__ void __launch_bounds__(256)
k_scatter_kspace(long long M, int T, const int32_t* __restrict__ perm,
                 const float2* __restrict__ kt, float2* __restrict__ ksp, float scale,
                 const float2* __restrict__ obs) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long s = idx >> 5;
  int t = (int)(idx & 31);
  if (s >= M || t >= T) return;
  const int j = perm[s];
  float2 v = kt[s * 32 + t];
  v.x *= scale;
  v.y *= scale;
  const long long oi = (long long)t * M + j;
  if (obs) {
    const float2 y = obs[oi];
    v.x -= y.x;
    v.y -= y.y;
  }
  ksp[oi] = v;
}

// ---------------------------------------------------------------------------------- row kernels
template <int W, int OFF>
__device__ __forceinline__ void taps_spread(float2 (&acc)[CX], const float (&wx)[8], float2 v) {
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (OFF + i >= 0 && OFF + i < CX) {
      acc[OFF + i].x = fmaf(v.x, wx[i], acc[OFF + i].x);
      acc[OFF + i].y = fmaf(v.y, wx[i], acc[OFF + i].y);
    }
  }
}

template <int W, int OFF>
__device__ __forceinline__ float2 taps_interp(const float2 (&acc)[CX], const float (&wx)[8]) {
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (OFF + i >= 0 && OFF + i < CX) {
      r.x = fmaf(acc[OFF + i].x, wx[i], r.x);
      r.y = fmaf(acc[OFF + i].y, wx[i], r.y);
    }
  }
  return r;
}

#define OFF_CASES(M_)                                                                           \
  M_(-6) M_(-5) M_(-4) M_(-3) M_(-2) M_(-1) M_(0) M_(1) M_(2) M_(3) M_(4) M_(5) M_(6) M_(7)     \
  M_(8) M_(9) M_(10) M_(11) M_(12) M_(13) M_(14) M_(15) M_(16) M_(17) M_(18) M_(19) M_(20)      \
  M_(21) M_(22) M_(23) M_(24) M_(25) M_(26) M_(27) M_(28) M_(29) M_(30) M_(31)

struct RowJob {
  int z, y, bx;       // row coordinates (z = 0 in 2-D), x-tile index
  long long rowbase;  // linear index of (z, y, 0) in one coil's grid
};

// Decode tile -> this warp's row.  Tiles: (ROWS_Z x ROWS_Y) rows x one x-tile.
template <int DIM>
__device__ __forceinline__ bool decode_row(const Geom& g, long long tile, int warp, RowJob* job) {
  const int nfx = g.nf[DIM - 1];
  const int nbx = (nfx + CX - 1) / CX;
  if (DIM == 3) {
    constexpr int RZ = 2, RY = 4;
    const int nty = (g.nf[1] + RY - 1) / RY;
    const int bx = (int)(tile % nbx);
    long long r = tile / nbx;
    const int ty = (int)(r % nty);
    const int tz = (int)(r / nty);
    job->z = tz * RZ + warp / RY;
    job->y = ty * RY + warp % RY;
    job->bx = bx;
    if (job->z >= g.nf[0] || job->y >= g.nf[1]) return false;
    job->rowbase = ((long long)job->z * g.nf[1] + job->y) * nfx;
  } else {
    const int bx = (int)(tile % nbx);
    const int ty = (int)(tile / nbx);
    job->z = 0;
    job->y = ty * WARPS + warp;
    job->bx = bx;
    if (job->y >= g.nf[0]) return false;
    job->rowbase = (long long)job->y * nfx;
  }
  return true;
}

template <int DIM>
long long num_tiles(const Geom& g) {
  const int nbx = (g.nf[DIM - 1] + CX - 1) / CX;
  if (DIM == 3) return (long long)((g.nf[0] + 1) / 2) * ((g.nf[1] + 3) / 4) * nbx;
  return (long long)((g.nf[0] + WARPS - 1) / WARPS) * nbx;
}

// ---- row visits ---------------------------------------------------------------------------
// Every sorted point whose footprint touches the row is a "visit".  Visits are produced in two
// phases so that no load sits on the critical path of the accumulation loop:
//   1. FILTER (lane-parallel): the 32 lanes test 32 candidates of a key range at once (x offset
//      inside the tile?), compute wy*wz for their candidate and compact the survivors into a
//      per-warp shared-memory list of {s, off, wyz};
//   2. CONSUME (warp-uniform): the list is walked with the next entry's loads (list entry, the
//      point's x weights, the point's 32 coil values) issued before the current entry's FFMAs.
//   key order (setpts.cu): 3-D (z0 * nbx + bx) * nfy + y0 ; 2-D bx * nfy + y0
constexpr int LIST = 128;  // list entries per warp (int4 each)

template <int DIM, int W, typename Consume>
__device__ __forceinline__ void for_each_row_visit(const Geom& g, const RowJob& job,
                                                   const int32_t* __restrict__ bin_start,
                                                   const float* __restrict__ rec, int4* list,
                                                   int lane, Consume&& consume) {
  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2 >= 0 ? DIM - 2 : 0];
  const int nbx = (nfx + CX - 1) / CX;
  const int NZ = (DIM == 3) ? W : 1;
  int n_list = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int dz = 0; dz < NZ; ++dz) {
    int z0 = 0;
    if (DIM == 3) {
      z0 = job.z - dz;
      if (z0 < 0) z0 += g.nf[0];
    }
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      // side 0: the x-bin to the left (taps reaching into this tile), side 1: this tile's bin
      int bxx = job.bx - 1 + side;
      int xshift = job.bx * CX;  // off = x0 - xshift
      if (bxx < 0) {
        bxx = nbx - 1;
        xshift = nfx;  // wrapped: off = x0 - nfx
      }
      const long long kbase = ((long long)z0 * nbx + bxx) * nfy;
      const int ylo = job.y - (W - 1);
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        int a, b;  // inclusive y0 range of this part (periodic wrap -> up to two ranges)
        if (ylo >= 0) {
          if (part == 1) break;
          a = ylo;
          b = job.y;
        } else if (part == 0) {
          a = 0;
          b = job.y;
        } else {
          a = ylo + nfy;
          b = nfy - 1;
        }
        const int s_begin = __ldg(bin_start + kbase + a);
        const int s_end = __ldg(bin_start + kbase + b + 1);
        for (int base = s_begin; base < s_end; base += 32) {
          const int s = base + lane;
          bool ok = false;
          int off = 0;
          float wyz = 0.f;
          if (s < s_end) {
            const float* r = rec + (long long)s * REC;
            off = __float_as_int(__ldg(r + 7)) - xshift;
            ok = (off > -W) && (off < CX);
            if (ok) {
              int dy = job.y - __float_as_int(__ldg(r + 15));
              if (dy < 0) dy += nfy;
              wyz = __ldg(r + 8 + dy);
              if (DIM == 3) wyz *= __ldg(r + 16 + dz);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) list[n_list + __popc(m & lt_mask)] = make_int4(s, off, __float_as_int(wyz), 0);
          n_list += __popc(m);
          if (n_list > LIST - 32) {
            __syncwarp();
            consume(n_list);
            __syncwarp();
            n_list = 0;
          }
        }
      }
    }
  }
  __syncwarp();
  consume(n_list);
  __syncwarp();
}

struct Visit {
  int off;
  float wyz;
  float4 wa, wb;
  float2 v;
};

template <int DIM, int W>
__global__ void __launch_bounds__(THREADS, 2)
k_spread_rows(Geom g, int T, long long ntiles, const int32_t* __restrict__ bin_start,
              const float* __restrict__ rec, const float2* __restrict__ kt,
              float2* __restrict__ fw, int* __restrict__ counter) {
  extern __shared__ float2 sbuf[];  // [WARPS][32][33] transpose buffers, then [WARPS][LIST] int4
  __shared__ long long s_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* buf = sbuf + warp * (32 * 33);
  int4* list = reinterpret_cast<int4*>(sbuf + WARPS * 32 * 33) + warp * LIST;
  const int nfx = g.nf[DIM - 1];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= ntiles) break;
    RowJob job;
    if (!decode_row<DIM>(g, tile, warp, &job)) continue;
    float2 acc[CX];
#pragma unroll
    for (int i = 0; i < CX; ++i) acc[i] = make_float2(0.f, 0.f);

    auto load_visit = [&](int k) {
      Visit e;
      const int4 q = list[k];
      const float4* r = reinterpret_cast<const float4*>(rec + (long long)q.x * REC);
      e.off = q.y;
      e.wyz = __int_as_float(q.z);
      e.wa = __ldg(r);
      e.wb = __ldg(r + 1);
      e.v = __ldg(kt + (long long)q.x * 32 + lane);
      return e;
    };
    for_each_row_visit<DIM, W>(g, job, bin_start, rec, list, lane, [&](int n) {
      if (n == 0) return;
      Visit cur = load_visit(0);
      for (int k = 0; k < n; ++k) {
        Visit nxt = cur;
        if (k + 1 < n) nxt = load_visit(k + 1);
        const float wx[8] = {cur.wa.x, cur.wa.y, cur.wa.z, cur.wa.w, cur.wb.x, cur.wb.y, cur.wb.z, 0.f};
        const float2 v = make_float2(cur.v.x * cur.wyz, cur.v.y * cur.wyz);
        switch (cur.off) {
#define CASE_(O) case O: taps_spread<W, O>(acc, wx, v); break;
          OFF_CASES(CASE_)
#undef CASE_
          default: break;
        }
        cur = nxt;
      }
    });
    // flush: registers (lane = coil, i = cell) -> smem transpose -> coalesced rows per coil
#pragma unroll
    for (int i = 0; i < CX; ++i) buf[lane * 33 + i] = acc[i];
    __syncwarp();
    const int x = job.bx * CX + lane;
    if (x < nfx) {
      float2* dst = fw + job.rowbase + x;
      for (int t = 0; t < T; ++t) dst[(long long)t * g.nftot] = buf[t * 33 + lane];
    }
    __syncwarp();
  }
}

template <int DIM, int W>
__global__ void __launch_bounds__(THREADS, 2)
k_interp_rows(Geom g, int T, long long ntiles, const int32_t* __restrict__ bin_start,
              const float* __restrict__ rec, const float2* __restrict__ fw,
              float2* __restrict__ kt, int* __restrict__ counter) {
  extern __shared__ float2 sbuf[];
  __shared__ long long s_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* buf = sbuf + warp * (32 * 33);
  int4* list = reinterpret_cast<int4*>(sbuf + WARPS * 32 * 33) + warp * LIST;
  const int nfx = g.nf[DIM - 1];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= ntiles) break;
    RowJob job;
    if (!decode_row<DIM>(g, tile, warp, &job)) continue;
    // load the row: coalesced per coil -> smem -> registers (lane = coil)
    {
      const int x = job.bx * CX + lane;
      const float2* src = fw + job.rowbase + x;
      for (int t = 0; t < 32; ++t) {
        float2 v = make_float2(0.f, 0.f);
        if (t < T && x < nfx) v = __ldg(src + (long long)t * g.nftot);
        buf[t * 33 + lane] = v;
      }
    }
    __syncwarp();
    float2 acc[CX];
#pragma unroll
    for (int i = 0; i < CX; ++i) acc[i] = buf[lane * 33 + i];
    __syncwarp();

    auto load_visit = [&](int k) {
      Visit e;
      const int4 q = list[k];
      const float4* r = reinterpret_cast<const float4*>(rec + (long long)q.x * REC);
      e.off = q.y;
      e.wyz = __int_as_float(q.z);
      e.wa = __ldg(r);
      e.wb = __ldg(r + 1);
      e.v = make_float2(__int_as_float(q.x), 0.f);  // carries s
      return e;
    };
    for_each_row_visit<DIM, W>(g, job, bin_start, rec, list, lane, [&](int n) {
      if (n == 0) return;
      Visit cur = load_visit(0);
      for (int k = 0; k < n; ++k) {
        Visit nxt = cur;
        if (k + 1 < n) nxt = load_visit(k + 1);
        const float wx[8] = {cur.wa.x, cur.wa.y, cur.wa.z, cur.wa.w, cur.wb.x, cur.wb.y, cur.wb.z, 0.f};
        float2 p = make_float2(0.f, 0.f);
        switch (cur.off) {
#define CASE_(O) case O: p = taps_interp<W, O>(acc, wx); break;
          OFF_CASES(CASE_)
#undef CASE_
          default: break;
        }
        const int s = __float_as_int(cur.v.x);
        if (lane < T)
          atomicAdd(kt + (long long)s * 32 + lane, make_float2(p.x * cur.wyz, p.y * cur.wyz));
        cur = nxt;
      }
    });
  }
}

size_t row_smem() { return (size_t)WARPS * (32 * 33 * sizeof(float2) + LIST * sizeof(int4)); }

int ensure_state(b200_plan* p, int T, cudaStream_t st) {
  TiledState* ts = state(p);
  const long long M = p->M;
  if (!ts->d_counter) CUDA_TRY(cudaMalloc(&ts->d_counter, 64));
  if (!ts->rec_valid || ts->M != M) {
    if (ts->d_rec) cudaFree(ts->d_rec);
    ts->d_rec = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_rec, (size_t)(M > 0 ? M : 1) * REC * sizeof(float)));
    if (M > 0) {
      const int nb = ceil_div(M, 256);
#define LAUNCH_REC(W_)                                                                          \
  k_point_records<W_><<<nb, 256, 0, st>>>(p->g, M, p->d_poly, p->d_org_s[0], p->d_org_s[1],      \
                                          p->d_org_s[2], p->d_x1_s[0], p->d_x1_s[1],             \
                                          p->d_x1_s[2], ts->d_rec)
      switch (p->g.w) {
        case 4: LAUNCH_REC(4); break;
        case 5: LAUNCH_REC(5); break;
        case 6: LAUNCH_REC(6); break;
        default: LAUNCH_REC(7); break;
      }
#undef LAUNCH_REC
      CHECK_LAUNCH();
    }
    ts->M = M;
    ts->rec_valid = true;
  }
  const size_t need = (size_t)(M > 0 ? M : 1) * 32 * sizeof(float2);
  if (ts->kt_bytes < need) {
    if (ts->d_kt) cudaFree(ts->d_kt);
    ts->d_kt = nullptr;
    ts->kt_bytes = 0;
    CUDA_TRY(cudaMalloc(&ts->d_kt, need));
    ts->kt_bytes = need;
  }
  (void)T;
  return B200_OK;
}

template <int DIM, int W>
int launch_spread(b200_plan* p, TiledState* ts, float2* fw, int T, cudaStream_t st) {
  const long long nt = num_tiles<DIM>(p->g);
  auto kern = k_spread_rows<DIM, W>;
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem()));
    attr_done = true;
  }
  const int grid = (int)(nt < 2LL * B200_NUM_SMS ? nt : 2LL * B200_NUM_SMS);
  kern<<<grid, THREADS, row_smem(), st>>>(p->g, T, nt, p->d_bin_start, ts->d_rec, ts->d_kt, fw,
                                          ts->d_counter);
  CHECK_LAUNCH();
  return B200_OK;
}

template <int DIM, int W>
int launch_interp(b200_plan* p, TiledState* ts, const float2* fw, int T, cudaStream_t st) {
  const long long nt = num_tiles<DIM>(p->g);
  auto kern = k_interp_rows<DIM, W>;
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem()));
    attr_done = true;
  }
  const int grid = (int)(nt < 2LL * B200_NUM_SMS ? nt : 2LL * B200_NUM_SMS);
  kern<<<grid, THREADS, row_smem(), st>>>(p->g, T, nt, p->d_bin_start, ts->d_rec, fw, ts->d_kt,
                                          ts->d_counter);
  CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace

bool tiled_supported(const b200_plan* p, int T) {
  const Geom& g = p->g;
  if (g.dim < 2 || g.dim > 3) return false;
  if (g.w < 4 || g.w > 7) return false;
  if (T > 32) return false;
  const int nfx = g.nf[g.dim - 1];
  const int rem = nfx % CX;
  if (rem != 0 && rem < g.w - 1) return false;  // wrapped taps must come from the last bin only
  if (nfx < CX && nfx < 2 * g.w) return false;
  for (int a = 0; a < g.dim; ++a)
    if (g.nf[a] < 2 * g.w) return false;
  return true;
}

void tiled_free(b200_plan* p) {
  if (!p->tiled) return;
  TiledState* ts = (TiledState*)p->tiled;
  if (ts->d_rec) cudaFree(ts->d_rec);
  if (ts->d_kt) cudaFree(ts->d_kt);
  if (ts->d_counter) cudaFree(ts->d_counter);
  delete ts;
  p->tiled = nullptr;
}

#define DISPATCH_DW(FN, ...)                                   \
  do {                                                         \
    const int d_ = p->g.dim, w_ = p->g.w;                      \
    if (d_ == 3) {                                             \
      if (w_ == 7) return FN<3, 7>(__VA_ARGS__);               \
      if (w_ == 6) return FN<3, 6>(__VA_ARGS__);               \
      if (w_ == 5) return FN<3, 5>(__VA_ARGS__);               \
      return FN<3, 4>(__VA_ARGS__);                            \
    } else {                                                   \
      if (w_ == 7) return FN<2, 7>(__VA_ARGS__);               \
      if (w_ == 6) return FN<2, 6>(__VA_ARGS__);               \
      if (w_ == 5) return FN<2, 5>(__VA_ARGS__);               \
      return FN<2, 4>(__VA_ARGS__);                            \
    }                                                          \
  } while (0)

int spread_tiled(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                 cudaStream_t st) {
  B200_TRY(ensure_state(p, T, st));
  TiledState* ts = state(p);
  const long long M = p->M;
  if (M > 0) {
    k_gather_kspace<<<ceil_div(M * 32, 256), 256, 0, st>>>(M, T, p->d_perm, ksp, density, ts->d_kt);
    CHECK_LAUNCH();
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counter, 0, sizeof(int), st));
  DISPATCH_DW(launch_spread, p, ts, fw, T, st);
}

int interp_tiled(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                 const float2* obs, cudaStream_t st) {
  B200_TRY(ensure_state(p, T, st));
  TiledState* ts = state(p);
  const long long M = p->M;
  if (M == 0) return B200_OK;
  CUDA_TRY(cudaMemsetAsync(ts->d_kt, 0, (size_t)M * 32 * sizeof(float2), st));
  CUDA_TRY(cudaMemsetAsync(ts->d_counter, 0, sizeof(int), st));
  int rc = [&]() -> int { DISPATCH_DW(launch_interp, p, ts, fw, T, st); }();
  if (rc != B200_OK) return rc;
  k_scatter_kspace<<<ceil_div(M * 32, 256), 256, 0, st>>>(M, T, p->d_perm, ts->d_kt, ksp, scale, obs);
  CHECK_LAUNCH();
  return B200_OK;
}
