// double_rows.cu -- tile-owned, atomic-free spreading in float64 / complex128 for plans created with B200_DOUBLE:
// the register-tile idea of rows_common.cuh without its single-precision machinery.
//
// The point-driven double spreader (double_path.cu) issues 2 w^d global double atomics per point and coil and
// runs at the L2's atomic rate (2e11 / s measured: 52 ms for 2^21 points x 8 coils at w = 7, 20 x the
// single-precision path); shared-memory tiles do not help, shared double atomics are compare-and-swap loops on
// sm_100a (108 .. 220 ms, tools/bench_double.py).  Here a warp OWNS a tile of 2 rows x 16 cells at fixed z for up
// to 16 coils:
//
//     lane = (x-half h, coil t):  32 complex128 accumulators = rows y, y + 1 x cells 8 h .. 8 h + 7 of coil t
//
// The visitors of a tile are the same contiguous ranges of bin-sorted points as in single precision (the fold
// runs in double, the sort on the integer origins: setpts.cu `k1_sort_origins`), written out once per
// trajectory as a visit stream {x offset, dy, dz, sorted point}.  Entries are staged 32 at a time with
// asynchronous copies, one block ahead of the arithmetic: per visit the x weights laid over the tile's 16 cells
// (zero-filled outside the footprint, so consuming a visit needs no case analysis and is plain C++: no generated
// PTX here), the y / z weights of the two rows, and the point's 16 coil values.  Consuming a visit is 6
// shared-memory loads, 4 multiplies and 32 DFMA per lane: the kernel is bound by the FP64 pipe.  A finished tile
// leaves with plain 128-bit stores (a lane's 8 cells of a row are one 128-byte line) -- every cell of the grid is
// written, the caller does not clear it; tiles cut by a chunk boundary are added with double atomics to
// pre-zeroed rows.  Any kernel width up to 16 (eps down to 1e-14), 2-D and 3-D.  Interpolation stays
// point-driven (a gather, no atomics).
//
// Replaces finufft's double-precision spreading stage (`dtype = samples.dtype`,
// src/mrinufft/operators/base.py:934; call site src/mrinufft/operators/interfaces/finufft.py:76).
#include "rows_common.cuh"

using namespace rows;

namespace {

constexpr int WT = 16;            // weights per axis in the point table
constexpr int DTC = 16;           // coils per call of the kernel
constexpr unsigned DHDR = 0xffffffffu;  // entry.z of a header entry (and of the all-ones padding)

struct DRowsState {
  double* d_wtab = nullptr;        // [M][3][WT] weights of the sorted points: x, y, z axis roles
  double2* d_kt = nullptr;         // [M][DTC] (sorted point, coil) values
  int32_t* d_tot = nullptr;
  uint32_t* d_start = nullptr;
  uint4* d_ent = nullptr;          // visit stream: {x offset, dy, dz, sorted point} / header {0, 0, DHDR, tile}
  int32_t* d_chunk_row = nullptr;
  int32_t* d_split_rows = nullptr;
  int* d_counters = nullptr;
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0, pts_cap = 0, ent_cap = 0, chunk_cap = 0;
  long long nrows = 0, visits = 0, nsplit = 0;
  unsigned S = 0;
  int nchunks = 0, lch = LCH;
  bool valid = false, unsupported = false;
};

DRowsState* dstate(b200_plan* p) {
  if (!p->drows) p->drows = new DRowsState();
  return (DRowsState*)p->drows;
}

__device__ __forceinline__ double es_phi(double x, double hw, double beta) {
  const double r = x / hw;
  const double a = 1.0 - r * r;
  return a < 0.0 ? 0.0 : exp(beta * (sqrt(a) - 1.0));
}

// weight table of the sorted points (direct evaluation of the kernel in double)
__global__ void __launch_bounds__(256)
kd_weights(Geom g, double beta, long long M, const int32_t* __restrict__ perm, const double* __restrict__ x0,
           const double* __restrict__ x1, const double* __restrict__ x2, double* __restrict__ wtab) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= M) return;
  const int j = perm[s];
  const double* xs[3] = {x0, x1, x2};
  const double hw = 0.5 * g.w;
  for (int r = 0; r < 3; ++r) {  // role r = 0: fastest axis (x), 1: y, 2: z
    const int a = g.dim - 1 - r;
    double* out = wtab + ((long long)s * 3 + r) * WT;
    for (int i = 0; i < WT; ++i) {
      double v = 0.0;
      if (a < 0) v = i == 0 ? 1.0 : 0.0;
      else if (i < g.w) v = es_phi(xs[a][j] + i, hw, beta);
      out[i] = v;
    }
  }
}

// ranges of sorted points that visit tile `rc` (run-time kernel width; cf. rows::slot_range)
template <int DIM>
__device__ __forceinline__ void slot_range_rt(const Geom& g, const TileCoord& rc, int slot, int w,
                                              const int32_t* __restrict__ bin_start, int* begin, int* len) {
  const int nzs = DIM == 3 ? w : 1;
  *begin = 0;
  *len = 0;
  if (slot >= nzs * 6) return;
  const int part = slot & 1, sub = (slot >> 1) % 3, zs = (slot >> 1) / 3;
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int ylo = rc.y - (w - 1);
  int a, b;
  if (part == 0) {
    a = ylo > 0 ? ylo : 0;
    b = rc.y + 1;
  } else {
    if (ylo >= 0) return;
    a = ylo + nfy;
    b = nfy - 1;
  }
  int z0 = 0;
  if (DIM == 3) {
    z0 = rc.z - zs;
    if (z0 < 0) z0 += g.nf[0];
  }
  const int bxx = (sub == 2) ? (rc.bx == 0 ? nbx - 1 : rc.bx - 1) : rc.bx;
  const long long kb = (((long long)z0 * nbx + bxx) * 2 + (sub != 0 ? 1 : 0)) * nfy;
  const int s0 = __ldg(bin_start + kb + a);
  *begin = s0;
  *len = __ldg(bin_start + kb + b + 1) - s0;
}

template <int DIM>
__global__ void __launch_bounds__(256)
kd_row_totals(Geom g, long long nrows, const int32_t* __restrict__ bin_start, int32_t* __restrict__ tot,
              unsigned long long* __restrict__ grand) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = 0;
  TileCoord rc;
  const bool ok = row < nrows && decode_tile<DIM, 32>(g, row, &rc);
  if (ok) {
    const int ns = (DIM == 3 ? g.w : 1) * 6;
    for (int slot = 0; slot < ns; ++slot) {
      int b, l;
      slot_range_rt<DIM>(g, rc, slot, g.w, bin_start, &b, &l);
      total += l;
    }
  }
  if (row <= nrows) tot[row] = ok ? (int32_t)min(total, (long long)INT32_MAX) : -1;
  long long wsum = total;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, d);
  if ((threadIdx.x & 31) == 0 && wsum > 0) atomicAdd(grand, (unsigned long long)wsum);
}

__global__ void __launch_bounds__(256)
kd_scan_inputs(long long n, const int32_t* __restrict__ tot, uint32_t* __restrict__ words) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tot[i];
  words[i] = t < 0 ? 0u : (uint32_t)t + 1u;
}

// one warp per tile: header entry + one entry per visit, in range order
template <int DIM>
__global__ void __launch_bounds__(256)
kd_build_stream(Geom g, long long nrows, const int32_t* __restrict__ bin_start, const int32_t* __restrict__ tot,
                const uint32_t* __restrict__ start, const int32_t* __restrict__ ox, const int32_t* __restrict__ oy,
                uint4* __restrict__ ent, int32_t* __restrict__ chunk_row, int32_t* __restrict__ split_rows,
                int* __restrict__ split_counter, uint32_t lch) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int t = tot[row];
  if (t < 0) return;
  const uint32_t hs = start[row], he = hs + 1u + (uint32_t)t;
  if (lane == 0) {
    ent[hs] = make_uint4(0u, 0u, DHDR, (uint32_t)row);
    for (uint32_t c = (hs + lch - 1) / lch; c * lch < he; ++c) chunk_row[c] = (int32_t)row;
    if (hs / lch != (he - 1) / lch) split_rows[atomicAdd(split_counter, 1)] = (int32_t)row;
  }
  if (t == 0) return;
  TileCoord rc;
  decode_tile<DIM, 32>(g, row, &rc);
  const int nfx = g.nf[DIM - 1], nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int left_len = (rc.bx == 0) ? (nfx - (nbx - 1) * CX) : CX;  // length of the left neighbour tile
  uint4* out = ent + hs + 1;
  auto entry = [&](int s, int zs, int left) -> uint4 {
    const int xo = ox[s] % CX - left * left_len;  // x offset of the footprint inside this tile (may be < 0)
    int dy = rc.y - oy[s];
    if (dy < -1) dy += nfy;
    return make_uint4((unsigned)xo, (unsigned)dy, (unsigned)zs, (unsigned)s);
  };
  const int ns = (DIM == 3 ? g.w : 1) * 6;
  int run = 0;
  for (int h = 0; h * 32 < ns; ++h) {
    const int slot = lane + 32 * h;
    int b, l;
    slot_range_rt<DIM>(g, rc, slot, g.w, bin_start, &b, &l);
    int inc = l;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += u;
    }
    const int pre = run + inc - l;
    run += __shfl_sync(0xffffffffu, inc, 31);
    const int zs = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
    // short ranges: written by the owning lane; long ranges (dense k-space centre): by the whole warp
    const bool is_long = l > 64;
    if (!is_long)
      for (int i = 0; i < l; ++i) out[pre + i] = entry(b + i, zs, left);
    unsigned longs = __ballot_sync(0xffffffffu, is_long);
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const int bb = __shfl_sync(0xffffffffu, b, src), ll = __shfl_sync(0xffffffffu, l, src);
      const int pp = __shfl_sync(0xffffffffu, pre, src);
      const int szs = __shfl_sync(0xffffffffu, zs, src), sl = __shfl_sync(0xffffffffu, left, src);
      for (int i = lane; i < ll; i += 32) out[pp + i] = entry(bb + i, szs, sl);
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(128)
kd_zero_split_rows(Geom g, int T, long long nsplit, const int32_t* __restrict__ split_rows, double2* __restrict__ fw) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nsplit) return;
  TileCoord rc;
  if (!decode_tile<DIM, 32>(g, split_rows[w], &rc)) return;
  const int nfx = g.nf[DIM - 1], nfy = g.nf[DIM - 2];
  const int x = rc.bx * CX + (lane & 15);
  if (x >= nfx) return;
  double2* dst = fw + ((long long)rc.z * nfy + rc.y + (lane >> 4)) * nfx + x;
  for (int t = 0; t < T; ++t) dst[(long long)t * g.nftot] = make_double2(0.0, 0.0);
}

// kt[s][t] = ksp[t][perm[s]] * density   (t < T, zero for T <= t < DTC)
__global__ void __launch_bounds__(256)
kd_gather(long long M, int T, const int32_t* __restrict__ perm, const double2* __restrict__ ksp,
          const double* __restrict__ density, double2* __restrict__ kt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * DTC) return;
  const long long s = i / DTC;
  const int t = (int)(i % DTC);
  double2 v = make_double2(0.0, 0.0);
  if (t < T) {
    const int j = perm[s];
    v = ksp[(long long)t * M + j];
    if (density) {
      const double d = density[j];
      v.x *= d;
      v.y *= d;
    }
  }
  kt[i] = v;
}

// 8- and 16-byte asynchronous copies to shared memory, `n` bytes read and the rest zero-filled
__device__ __forceinline__ void cp_async8_zfill(unsigned smem_dst, const void* gsrc, int n) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(unsigned smem_dst, const void* gsrc, int n) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(n) : "memory");
}

// per-warp shared memory: two buffers of NB staged visits
constexpr int NB = 16;                  // visits per staged block
constexpr int DCTAS = 3;                // CTAs per SM the row kernel is compiled for
constexpr int DSM_E = NB * 16 * 8;      // cell weights: [visit][16 cells]
constexpr int DSM_V = NB * DTC * 16;    // coil values: [visit][DTC] complex128
constexpr int DSM_R = NB * 32;          // [visit] {wy[dy], wy[dy + 1], wz[dz], -} -> row scales (s0, s1)
constexpr int DSM_BUF = DSM_E + DSM_V + DSM_R;
constexpr int DSM_WARP = 2 * DSM_BUF;

template <int DIM>
__global__ void __launch_bounds__(THREADS, DCTAS)
kd_rows(Geom g, int T, int nchunks, unsigned S, const uint4* __restrict__ ent, const int32_t* __restrict__ chunk_row,
        const double* __restrict__ wtab, const double2* __restrict__ kt, double2* __restrict__ fw,
        int* __restrict__ counter, int lch_log2) {
  extern __shared__ __align__(128) unsigned char dsm_raw[];
  constexpr unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xh = lane >> 4, t = lane & 15;
  unsigned char* wsm = dsm_raw + (size_t)warp * DSM_WARP;
  const unsigned wsa = smem_u32(wsm);
  const int nfx = g.nf[DIM - 1], nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int w = g.w;
  const bool coil_on = t < T;
  const uint4 pad = make_uint4(0u, 0u, DHDR, 0xffffffffu);

  // accumulators: [row][cell] complex128 of this lane's coil, cells 8 xh .. 8 xh + 7
  double2 acc[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = make_double2(0.0, 0.0);

  for (;;) {
    int ch = 0;
    if (lane == 0) ch = atomicAdd(counter, 1);
    ch = __shfl_sync(FULL, ch, 0);
    if (ch >= nchunks) break;
    const unsigned base = (unsigned)ch << lch_log2;
    const uint4* v = ent + base;
    const int nw = (int)min(1u << lch_log2, S - base);
    const int nblk = (nw + NB - 1) / NB;
    const bool tail_whole = __ldg(reinterpret_cast<const unsigned*>(v + nw) + 2) == DHDR;

    double2* tbase = nullptr;  // this lane's first cell of row 0 of the tile
    int xlim = 0;              // this lane's cells inside the grid (8, fewer in a short last tile)
    auto tile_setup = [&](int row) {
      const int nz = DIM == 3 ? g.nf[0] : 1;
      const int ps = row & 3;
      int r = row >> 2;
      const int bx = r % nbx;
      r /= nbx;
      const int pg = r % 4;
      r /= 4;
      const int z = r % nz;
      const int yb = r / nz;
      const int y = yb * YB + (pg * 4 + ps) * 2;
      xlim = nfx - bx * CX - xh * 8;
      xlim = xlim > 8 ? 8 : xlim;
      tbase = fw + (long long)t * g.nftot + ((long long)z * nfy + y) * nfx + bx * CX + xh * 8;
    };
    tile_setup(__ldg(chunk_row + ch));
    bool started = false, dirty = false;

    auto flush = [&](bool shared) {
      if (!coil_on) return;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < xlim) {
            double2* a = tbase + (long long)r * nfx + c;
            if (!shared) {
              __stcs(a, acc[r][c]);
            } else {
              atomicAdd(&a->x, acc[r][c].x);
              atomicAdd(&a->y, acc[r][c].y);
            }
          }
    };
    auto store_zero = [&]() {
      if (!coil_on) return;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < xlim) __stcs(tbase + (long long)r * nfx + c, make_double2(0.0, 0.0));
    };
    auto load_entries = [&](int blk) -> uint4 {
      const int i = blk * NB + lane;
      return (lane < NB && i < nw) ? __ldg(v + i) : pad;
    };
    // Stage a block of NB entries, all copies asynchronous.  A lane fetches the y / z weights of its own entry;
    // the 16 cell weights (the x weights laid over the tile's cells, zero outside the footprint) and the 16
    // coil values of a visit are fetched by a half-warp, two visits per step (coalesced, conflict-free).
    auto issue = [&](int buf, const uint4& e) {
      const unsigned sE = wsa + buf * DSM_BUF, sV = sE + DSM_E, sR = sV + DSM_V;
      const bool hdr = e.z == DHDR;
      const double* wt = wtab + (hdr ? 0LL : (long long)e.w * (3 * WT));
      const int dy = (int)e.y;
      const bool ok0 = !hdr && dy >= 0 && dy < w, ok1 = !hdr && dy + 1 < w;
      if (lane < NB) {
        cp_async8_zfill(sR + lane * 32, wt + WT + (ok0 ? dy : 0), ok0 ? 8 : 0);
        cp_async8_zfill(sR + lane * 32 + 8, wt + WT + (ok1 ? dy + 1 : 0), ok1 ? 8 : 0);
        if (DIM == 3) cp_async8_zfill(sR + lane * 32 + 16, wt + 2 * WT + (hdr ? 0 : (int)e.z), hdr ? 0 : 8);
      }
#pragma unroll 4
      for (int q = 0; q < NB / 2; ++q) {
        const int vi = 2 * q + xh;
        const unsigned pw = __shfl_sync(FULL, e.w, vi);
        const int pxo = (int)__shfl_sync(FULL, e.x, vi);
        const bool ph = __shfl_sync(FULL, e.z, vi) == DHDR;
        const int k = t - pxo;
        const bool ok = !ph && k >= 0 && k < w;
        // 32-bit element indices: M * 48 < 2^32 is a condition of `build`
        cp_async8_zfill(sE + (vi * 16 + t) * 8, wtab + (ok ? pw * (3u * WT) + (unsigned)k : 0u), ok ? 8 : 0);
        cp_async16_zfill(sV + (vi * DTC + t) * 16, kt + (ph ? 0u : pw * (unsigned)DTC + (unsigned)t), ph ? 0 : 16);
      }
      cp_async_commit();
    };

    uint4 e0 = load_entries(0), e1 = nblk > 1 ? load_entries(1) : pad;
    issue(0, e0);
    for (int blk = 0; blk < nblk; ++blk) {
      const int buf = blk & 1;
      uint4 e2 = pad;
      if (blk + 1 < nblk) {
        issue(buf ^ 1, e1);
        if (blk + 2 < nblk) e2 = load_entries(blk + 2);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const unsigned char* bp = wsm + buf * DSM_BUF;
      const double* sE = reinterpret_cast<const double*>(bp);
      const double2* sV = reinterpret_cast<const double2*>(bp + DSM_E);
      double2* sR = reinterpret_cast<double2*>(const_cast<unsigned char*>(bp) + DSM_E + DSM_V);
      if (DIM == 3) {  // row scales of this lane's entry: wy[dy] wz[dz], wy[dy + 1] wz[dz]
        if (lane < NB) {
          const double2 wy = sR[lane * 2];
          const double wz = sR[lane * 2 + 1].x;
          sR[lane * 2] = make_double2(wy.x * wz, wy.y * wz);
        }
        __syncwarp();
      }
      const int n = min(NB, nw - blk * NB);
      // ---- consume: runs of visits separated by header entries
      unsigned hm = __ballot_sync(FULL, e0.z == DHDR && e0.w != 0xffffffffu);
      int k0 = 0;
      for (;;) {
        const int k1 = hm ? (__ffs(hm) - 1) : n;
        if (k1 > k0) {
          dirty = true;
#pragma unroll 2
          for (int k = k0; k < k1; ++k) {
            const double2 sc = sR[k * 2];
            const double2 val = sV[k * DTC + t];
            double ew[8];
            const double2* ep = reinterpret_cast<const double2*>(sE + k * 16 + xh * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const double2 p2 = ep[q];
              ew[2 * q] = p2.x;
              ew[2 * q + 1] = p2.y;
            }
            const double2 a0 = make_double2(val.x * sc.x, val.y * sc.x);
            const double2 a1 = make_double2(val.x * sc.y, val.y * sc.y);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              acc[0][c].x = fma(a0.x, ew[c], acc[0][c].x);
              acc[0][c].y = fma(a0.y, ew[c], acc[0][c].y);
              acc[1][c].x = fma(a1.x, ew[c], acc[1][c].x);
              acc[1][c].y = fma(a1.y, ew[c], acc[1][c].y);
            }
          }
        }
        const bool at_end = k1 >= n;
        if (at_end && blk + 1 < nblk) break;
        // the tile ends here (next header) or the chunk does
        const bool whole = started && (!at_end || tail_whole);
        if (dirty) {
          flush(!whole);
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = make_double2(0.0, 0.0);
        } else if (whole) {
          store_zero();
        }
        if (at_end) break;
        hm &= hm - 1;
        tile_setup((int)__shfl_sync(FULL, e0.w, k1));
        started = true;
        dirty = false;
        k0 = k1 + 1;
      }
      __syncwarp();
      e0 = e1;
      e1 = e2;
    }
  }
}

template <int DIM>
int build(b200_plan* p, DRowsState* ds, const double* const* x1u, cudaStream_t st) {
  const long long M = p->M;
  const Geom& g = p->g;
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  ds->valid = false;
  ds->unsupported = false;
  if (!ds->d_counters) CUDA_TRY(cudaMalloc(&ds->d_counters, 64));
  if ((size_t)M > ds->pts_cap || !ds->d_wtab) {
    fr(ds->d_wtab);
    fr(ds->d_kt);
    ds->d_wtab = nullptr;
    ds->d_kt = nullptr;
    ds->pts_cap = 0;
    const size_t cap = (size_t)(M > 0 ? M : 1);
    if (cudaMalloc(&ds->d_wtab, cap * 3 * WT * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&ds->d_kt, cap * DTC * sizeof(double2)) != cudaSuccess) {
      cudaGetLastError();
      ds->unsupported = true;  // not enough memory for the tables: the point-driven kernels serve the plan
      return 1;
    }
    ds->pts_cap = cap;
  }
  B200_TRY(k1_sort_origins(p, st));
  if (M > 0) {
    kd_weights<<<ceil_div(M, 256), 256, 0, st>>>(g, p->beta, M, p->d_perm, x1u[0], x1u[1], x1u[2], ds->d_wtab);
    CHECK_LAUNCH();
  }
  const long long nrows = num_tiles<DIM, 32>(g);
  if (nrows >= (1LL << 30) || M >= (1LL << 32) / (3 * WT)) {
    ds->unsupported = true;
    return 1;
  }
  if (ds->nrows != nrows || !ds->d_tot) {
    fr(ds->d_tot);
    fr(ds->d_start);
    ds->d_tot = nullptr;
    ds->d_start = nullptr;
    CUDA_TRY(cudaMalloc(&ds->d_tot, (size_t)(nrows + 1) * 4));
    CUDA_TRY(cudaMalloc(&ds->d_start, (size_t)(nrows + 1) * 4));
    ds->nrows = nrows;
  }
  CUDA_TRY(cudaMemsetAsync(ds->d_counters, 0, 64, st));
  kd_row_totals<DIM><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(
      g, nrows, p->d_bin_start, ds->d_tot, reinterpret_cast<unsigned long long*>(ds->d_counters + 2));
  CHECK_LAUNCH();
  unsigned long long grand = 0;
  CUDA_TRY(cudaMemcpyAsync(&grand, ds->d_counters + 2, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (grand + (unsigned long long)nrows >= (1ULL << 31)) {
    ds->unsupported = true;
    return 1;
  }
  kd_scan_inputs<<<ceil_div(nrows + 1, 256), 256, 0, st>>>(nrows + 1, ds->d_tot, ds->d_start);
  CHECK_LAUNCH();
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, ds->d_start, ds->d_start, (int)(nrows + 1), st);
  if (need > ds->scan_tmp_bytes) {
    fr(ds->d_scan_tmp);
    ds->d_scan_tmp = nullptr;
    ds->scan_tmp_bytes = 0;
    CUDA_TRY(cudaMalloc(&ds->d_scan_tmp, need));
    ds->scan_tmp_bytes = need;
  }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ds->d_scan_tmp, need, ds->d_start, ds->d_start, (int)(nrows + 1), st));
  g_kernel_launches += 2;
  uint32_t S = 0;
  CUDA_TRY(cudaMemcpyAsync(&S, ds->d_start + nrows, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  int lch = LCH;
  while (lch > 128 && (long long)S / lch < 8LL * p->num_sms * 4 * WARPS) lch >>= 1;
  ds->lch = lch;
  const int nchunks = (int)((S + lch - 1) / lch);
  if ((size_t)S + 64 > ds->ent_cap) {
    fr(ds->d_ent);
    ds->d_ent = nullptr;
    ds->ent_cap = 0;
    const size_t cap = (size_t)S + 64 + (size_t)S / 16;
    if (cudaMalloc(&ds->d_ent, cap * sizeof(uint4)) != cudaSuccess) {
      cudaGetLastError();
      ds->unsupported = true;
      return 1;
    }
    ds->ent_cap = cap;
  }
  if ((size_t)nchunks + 1 > ds->chunk_cap) {
    fr(ds->d_chunk_row);
    fr(ds->d_split_rows);
    ds->d_chunk_row = nullptr;
    ds->d_split_rows = nullptr;
    ds->chunk_cap = 0;
    const size_t cap = (size_t)nchunks + 1 + (size_t)nchunks / 16;
    CUDA_TRY(cudaMalloc(&ds->d_chunk_row, cap * 4));
    CUDA_TRY(cudaMalloc(&ds->d_split_rows, cap * 4));
    ds->chunk_cap = cap;
  }
  CUDA_TRY(cudaMemsetAsync(ds->d_ent, 0xff, ((size_t)S + 64) * sizeof(uint4), st));
  CUDA_TRY(cudaMemsetAsync(ds->d_chunk_row, 0, (size_t)(nchunks + 1) * 4, st));
  kd_build_stream<DIM><<<ceil_div(nrows * 32, 256), 256, 0, st>>>(
      g, nrows, p->d_bin_start, ds->d_tot, ds->d_start, p->d_org_s[DIM - 1], p->d_org_s[DIM - 2], ds->d_ent,
      ds->d_chunk_row, ds->d_split_rows, ds->d_counters + 1, (uint32_t)lch);
  CHECK_LAUNCH();
  int nsplit = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsplit, ds->d_counters + 1, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  ds->S = S;
  ds->visits = (long long)grand;
  ds->nchunks = nchunks;
  ds->nsplit = nsplit;
  ds->valid = true;
  return B200_OK;
}

template <int DIM>
int launch(b200_plan* p, DRowsState* ds, double2* fw, int T, cudaStream_t st) {
  auto kern = kd_rows<DIM>;
  const size_t smem = (size_t)WARPS * DSM_WARP;
  static PerDeviceOnce once;
  static int ctas_per_sm = 1;
  if (once.first()) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, THREADS, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  if (ds->nsplit > 0) {
    kd_zero_split_rows<DIM><<<ceil_div(ds->nsplit * 32, 128), 128, 0, st>>>(p->g, T, ds->nsplit, ds->d_split_rows, fw);
    CHECK_LAUNCH();
  }
  CUDA_TRY(cudaMemsetAsync(ds->d_counters, 0, sizeof(int), st));
  const long long want = (ds->nchunks + WARPS - 1) / WARPS;
  const long long cap = (long long)p->num_sms * ctas_per_sm;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  kern<<<grid, THREADS, smem, st>>>(p->g, T, ds->nchunks, ds->S, ds->d_ent, ds->d_chunk_row, ds->d_wtab, ds->d_kt, fw,
                                    ds->d_counters, 31 - __builtin_clz(ds->lch));
  CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace

bool drows_supported(const b200_plan* p) {
  const Geom& g = p->g;
  if (g.dim < 2 || g.dim > 3 || g.w > WT) return false;
  if (p->rows_dbg & 32) return false;  // option 3, bit 5: the point-driven double kernels (A/B, tests)
  const int rem = g.nf[g.dim - 1] % CX;
  if (rem != 0 && rem < g.w - 1) return false;  // a footprint may touch at most two tiles
  if (g.w - 1 > CX) return false;
  for (int a = 0; a < g.dim; ++a)
    if (g.nf[a] < 2 * g.w) return false;
  return true;
}

int drows_setpts(b200_plan* p, const double* const* x1u, cudaStream_t st) {
  DRowsState* ds = dstate(p);
  if (p->g.dim == 3) return build<3>(p, ds, x1u, st);
  return build<2>(p, ds, x1u, st);
}

int drows_spread(b200_plan* p, const double2* ksp, const double* density, double2* fw, int T, cudaStream_t st) {
  DRowsState* ds = dstate(p);
  if (!ds->valid || ds->unsupported) return 1;
  const long long M = p->M;
  for (int t0 = 0; t0 < T; t0 += DTC) {
    const int tn = T - t0 < DTC ? T - t0 : DTC;
    if (M > 0) {
      kd_gather<<<ceil_div(M * DTC, 256), 256, 0, st>>>(M, tn, p->d_perm, ksp + (long long)t0 * M, density, ds->d_kt);
      CHECK_LAUNCH();
    }
    double2* fwt = fw + (long long)t0 * p->g.nftot;
    if (p->g.dim == 3) B200_TRY(launch<3>(p, ds, fwt, tn, st));
    else B200_TRY(launch<2>(p, ds, fwt, tn, st));
  }
  return B200_OK;
}

void drows_info(const b200_plan* p, int64_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  const DRowsState* ds = (const DRowsState*)p->drows;
  if (!ds || !drows_supported(p)) return;
  out[3] = ds->unsupported ? 1 : 0;
  if (!ds->valid) return;
  out[0] = DTC;
  out[1] = ds->visits;
  out[2] = ds->S;
}

void drows_free(b200_plan* p) {
  if (!p->drows) return;
  DRowsState* ds = (DRowsState*)p->drows;
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  fr(ds->d_wtab);
  fr(ds->d_kt);
  fr(ds->d_tot);
  fr(ds->d_start);
  fr(ds->d_ent);
  fr(ds->d_chunk_row);
  fr(ds->d_split_rows);
  fr(ds->d_counters);
  fr(ds->d_scan_tmp);
  delete ds;
  p->drows = nullptr;
}
