// spread_rows.cu -- "method 2": output-owned, register-accumulating spreading (K2) and its
// transpose for interpolation (K3), written for sm_100a.
//
// Why not the usual shared-memory sub-grid + atomicAdd design: a 3-D width-7 kernel needs 343
// complex accumulations per point per coil; with 32 coils batched that is 1.8e11 float atomics per
// transform, and on sm_100a a float atomicAdd on shared memory is an ATOMS.CAST.SPIN
// compare-and-swap loop (checked with cuobjdump).  Shared memory cannot feed the FMA pipe either
// (128 B/clk/SM against 128 FFMA/clk/SM).  The register file can.  So every fine-grid TILE
//
//        2 rows (y, y+1) x 16 consecutive cells along the fastest axis at fixed z, all coils
//
// is OWNED by one warp at a time and lives in REGISTERS: lane = coil, 32 64-bit registers hold the
// (cell 2j, cell 2j+1) pairs of the real and imaginary parts of both rows.  The points whose
// footprint covers the tile ("visits") are found through the bin sort of K1 without any search:
//
//   * bins are pencils of 1 x 1 x 16 cells, split in two sub-bins: points whose footprint stays
//     inside the 16-cell tile ("interior") and points whose footprint crosses into the next tile
//     ("crossing"); key order (z0, x-tile, crossing, y0);
//   * the visits of tile (z, y..y+1, bx) are therefore exactly 3 contiguous ranges of sorted points
//     per slow-axis offset dz (own interior, own crossing, left neighbour's crossing), each spanning
//     y0 in [y-w+1, y+1] (twice that when the range wraps periodically);
//   * the ranges of a tile are concatenated by a warp prefix sum; blocks of 16 visits are staged
//     lane-parallel (one lane = one visit: load the point's weight record, pair-pack the x weights
//     by the parity of the x offset, form the two row scales wy[dy] wz[dz], write a 48-byte packet
//     to shared memory) while the sample values of those 16 points for all coils stream in with
//     cp.async (one coalesced 256-byte row per point from the (sorted point, coil) transposed
//     k-space batch), double buffered;
//   * the consume loop costs 3 LDS.128 (packet, warp broadcast) + 1 LDS.64 (this lane's coil
//     value) + 8 FMUL + one indexed branch on the x offset + 16 packed FFMA2 per visit, for two
//     grid rows: the weights broadcast -- the scarce resource, one L1 wavefront per clock per SM --
//     is amortised over twice the FMAs of a one-row design;
//   * a tile is written to HBM exactly once as 128-byte coalesced stores per coil and row (through
//     a shared-memory transpose): no memset of the oversampled grid, no halo flush, no atomics.
//
// Load balance: a trajectory like 3-D radial puts ~1e5 visits on the few tiles through the k-space
// centre.  Tiles with more than CHUNK visits are split into several work items whose partial
// results are merged with vector red.global.add on pre-zeroed rows; all items are handed out
// dynamically (groups of 4 vertically adjacent tiles per atomic fetch, so a warp re-uses the point
// records it just pulled into L1), in an order that sweeps z inside slabs of 32 rows so that a
// point's data is still in L2 when the next plane needs it.
//
// Interpolation is the exact transpose: the warp loads its tile into registers once (coalesced),
// every visiting point takes its tap dot products from registers and adds the partial sum into
// the (sorted point, coil) accumulator with one vector `red.global.add.v2.f32` per lane.
//
// Replaces finufft's spread/interp stage (call sites
// src/mrinufft/operators/interfaces/finufft.py:69,76; algorithm docs/explanations/nufft.rst:253-309).
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "device_utils.cuh"

namespace {

typedef unsigned long long u64;

constexpr int CX = 16;          // cells per tile row (== pencil-bin width, B200_BIN_X)
constexpr int NACC = 32;        // 64-bit accumulator registers per lane: [row 2][re/im 2][cell pair 8]
constexpr int REC = 28;         // floats per point record (112 B):
constexpr int R_WY = 8;         //   [0..7]  pair-packed x weights P | [8..16] 0, wy[0..6], 0
constexpr int R_WZ = 17;        //   [17..23] wz[0..6] | [24] (xo >> 1) + W/2 | [25] y0 | [26..27] pad
constexpr int R_JY = 24;
constexpr int VBLK = 16;        // visits per value block (cp.async / bulk-copy granularity)
constexpr int MBLK = 32;        // visits per packet block (one lane stages one visit)
constexpr int WARPS = 4;        // warps per CTA
constexpr int THREADS = WARPS * 32;
constexpr int CHUNK = 4096;     // visits per work item
constexpr int GROUP = 4;        // work items fetched per atomic

constexpr int ITEM_EMPTY = 1;   // item flags: tile without visits
constexpr int ITEM_SPLIT = 2;   //             tile shared by several items

// visit word: sorted point index | dz << 27 | (visit comes from the left neighbour's crossing bin) << 30
constexpr unsigned VIS_SBITS = 27;
constexpr unsigned VIS_SMASK = (1u << VIS_SBITS) - 1;
constexpr unsigned VIS_NONE = 0xffffffffu;

// per-warp shared memory (bytes)
constexpr int SM_VBUF = 2 * VBLK * 32 * 8;  // double-buffered coil values of VBLK points
constexpr int SM_META = 2 * MBLK * 48;      // double-buffered packets {P0..P3, s0, s1, idx, s}
constexpr int SM_MBAR = 16;                 // two mbarriers (bulk-copy variant)
constexpr int SM_WARP = SM_VBUF + SM_META + SM_MBAR;
constexpr int TS = 34;          // row stride (floats) of the transpose planes
static_assert(SM_VBUF + SM_META >= 2 * 32 * TS * 4, "transpose buffers must fit in vbuf+meta");

struct RowsState {
  float* d_rec = nullptr;        // [M][REC] per sorted point
  float2* d_kt = nullptr;        // [M][32] transposed (sorted point, coil) k-space batch
  size_t kt_bytes = 0;
  int32_t* d_iperm = nullptr;    // [M] point index -> sorted position
  int32_t* d_tot = nullptr;      // [nrows + 1] visits per tile
  uint32_t* d_vis_start = nullptr;  // [nrows + 1] first visit word of a tile
  int32_t* d_item_start = nullptr;  // [nrows + 1]
  int4* d_items = nullptr;       // [nitems] {tile, flags, first visit word, visit count}
  uint32_t* d_vis = nullptr;     // [nvis] visit words, tile by tile
  int32_t* d_split_rows = nullptr;
  int* d_counters = nullptr;     // [0] work counter, [1] split-row counter, [2..3] total visits (u64)
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  long long nrows = 0, nitems = 0, nsplit = 0, nvis = 0;
  long long M = -1;
  bool valid = false;
  bool unsupported = false;      // too many visits / points for the 32-bit visit words
};

RowsState* state(b200_plan* p) {
  if (!p->tiled) p->tiled = new RowsState();
  return (RowsState*)p->tiled;
}

// ------------------------------------------------------------------------------ geometry helpers
template <int DIM>
__host__ __device__ __forceinline__ int num_xtiles(const Geom& g) {
  return (g.nf[DIM - 1] + CX - 1) / CX;
}

// Tile ids enumerate (y-block of YB rows, z, group of 4 row pairs inside the block, x-tile, pair
// inside the group): the GROUP = 4 ids fetched together are 4 vertically adjacent tiles, and the
// sweep over z stays inside a slab of YB rows, so that the coil rows / records of the points
// (re-visited by the next w - 1 planes) are still in L2: one z step streams
// YB * nfx * 8 B * T = 4 MB of grid, not a whole 67 MB plane.
constexpr int YB = 32;

template <int DIM>
__host__ __device__ __forceinline__ long long num_rows(const Geom& g) {
  const long long nyb = (g.nf[DIM - 2] + YB - 1) / YB;
  const long long nz = DIM == 3 ? g.nf[0] : 1;
  return nyb * nz * (YB / 8) * num_xtiles<DIM>(g) * 4;
}

struct RowCoord {
  int z, y, bx;       // y = first (even) row of the pair
  long long rowbase;  // linear index of (z, y, 0) in one coil's grid
};

template <int DIM>
__device__ __forceinline__ bool decode_row(const Geom& g, long long row, RowCoord* rc) {
  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int nz = DIM == 3 ? g.nf[0] : 1;
  const int ps = (int)(row & 3);
  long long r = row >> 2;
  rc->bx = (int)(r % nbx);
  r /= nbx;
  const int pg = (int)(r % (YB / 8));
  r /= (YB / 8);
  rc->z = (int)(r % nz);
  const int yb = (int)(r / nz);
  rc->y = yb * YB + pg * 8 + ps * 2;
  if (rc->y >= nfy) return false;  // nfy is even: both rows of a pair are valid or neither
  rc->rowbase = ((long long)rc->z * nfy + rc->y) * nfx;
  return true;
}

// Range slot -> [begin, begin + len) in sorted point order.
//   slot = ((dz * 3) + sub) * 2 + part ; sub 0: own interior, 1: own crossing, 2: left crossing ;
//   part 0: y0 in [max(y-w+1, 0), y+1], part 1: the periodic wrap [y-w+1+nfy, nfy-1] (if any).
template <int DIM, int W>
__device__ __forceinline__ void slot_range(const Geom& g, const RowCoord& rc, int slot,
                                           const int32_t* __restrict__ bin_start, int* begin,
                                           int* len) {
  constexpr int NZ = (DIM == 3) ? W : 1;
  *begin = 0;
  *len = 0;
  if (slot >= NZ * 6) return;
  const int part = slot & 1;
  const int sub = (slot >> 1) % 3;
  const int dz = (slot >> 1) / 3;
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int ylo = rc.y - (W - 1);
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
    z0 = rc.z - dz;
    if (z0 < 0) z0 += g.nf[0];
  }
  const int bxx = (sub == 2) ? (rc.bx == 0 ? nbx - 1 : rc.bx - 1) : rc.bx;
  const int cross = sub != 0;
  const long long kb = (((long long)z0 * nbx + bxx) * 2 + cross) * nfy;
  const int s0 = __ldg(bin_start + kb + a);
  *begin = s0;
  *len = __ldg(bin_start + kb + b + 1) - s0;
}

// ------------------------------------------------------------------------------ pre-passes
// Per-point record, computed once per trajectory: everything of a visit that does not depend on
// the visiting tile.  The x weights are stored pair-packed for the parity of the point's offset
// inside its tile (tile lengths are even, so the parity is the same seen from the right
// neighbour): P_q = (w[2q - par], w[2q + 1 - par]).
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
  float wgt[3][8];
  int org[3] = {0, 0, 0};
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // r = 0: x, 1: y, 2: z
#pragma unroll
    for (int i = 0; i < 8; ++i) wgt[r][i] = 0.f;
    const int a = g.dim - 1 - r;
    if (a < 0) {
      wgt[r][0] = 1.f;  // unused axis: single unit tap
      continue;
    }
    const float z = fmaf(2.f, fp[a][s], (float)(W - 1));
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float acc = spoly[g.deg * W + i];
      for (int k = g.deg - 1; k >= 0; --k) acc = fmaf(acc, z, spoly[k * W + i]);
      wgt[r][i] = acc;
    }
    org[r] = op[a][s];
  }
  const int xo = org[0] % CX;  // offset inside the point's own tile
  const bool odd = xo & 1;
  float out[REC];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float we = wgt[0][i];                    // W <= 7: wgt[0][7] = 0
    const float wo = i >= 1 ? wgt[0][i - 1] : 0.f;
    out[i] = odd ? wo : we;
  }
  out[R_WY] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) out[R_WY + 1 + i] = wgt[1][i];
#pragma unroll
  for (int i = 0; i < 7; ++i) out[R_WZ + i] = wgt[2][i];
  out[R_JY] = __int_as_float((xo >> 1) + W / 2);
  out[R_JY + 1] = __int_as_float(org[1]);
  out[26] = out[27] = 0.f;
  float4* dst = reinterpret_cast<float4*>(rec + s * REC);
#pragma unroll
  for (int q = 0; q < REC / 4; ++q)
    dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

__global__ void __launch_bounds__(256)
k_invert_perm(long long M, const int32_t* __restrict__ perm, int32_t* __restrict__ iperm) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < M) iperm[perm[s]] = (int32_t)s;
}

// visits per tile (0 for ids outside the grid) and their grand total
template <int DIM, int W>
__global__ void __launch_bounds__(256)
k_row_totals(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
             int32_t* __restrict__ tot, unsigned long long* __restrict__ grand) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = 0;
  RowCoord rc;
  if (row < nrows && decode_row<DIM>(g, row, &rc)) {
    constexpr int NZ = (DIM == 3) ? W : 1;
    for (int slot = 0; slot < NZ * 6; ++slot) {
      int b, l;
      slot_range<DIM, W>(g, rc, slot, bin_start, &b, &l);
      total += l;
    }
  }
  // tiles outside the grid get -1: no item at all
  if (row <= nrows)
    tot[row] = (row == nrows || !decode_row<DIM>(g, row, &rc)) ? -1 : (int32_t)min(total, (long long)INT32_MAX);
  long long wsum = total;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, d);
  if ((threadIdx.x & 31) == 0 && wsum > 0) atomicAdd(grand, (unsigned long long)wsum);
}

// per tile: (visits, items) as inputs of the two exclusive scans
__global__ void __launch_bounds__(256)
k_scan_inputs(long long n, const int32_t* __restrict__ tot, uint32_t* __restrict__ nvis,
              int32_t* __restrict__ nitem) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tot[i];
  nvis[i] = t > 0 ? (uint32_t)t : 0u;
  nitem[i] = t < 0 ? 0 : (t == 0 ? 1 : (t + CHUNK - 1) / CHUNK);  // an empty tile still owns one item
}

__global__ void __launch_bounds__(256)
k_fill_items(long long nrows, const int32_t* __restrict__ tot,
             const uint32_t* __restrict__ vis_start, const int32_t* __restrict__ item_start,
             int4* __restrict__ items, int32_t* __restrict__ split_rows,
             int* __restrict__ split_counter) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int t = tot[row];
  if (t < 0) return;
  const int at = item_start[row];
  if (t == 0) {
    items[at] = make_int4((int)row, ITEM_EMPTY, 0, 0);
    return;
  }
  const int n = (t + CHUNK - 1) / CHUNK;
  const int flags = n > 1 ? ITEM_SPLIT : 0;
  for (int c = 0; c < n; ++c)
    items[at + c] = make_int4((int)row, flags, (int)(vis_start[row] + (uint32_t)c * CHUNK),
                              min(CHUNK, t - c * CHUNK));
  if (n > 1) split_rows[atomicAdd(split_counter, 1)] = (int32_t)row;
}

// The visit list of every tile, written once per trajectory: one warp per tile concatenates the
// tile's key ranges (slot order = dz-major, as the prefix sums of k_row_totals assume).
template <int DIM, int W>
__global__ void __launch_bounds__(256)
k_build_visits(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
               const int32_t* __restrict__ tot, const uint32_t* __restrict__ vis_start,
               uint32_t* __restrict__ vis) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows || tot[row] <= 0) return;
  RowCoord rc;
  decode_row<DIM>(g, row, &rc);
  uint32_t* out = vis + vis_start[row];
  int b[2], l[2];
  slot_range<DIM, W>(g, rc, lane, bin_start, &b[0], &l[0]);
  slot_range<DIM, W>(g, rc, lane + 32, bin_start, &b[1], &l[1]);
  int inc0 = l[0], inc1 = l[1];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u0 = __shfl_up_sync(0xffffffffu, inc0, d);
    const int u1 = __shfl_up_sync(0xffffffffu, inc1, d);
    if (lane >= d) {
      inc0 += u0;
      inc1 += u1;
    }
  }
  const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
  const int pre[2] = {inc0 - l[0], tot0 + inc1 - l[1]};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int slot = lane + 32 * h;
    const unsigned tag = ((unsigned)((slot >> 1) / 3) << VIS_SBITS) | ((((slot >> 1) % 3) == 2 ? 1u : 0u) << 30);
    // short ranges: written by the owning lane; long ranges (dense k-space centre): by the whole warp
    const bool is_long = l[h] > 64;
    if (!is_long)
      for (int i = 0; i < l[h]; ++i) out[pre[h] + i] = (unsigned)(b[h] + i) | tag;
    unsigned longs = __ballot_sync(0xffffffffu, is_long);
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const int bb = __shfl_sync(0xffffffffu, b[h], src);
      const int ll = __shfl_sync(0xffffffffu, l[h], src);
      const int pp = __shfl_sync(0xffffffffu, pre[h], src);
      const unsigned tg = __shfl_sync(0xffffffffu, tag, src);
      for (int i = lane; i < ll; i += 32) out[pp + i] = (unsigned)(bb + i) | tg;
    }
  }
}

// Transposes between the caller's k-space batch ksp[t][j] and the row kernels' kt[s][t]
// (s = sorted position of sample j).  A warp moves 32 consecutive samples x 32 coils: coalesced
// 256-byte rows per coil on the ksp side, one full 256-byte line per sample on the kt side.
constexpr int KT_WARPS = 4;
constexpr int KT_STRIDE = 33;  // u64 row stride of the transpose tile

// kt[iperm[j]][t] = ksp[t][j] * density[j]   (t < T; coils t >= T are zero-filled)
__global__ void __launch_bounds__(KT_WARPS * 32)
k_gather_kspace(long long M, int T, const int32_t* __restrict__ iperm,
                const float2* __restrict__ ksp, const float* __restrict__ density,
                float2* __restrict__ kt) {
  __shared__ u64 tile[KT_WARPS][32 * KT_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long j0 = ((long long)blockIdx.x * KT_WARPS + warp) * 32;
  if (j0 >= M) return;
  const long long j = j0 + lane;
  const bool live = j < M;
  const float d = (live && density) ? density[j] : 1.f;
  const int s = live ? iperm[j] : -1;
  u64* tl = tile[warp];
  const u64* src = reinterpret_cast<const u64*>(ksp);
#pragma unroll 8
  for (int t = 0; t < 32; ++t) {
    float2 v = make_float2(0.f, 0.f);
    if (t < T && live) {
      const u64 raw = __ldg(src + (long long)t * M + j);
      v = make_float2(__uint_as_float((unsigned)raw) * d, __uint_as_float((unsigned)(raw >> 32)) * d);
    }
    tl[lane * KT_STRIDE + t] = ((u64)__float_as_uint(v.y) << 32) | (u64)__float_as_uint(v.x);
  }
  __syncwarp();
  u64* dst = reinterpret_cast<u64*>(kt);
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int si = __shfl_sync(0xffffffffu, s, i);
    if (si >= 0) dst[(long long)si * 32 + lane] = tl[i * KT_STRIDE + lane];
  }
}

// ksp[t][j] = scale * kt[iperm[j]][t] (- obs[t][j])
__global__ void __launch_bounds__(KT_WARPS * 32)
k_scatter_kspace(long long M, int T, const int32_t* __restrict__ iperm,
                 const float2* __restrict__ kt, float2* __restrict__ ksp, float scale,
                 const float2* __restrict__ obs) {
  __shared__ u64 tile[KT_WARPS][32 * KT_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long j0 = ((long long)blockIdx.x * KT_WARPS + warp) * 32;
  if (j0 >= M) return;
  const long long j = j0 + lane;
  const bool live = j < M;
  const int s = live ? iperm[j] : -1;
  u64* tl = tile[warp];
  const u64* src = reinterpret_cast<const u64*>(kt);
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int si = __shfl_sync(0xffffffffu, s, i);
    tl[i * KT_STRIDE + lane] = si >= 0 ? __ldg(src + (long long)si * 32 + lane) : 0ull;
  }
  __syncwarp();
  if (!live) return;
#pragma unroll 8
  for (int t = 0; t < T; ++t) {
    const u64 raw = tl[lane * KT_STRIDE + t];
    float2 v = make_float2(__uint_as_float((unsigned)raw) * scale, __uint_as_float((unsigned)(raw >> 32)) * scale);
    const long long oi = (long long)t * M + j;
    if (obs) {
      const float2 y = __ldg(obs + oi);
      v.x -= y.x;
      v.y -= y.y;
    }
    ksp[oi] = v;
  }
}

// rows shared by several work items are accumulated with red.add: zero them first
template <int DIM>
__global__ void __launch_bounds__(128)
k_zero_split_rows(Geom g, int T, long long nsplit, const int32_t* __restrict__ split_rows,
                  float2* __restrict__ fw) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nsplit) return;
  RowCoord rc;
  if (!decode_row<DIM>(g, split_rows[w], &rc)) return;
  const int nfx = g.nf[DIM - 1];
  const int x = rc.bx * CX + (lane & 15);
  if (x >= nfx) return;
  float2* dst = fw + rc.rowbase + (long long)(lane >> 4) * nfx + x;
  for (int t = 0; t < T; ++t) dst[(long long)t * g.nftot] = make_float2(0.f, 0.f);
}

// ------------------------------------------------------------------------------ row kernels
// Per-visit tap kernels: generated inline PTX (tools/gen_taps.py) -- one `brx.idx` on the x offset,
// then w packed `fma.rn.f32x2` (SASS FFMA2) on statically indexed 64-bit (re, im) accumulators.
// A C++ `switch` is lowered by nvcc to a compare/branch tree that cost 17 issue slots per visit.
#include "taps_generated.inc"

template <int W>
__device__ __forceinline__ void taps_spread(u64 (&acc)[NACC], unsigned idx, const u64 (&P)[4],
                                            const u64 (&A)[4]) {
  if (W == 7) taps_spread_w7(acc, idx, P, A);
  else if (W == 6) taps_spread_w6(acc, idx, P, A);
  else if (W == 5) taps_spread_w5(acc, idx, P, A);
  else taps_spread_w4(acc, idx, P, A);
}
template <int W>
__device__ __forceinline__ void taps_interp(u64 (&S)[4], u64 (&acc)[NACC], unsigned idx,
                                            const u64 (&P)[4]) {
  if (W == 7) taps_interp_w7(S, acc, idx, P);
  else if (W == 6) taps_interp_w6(S, acc, idx, P);
  else if (W == 5) taps_interp_w5(S, acc, idx, P);
  else taps_interp_w4(S, acc, idx, P);
}

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
__device__ __forceinline__ float lo32(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

// kt[addr] += p for lanes with pred != 0 (vector reduction, no branch)
__device__ __forceinline__ void red_add_f32x2(float2* addr, u64 p, int pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .f32 lo, hi;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "mov.b64 {lo, hi}, %1;\n"
      "@q red.global.add.v2.f32 [%0], {lo, hi};\n"
      "}\n" ::"l"(addr), "l"(p), "r"(pred)
      : "memory");
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// bulk-copy (TMA) variant of the value staging: one 256-byte copy per visit, completion on an mbarrier
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_LOOP;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void* gsrc, unsigned bytes, unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
      "l"(gsrc), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// one staged visit as the consume loop sees it: 48-byte packet, read with warp-broadcast loads
struct Packet {
  u64 P[4];     // pair-packed x weights
  float s0, s1; // row scales wy[dy] wz[dz], wy[dy + 1] wz[dz]
  unsigned idx; // tap-kernel case: floor(off / 2) + W / 2
  unsigned s;   // sorted point index
};
__device__ __forceinline__ void load_packet(unsigned addr, Packet& p) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];\n" : "=l"(p.P[0]), "=l"(p.P[1]) : "r"(addr));
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+16];\n" : "=l"(p.P[2]), "=l"(p.P[3]) : "r"(addr));
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+32];\n"
               : "=f"(p.s0), "=f"(p.s1), "=r"(p.idx), "=r"(p.s)
               : "r"(addr));
}
__device__ __forceinline__ u64 lds64(unsigned addr) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];\n" : "=l"(v) : "r"(addr));
  return v;
}

// registers of one lane's visit between "issue" (record loads in flight) and "finish" (packet written)
struct Staged {
  float4 p0, p1;  // pair-packed x weights
  float wz;
  int2 jy;        // (xo >> 1) + W / 2, y0
  const float* r;
};

template <int DIM, int W, bool SPREAD, bool BULK>
__global__ void __launch_bounds__(THREADS, 4)
k_rows(Geom g, int T, long long nitems, const int4* __restrict__ items,
       const uint32_t* __restrict__ vis, const float* __restrict__ rec,
       float2* __restrict__ kt, float2* __restrict__ fw, int* __restrict__ counter) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = smem_raw + (size_t)warp * SM_WARP;
  const unsigned vbuf_a = smem_u32(wsm);               // [2][VBLK][32] (re, im)
  const unsigned meta_a = smem_u32(wsm + SM_VBUF);     // [2][MBLK] packets of 48 bytes
  const unsigned mbar_a = smem_u32(wsm + SM_VBUF + SM_META);
  uint4* meta = reinterpret_cast<uint4*>(wsm + SM_VBUF);
  // transpose buffers (alias vbuf/meta): real and imaginary planes [32 coils][34] floats, so that a
  // lane reads / writes its (cell 2j, cell 2j+1) register pairs with one conflict-free 64-bit access
  float* tre = reinterpret_cast<float*>(wsm);
  float* tim = tre + 32 * TS;

  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  u64* fw64 = reinterpret_cast<u64*>(fw);
  unsigned phase = 0;  // bit b: parity the next wait on mbarrier b expects
  if (SPREAD && BULK) {
    if (lane == 0) {
      mbar_init(mbar_a, 1);
      mbar_init(mbar_a + 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncwarp();
  }

  for (;;) {
    long long item0 = 0;
    if (lane == 0) item0 = (long long)atomicAdd(counter, GROUP);
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    if (item0 >= nitems) break;
#pragma unroll 1
    for (int gi = 0; gi < GROUP; ++gi) {
      const long long item = item0 + gi;
      if (item >= nitems) break;
      const int4 it = __ldg(items + item);
      RowCoord rc;
      decode_row<DIM>(g, it.x, &rc);
      // flush / load role of this lane: row (lane >> 4) of the pair, cell (lane & 15)
      const int x = rc.bx * CX + (lane & 15);
      u64* gtile = fw64 + rc.rowbase + (long long)(lane >> 4) * nfx + x;
      if (it.y & ITEM_EMPTY) {
        if (SPREAD && x < nfx) {
          for (int t = 0; t < T; ++t) gtile[(long long)t * g.nftot] = 0ull;
        }
        continue;
      }
      const bool split = it.y & ITEM_SPLIT;
      const uint32_t* v = vis + (uint32_t)it.z;
      const int nvis = it.w;
      const int nsub = (nvis + VBLK - 1) / VBLK;
      // a left neighbour's crossing point lands at x offset (xo - length of the left tile)
      const int lhalf = ((rc.bx == 0) ? (nfx - (nbx - 1) * CX) : CX) >> 1;

      // ---- accumulators: acc[r*16 + c*8 + j] = (cell 2j, cell 2j+1) of row r, c = re / im
      u64 acc[NACC];
      if (SPREAD) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0ull;
      } else {
        // load the tile: coalesced per coil (128 B per row) -> smem -> registers (lane = coil)
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
          u64 q = 0ull;
          if (t < T && x < nfx) q = __ldg(gtile + (long long)t * g.nftot);
          tre[t * TS + lane] = lo32(q);
          tim[t * TS + lane] = hi32(q);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[r * 16 + j] = *reinterpret_cast<const u64*>(tre + lane * TS + r * 16 + 2 * j);
            acc[r * 16 + 8 + j] = *reinterpret_cast<const u64*>(tim + lane * TS + r * 16 + 2 * j);
          }
        __syncwarp();
      }

      // ---- staging: packets in blocks of 32 visits (one lane = one visit), values in blocks of 16
      auto meta_issue = [&](unsigned w, Staged& st) {
        if (w != VIS_NONE) {
          const float* r = rec + (long long)(w & VIS_SMASK) * REC;
          st.r = r;
          st.p0 = __ldg(reinterpret_cast<const float4*>(r));
          st.p1 = __ldg(reinterpret_cast<const float4*>(r) + 1);
          st.jy = __ldg(reinterpret_cast<const int2*>(r + R_JY));
          st.wz = (DIM == 3) ? __ldg(r + R_WZ + ((w >> VIS_SBITS) & 7u)) : 1.f;
        }
      };
      auto meta_finish = [&](int mb, unsigned w, const Staged& st) {
        if (w != VIS_NONE) {
          // row scales: row y takes wy[dy], row y+1 takes wy[dy+1]  (dy = y - y0 in [-1, W-1];
          // the record stores 0, wy[0..6], 0 so that both loads are unconditional)
          int dy = rc.y - st.jy.y;
          if (dy < -1) dy += nfy;
          const float s0 = __ldg(st.r + R_WY + 1 + dy) * st.wz;
          const float s1 = __ldg(st.r + R_WY + 2 + dy) * st.wz;
          uint4* m = meta + ((mb & 1) * MBLK + lane) * 3;
          m[0] = make_uint4(__float_as_uint(st.p0.x), __float_as_uint(st.p0.y), __float_as_uint(st.p0.z),
                            __float_as_uint(st.p0.w));
          m[1] = make_uint4(__float_as_uint(st.p1.x), __float_as_uint(st.p1.y), __float_as_uint(st.p1.z),
                            __float_as_uint(st.p1.w));
          m[2] = make_uint4(__float_as_uint(s0), __float_as_uint(s1), (unsigned)(st.jy.x - (int)(w >> 30) * lhalf), w & VIS_SMASK);
        }
      };
      // coil values of value block j (visits 16 j .. 16 j + 15; their words sit in lanes
      // 16 (j & 1) .. of `w`, the word register of packet block j >> 1)
      auto values_issue = [&](int j, unsigned w) {
        const int buf = j & 1;
        if (BULK) {
          const int n = min(VBLK, nvis - j * VBLK);
          fence_proxy_async();
          if (lane == 0) mbar_expect_tx(mbar_a + 8 * buf, (unsigned)n * 256u);
          __syncwarp();
          if ((lane >> 4) == buf && w != VIS_NONE)
            bulk_g2s(vbuf_a + (unsigned)(buf * VBLK + (lane & 15)) * 256u,
                     kt + (long long)(w & VIS_SMASK) * 32, 256u, mbar_a + 8 * buf);
        } else {
          // 2 points per instruction, 16 bytes per lane
#pragma unroll
          for (int i = 0; i < VBLK / 2; ++i) {
            const int kk = 2 * i + (lane >> 4);
            const unsigned wk = __shfl_sync(0xffffffffu, w, buf * VBLK + kk);
            if (wk != VIS_NONE)
              cp_async16(vbuf_a + (unsigned)((buf * VBLK + kk) * 32 + (lane & 15) * 2) * 8u,
                         kt + (long long)(wk & VIS_SMASK) * 32 + (lane & 15) * 2);
          }
          cp_async_commit();
        }
      };
      auto load_word = [&](int mb) -> unsigned {
        const int i = mb * MBLK + lane;
        return i < nvis ? __ldg(v + i) : VIS_NONE;
      };

      unsigned w_cur = load_word(0);   // words of the packet block being consumed / staged
      unsigned w_nxt = load_word(1);   // prefetched one block ahead
      Staged st;
      meta_issue(w_cur, st);
      meta_finish(0, w_cur, st);
      if (SPREAD) values_issue(0, w_cur);
      unsigned w_val = w_cur;          // words of the packet block the next value block belongs to

#pragma unroll 1
      for (int j = 0; j < nsub; ++j) {
        const bool more = j + 1 < nsub;
        const bool new_block = more && ((j + 1) & 1) == 0;
        if (new_block) {
          w_val = w_nxt;
          w_nxt = load_word(((j + 1) >> 1) + 1);
          meta_issue(w_val, st);
        }
        if (SPREAD) {
          if (more) values_issue(j + 1, w_val);
          if (BULK) {
            mbar_wait(mbar_a + 8 * (j & 1), (phase >> (j & 1)) & 1u);
            phase ^= 1u << (j & 1);
          } else {
            if (more) cp_async_wait<1>();
            else cp_async_wait<0>();
          }
        }
        __syncwarp();
        const int n = min(VBLK, nvis - j * VBLK);
        const unsigned pk_a = meta_a + (unsigned)((((j >> 1) & 1) * MBLK + (j & 1) * VBLK) * 48);
        const unsigned vb_a = vbuf_a + (unsigned)((j & 1) * VBLK * 32 + lane) * 8u;

        auto apply = [&](const Packet& p, u64 val) {
          if (SPREAD) {
            const float vx = lo32(val), vy = hi32(val);
            const float a0x = vx * p.s0, a0y = vy * p.s0, a1x = vx * p.s1, a1y = vy * p.s1;
            const u64 A[4] = {pack2(a0x, a0x), pack2(a0y, a0y), pack2(a1x, a1x), pack2(a1y, a1y)};
            taps_spread<W>(acc, p.idx, p.P, A);
          } else {
            u64 S[4];
            taps_interp<W>(S, acc, p.idx, p.P);
            const float px = p.s0 * (lo32(S[0]) + hi32(S[0])) + p.s1 * (lo32(S[2]) + hi32(S[2]));
            const float py = p.s0 * (lo32(S[1]) + hi32(S[1])) + p.s1 * (lo32(S[3]) + hi32(S[3]));
            // coils t >= T hold an all-zero tile: they add zero to their (unused) kt slot
            red_add_f32x2(kt + (long long)p.s * 32 + lane, pack2(px, py), 1);
          }
        };
        // software-pipelined by hand: the packet (and value) of visit k + 1 is in flight while
        // the taps of visit k execute
        Packet pa, pb;
        u64 va = 0ull, vb = 0ull;
        load_packet(pk_a, pa);
        if (SPREAD) va = lds64(vb_a);
        int k = 0;
#pragma unroll 1
        for (; k + 1 < n; k += 2) {
          load_packet(pk_a + (unsigned)(k + 1) * 48u, pb);
          if (SPREAD) vb = lds64(vb_a + (unsigned)(k + 1) * 256u);
          apply(pa, va);
          const unsigned k2 = (unsigned)(k + 2) & (VBLK - 1);
          load_packet(pk_a + k2 * 48u, pa);
          if (SPREAD) va = lds64(vb_a + k2 * 256u);
          apply(pb, vb);
        }
        if (k < n) apply(pa, va);

        if (new_block) meta_finish((j + 1) >> 1, w_val, st);
        __syncwarp();
      }

      if (SPREAD) {
        // flush: registers (lane = coil) -> smem transpose -> coalesced 128-byte rows per coil
        if (BULK) fence_proxy_async();
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<u64*>(tre + lane * TS + r * 16 + 2 * j) = acc[r * 16 + j];
            *reinterpret_cast<u64*>(tim + lane * TS + r * 16 + 2 * j) = acc[r * 16 + 8 + j];
          }
        __syncwarp();
        if (x < nfx) {
          if (split) {
            for (int t = 0; t < T; ++t)
              red_add_f32x2(reinterpret_cast<float2*>(gtile + (long long)t * g.nftot),
                            pack2(tre[t * TS + lane], tim[t * TS + lane]), 1);
          } else {
#pragma unroll 4
            for (int t = 0; t < T; ++t)
              gtile[(long long)t * g.nftot] = pack2(tre[t * TS + lane], tim[t * TS + lane]);
          }
        }
        __syncwarp();
      }
    }
  }
}

constexpr int EFALLBACK = 1;  // internal: the row kernels cannot serve this trajectory

template <int DIM, int W>
int build_items(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long nrows = num_rows<DIM>(p->g);
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  ts->unsupported = false;
  if (nrows >= (1LL << 31) - 2 || p->M >= (long long)VIS_SMASK) {
    ts->unsupported = true;
    return B200_OK;
  }
  if (ts->nrows != nrows) {
    fr(ts->d_tot);
    fr(ts->d_vis_start);
    fr(ts->d_item_start);
    ts->d_tot = ts->d_item_start = nullptr;
    ts->d_vis_start = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_tot, (size_t)(nrows + 1) * 4));
    CUDA_TRY(cudaMalloc(&ts->d_vis_start, (size_t)(nrows + 1) * 4));
    CUDA_TRY(cudaMalloc(&ts->d_item_start, (size_t)(nrows + 1) * 4));
    ts->nrows = nrows;
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, 64, st));
  k_row_totals<DIM, W><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(
      p->g, nrows, p->d_bin_start, ts->d_tot, reinterpret_cast<unsigned long long*>(ts->d_counters + 2));
  CHECK_LAUNCH();
  unsigned long long grand = 0;
  CUDA_TRY(cudaMemcpyAsync(&grand, ts->d_counters + 2, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (grand >= (1ULL << 31)) {
    ts->unsupported = true;
    return B200_OK;
  }
  k_scan_inputs<<<ceil_div(nrows + 1, 256), 256, 0, st>>>(nrows + 1, ts->d_tot, ts->d_vis_start,
                                                          ts->d_item_start);
  CHECK_LAUNCH();
  size_t need = 0, need2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, ts->d_item_start, ts->d_item_start, (int)(nrows + 1), st);
  cub::DeviceScan::ExclusiveSum(nullptr, need2, ts->d_vis_start, ts->d_vis_start, (int)(nrows + 1), st);
  if (need2 > need) need = need2;
  if (need > ts->scan_tmp_bytes) {
    fr(ts->d_scan_tmp);
    ts->d_scan_tmp = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_scan_tmp, need));
    ts->scan_tmp_bytes = need;
  }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ts->d_scan_tmp, need, ts->d_item_start, ts->d_item_start,
                                         (int)(nrows + 1), st));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ts->d_scan_tmp, need, ts->d_vis_start, ts->d_vis_start,
                                         (int)(nrows + 1), st));
  g_kernel_launches += 4;
  int32_t nitems = 0;
  CUDA_TRY(cudaMemcpyAsync(&nitems, ts->d_item_start + nrows, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  fr(ts->d_items);
  fr(ts->d_split_rows);
  fr(ts->d_vis);
  ts->d_items = nullptr;
  ts->d_split_rows = nullptr;
  ts->d_vis = nullptr;
  CUDA_TRY(cudaMalloc(&ts->d_items, (size_t)(nitems > 0 ? nitems : 1) * sizeof(int4)));
  CUDA_TRY(cudaMalloc(&ts->d_split_rows, (size_t)(nitems / 2 + 1) * 4));
  // + 64 words of slack: the kernel prefetches one packet block of words ahead (guarded, but cheap)
  if (cudaMalloc(&ts->d_vis, (size_t)(grand + 64) * 4) != cudaSuccess) {
    cudaGetLastError();
    ts->d_vis = nullptr;
    ts->unsupported = true;
    return B200_OK;
  }
  k_fill_items<<<ceil_div(nrows, 256), 256, 0, st>>>(nrows, ts->d_tot, ts->d_vis_start,
                                                     ts->d_item_start, ts->d_items,
                                                     ts->d_split_rows, ts->d_counters + 1);
  CHECK_LAUNCH();
  k_build_visits<DIM, W><<<ceil_div(nrows * 32, 256), 256, 0, st>>>(p->g, nrows, p->d_bin_start, ts->d_tot,
                                                                    ts->d_vis_start, ts->d_vis);
  CHECK_LAUNCH();
  int nsplit = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsplit, ts->d_counters + 1, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  ts->nitems = nitems;
  ts->nsplit = nsplit;
  ts->nvis = (long long)grand;
  return B200_OK;
}

template <int DIM, int W>
int prepare(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long M = p->M;
  if (!ts->d_counters) CUDA_TRY(cudaMalloc(&ts->d_counters, 64));
  if (ts->d_rec) cudaFree(ts->d_rec);
  if (ts->d_iperm) cudaFree(ts->d_iperm);
  ts->d_rec = nullptr;
  ts->d_iperm = nullptr;
  CUDA_TRY(cudaMalloc(&ts->d_rec, (size_t)(M > 0 ? M : 1) * REC * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ts->d_iperm, (size_t)(M > 0 ? M : 1) * 4));
  if (M > 0) {
    k_point_records<W><<<ceil_div(M, 256), 256, 0, st>>>(
        p->g, M, p->d_poly, p->d_org_s[0], p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],
        p->d_x1_s[1], p->d_x1_s[2], ts->d_rec);
    CHECK_LAUNCH();
    k_invert_perm<<<ceil_div(M, 256), 256, 0, st>>>(M, p->d_perm, ts->d_iperm);
    CHECK_LAUNCH();
  }
  B200_TRY((build_items<DIM, W>(p, ts, st)));
  ts->M = M;
  ts->valid = true;
  return B200_OK;
}

template <int DIM, int W, bool SPREAD>
int launch_rows(b200_plan* p, RowsState* ts, float2* fw, int T, cudaStream_t st) {
  const bool bulk = SPREAD && p->rows_bulk != 0;
  auto kern = SPREAD ? (bulk ? k_rows<DIM, W, SPREAD, true> : k_rows<DIM, W, SPREAD, false>)
                     : k_rows<DIM, W, false, false>;
  const size_t smem = (size_t)WARPS * SM_WARP;
  static bool attr_done[2] = {false, false};
  static int ctas_per_sm[2] = {1, 1};
  const int v = bulk ? 1 : 0;
  if (!attr_done[v]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[v], kern, THREADS, smem));
    if (ctas_per_sm[v] < 1) ctas_per_sm[v] = 1;
    attr_done[v] = true;
  }
  if (SPREAD && ts->nsplit > 0) {
    k_zero_split_rows<DIM><<<ceil_div(ts->nsplit * 32, 128), 128, 0, st>>>(p->g, T, ts->nsplit,
                                                                         ts->d_split_rows, fw);
    CHECK_LAUNCH();
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, sizeof(int), st));
  const long long want = (ts->nitems + GROUP * WARPS - 1) / (GROUP * WARPS);
  const long long cap = (long long)p->num_sms * ctas_per_sm[v];
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  // slot 4 of the plan's timing events brackets the row kernel alone (bench.py roofline)
  const bool timed = p->timing && p->ev_ok;
  if (timed) cudaEventRecord(p->ev[8], st);
  kern<<<grid, THREADS, smem, st>>>(p->g, T, ts->nitems, ts->d_items, ts->d_vis, ts->d_rec,
                                    ts->d_kt, fw, ts->d_counters);
  if (timed) {
    cudaEventRecord(p->ev[9], st);
    p->ev_used[4] = 1;
  }
  CHECK_LAUNCH();
  return B200_OK;
}

int ensure_kt(RowsState* ts, long long M) {
  const size_t need = (size_t)(M > 0 ? M : 1) * 32 * sizeof(float2);
  if (ts->kt_bytes < need) {
    if (ts->d_kt) cudaFree(ts->d_kt);
    ts->d_kt = nullptr;
    ts->kt_bytes = 0;
    CUDA_TRY(cudaMalloc(&ts->d_kt, need));
    ts->kt_bytes = need;
  }
  return B200_OK;
}

}  // namespace

bool tiled_supported(const b200_plan* p, int T) {
  const Geom& g = p->g;
  if (g.dim < 2 || g.dim > 3) return false;
  if (g.w < 4 || g.w > 7) return false;
  if (T > 32) return false;
  if (p->tiled && ((RowsState*)p->tiled)->valid && ((RowsState*)p->tiled)->M == p->M &&
      ((RowsState*)p->tiled)->unsupported)
    return false;
  const int nfx = g.nf[g.dim - 1];
  const int rem = nfx % CX;
  if (rem != 0 && rem < g.w - 1) return false;  // a footprint may touch at most two tiles
  for (int a = 0; a < g.dim; ++a)
    if (g.nf[a] < 2 * g.w) return false;
  return true;
}

void tiled_free(b200_plan* p) {
  if (!p->tiled) return;
  RowsState* ts = (RowsState*)p->tiled;
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  fr(ts->d_rec);
  fr(ts->d_kt);
  fr(ts->d_iperm);
  fr(ts->d_tot);
  fr(ts->d_vis_start);
  fr(ts->d_item_start);
  fr(ts->d_items);
  fr(ts->d_vis);
  fr(ts->d_split_rows);
  fr(ts->d_counters);
  fr(ts->d_scan_tmp);
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

#define DISPATCH_DWS(FN, S, ...)                               \
  do {                                                         \
    const int d_ = p->g.dim, w_ = p->g.w;                      \
    if (d_ == 3) {                                             \
      if (w_ == 7) return FN<3, 7, S>(__VA_ARGS__);            \
      if (w_ == 6) return FN<3, 6, S>(__VA_ARGS__);            \
      if (w_ == 5) return FN<3, 5, S>(__VA_ARGS__);            \
      return FN<3, 4, S>(__VA_ARGS__);                         \
    } else {                                                   \
      if (w_ == 7) return FN<2, 7, S>(__VA_ARGS__);            \
      if (w_ == 6) return FN<2, 6, S>(__VA_ARGS__);            \
      if (w_ == 5) return FN<2, 5, S>(__VA_ARGS__);            \
      return FN<2, 4, S>(__VA_ARGS__);                         \
    }                                                          \
  } while (0)

static int ensure_state(b200_plan* p, cudaStream_t st) {
  RowsState* ts = state(p);
  if (ts->valid && ts->M == p->M) return B200_OK;
  DISPATCH_DW(prepare, p, ts, st);
}

// Both entry points return 1 (not an error) when the row kernels cannot serve this trajectory
// (more than 2^31 visits or 2^27 points): the caller then uses the point-driven kernels.
int spread_tiled(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                 cudaStream_t st) {
  B200_TRY(ensure_state(p, st));
  RowsState* ts = state(p);
  if (ts->unsupported) return EFALLBACK;
  const long long M = p->M;
  B200_TRY(ensure_kt(ts, M));
  if (M > 0) {
    k_gather_kspace<<<ceil_div(M, 32 * KT_WARPS), KT_WARPS * 32, 0, st>>>(M, T, ts->d_iperm, ksp, density,
                                                                          ts->d_kt);
    CHECK_LAUNCH();
  }
  DISPATCH_DWS(launch_rows, true, p, ts, fw, T, st);
}

int interp_tiled(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                 const float2* obs, cudaStream_t st) {
  B200_TRY(ensure_state(p, st));
  RowsState* ts = state(p);
  if (ts->unsupported) return EFALLBACK;
  const long long M = p->M;
  if (M == 0) return B200_OK;
  B200_TRY(ensure_kt(ts, M));
  CUDA_TRY(cudaMemsetAsync(ts->d_kt, 0, (size_t)M * 32 * sizeof(float2), st));
  int rc = [&]() -> int { DISPATCH_DWS(launch_rows, false, p, ts, const_cast<float2*>(fw), T, st); }();
  if (rc != B200_OK) return rc;
  k_scatter_kspace<<<ceil_div(M, 32 * KT_WARPS), KT_WARPS * 32, 0, st>>>(M, T, ts->d_iperm, ts->d_kt, ksp,
                                                                         scale, obs);
  CHECK_LAUNCH();
  return B200_OK;
}
