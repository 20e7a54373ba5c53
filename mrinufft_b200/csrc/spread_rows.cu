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

constexpr int CX = 16;          // cells per tile row (== pencil-bin width, B200_BIN_X)
constexpr int NACC = 32;        // 64-bit accumulator registers per lane: [row 2][re/im 2][cell pair 8]
constexpr int REC = 24;         // floats per point record: wx[7] xo | wy[7] y0 | wz[7] z0
constexpr int BLK = 16;         // visits staged per block
constexpr int WARPS = 4;        // warps per CTA
constexpr int THREADS = WARPS * 32;
constexpr int CHUNK = 4096;     // visits per work item
constexpr int NSLOT = 64;       // range slots per row (2 per lane)
constexpr int GROUP = 4;        // work items fetched per atomic

constexpr int ITEM_EMPTY = -1;        // chunk field: row without visits
constexpr int ITEM_SPLIT = 1 << 30;   // chunk flag: row is shared by several items

// per-warp shared memory (bytes)
constexpr int SM_VBUF = 2 * BLK * 32 * 8;   // double-buffered coil values of BLK points
constexpr int SM_META = 2 * BLK * 48;       // double-buffered packets {P0..P3, s0, s1, idx, s}
constexpr int SM_SLOT = 2 * NSLOT * 4;      // pre[], begin[]
constexpr int SM_WARP = SM_VBUF + SM_META + SM_SLOT;
constexpr int TS = 34;          // row stride (floats) of the transpose planes
static_assert(SM_VBUF + SM_META >= 2 * 32 * TS * 4, "transpose buffers must fit in vbuf+meta");

struct RowsState {
  float* d_rec = nullptr;        // [M][REC] per sorted point
  float2* d_kt = nullptr;        // [M][32] transposed (sorted point, coil) k-space batch
  size_t kt_bytes = 0;
  int32_t* d_nchunks = nullptr;  // [nrows + 1]
  int32_t* d_item_start = nullptr;  // [nrows + 1]
  int2* d_items = nullptr;       // [nitems] {row, chunk | flags}
  int32_t* d_split_rows = nullptr;
  int* d_counters = nullptr;     // [0] work counter, [1] split-row counter
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  long long nrows = 0, nitems = 0, nsplit = 0;
  long long M = -1;
  bool valid = false;
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
    int o = op[a][s];
    if (r == 0) o = o % CX;  // x: offset inside the point's own tile
    out[r * 8 + 7] = __int_as_float(o);
  }
  float4* dst = reinterpret_cast<float4*>(rec + s * REC);
#pragma unroll
  for (int q = 0; q < REC / 4; ++q)
    dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

// visits per row -> number of work items of the row (0 for ids outside the grid)
template <int DIM, int W>
__global__ void __launch_bounds__(256)
k_row_chunks(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
             int32_t* __restrict__ nchunks) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row > nrows) return;
  if (row == nrows) {
    nchunks[row] = 0;
    return;
  }
  RowCoord rc;
  if (!decode_row<DIM>(g, row, &rc)) {
    nchunks[row] = 0;
    return;
  }
  constexpr int NZ = (DIM == 3) ? W : 1;
  long long total = 0;
  for (int slot = 0; slot < NZ * 6; ++slot) {
    int b, l;
    slot_range<DIM, W>(g, rc, slot, bin_start, &b, &l);
    total += l;
  }
  // rows without visits still get one (empty) item, flagged -1: the spreader writes their zeros
  nchunks[row] = total == 0 ? -1 : (int32_t)((total + CHUNK - 1) / CHUNK);
}

__global__ void __launch_bounds__(256)
k_abs_chunks(long long n, const int32_t* __restrict__ in, int32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] < 0 ? 1 : in[i];
}

__global__ void __launch_bounds__(256)
k_fill_items(long long nrows, const int32_t* __restrict__ nchunks,
             const int32_t* __restrict__ item_start, int2* __restrict__ items,
             int32_t* __restrict__ split_rows, int* __restrict__ split_counter) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int n = nchunks[row];
  const int at = item_start[row];
  if (n < 0) {
    items[at] = make_int2((int)row, ITEM_EMPTY);
  } else if (n == 1) {
    items[at] = make_int2((int)row, 0);
  } else if (n > 1) {
    for (int c = 0; c < n; ++c) items[at + c] = make_int2((int)row, c | ITEM_SPLIT);
    split_rows[atomicAdd(split_counter, 1)] = (int32_t)row;
  }
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
typedef unsigned long long u64;
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// registers of one lane's visit between "issue" (loads in flight) and "finish" (meta written)
struct Staged {
  float4 a, b;   // wx[0..6], xo
  float4 c, d;   // wy[0..6], y0
  float wz;
  int s;         // sorted point index, -1 if this lane has no visit in the block
  int left_len;  // 0, or the length of the left tile for left-crossing visits
};

__device__ __forceinline__ float lo32(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int DIM, int W, bool SPREAD>
__global__ void __launch_bounds__(THREADS, 4)
k_rows(Geom g, int T, long long nitems, const int2* __restrict__ items,
       const int32_t* __restrict__ bin_start, const float* __restrict__ rec,
       float2* __restrict__ kt, float2* __restrict__ fw, int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = smem_raw + (size_t)warp * SM_WARP;
  u64* vbuf = reinterpret_cast<u64*>(wsm);                              // [2][BLK][32] (re, im)
  uint4* meta = reinterpret_cast<uint4*>(wsm + SM_VBUF);                // [2][BLK][3]
  int* s_pre = reinterpret_cast<int*>(wsm + SM_VBUF + SM_META);         // [NSLOT]
  int* s_beg = s_pre + NSLOT;                                           // [NSLOT]
  // transpose buffers (alias vbuf/meta): real and imaginary planes [32 coils][34] floats, so that a
  // lane reads / writes its (cell 2j, cell 2j+1) register pairs with one conflict-free 64-bit access
  float* tre = reinterpret_cast<float*>(wsm);
  float* tim = tre + 32 * TS;
  constexpr int JB0 = W / 2;  // idx = floor(off / 2) + JB0

  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  u64* fw64 = reinterpret_cast<u64*>(fw);

  for (;;) {
    long long item0 = 0;
    if (lane == 0) item0 = (long long)atomicAdd(counter, GROUP);
    item0 = __shfl_sync(0xffffffffu, item0, 0);
    if (item0 >= nitems) break;
#pragma unroll 1
    for (int gi = 0; gi < GROUP; ++gi) {
      const long long item = item0 + gi;
      if (item >= nitems) break;
      const int2 it = __ldg(items + item);
      RowCoord rc;
      if (!decode_row<DIM>(g, it.x, &rc)) continue;
      // flush / load role of this lane: row (lane >> 4) of the pair, cell (lane & 15)
      const int x = rc.bx * CX + (lane & 15);
      u64* gtile = fw64 + rc.rowbase + (long long)(lane >> 4) * nfx + x;
      const bool split = (it.y != ITEM_EMPTY) && (it.y & ITEM_SPLIT);

      if (it.y == ITEM_EMPTY) {
        if (SPREAD && x < nfx) {
          for (int t = 0; t < T; ++t) gtile[(long long)t * g.nftot] = 0ull;
        }
        continue;
      }
      const int chunk = it.y & (ITEM_SPLIT - 1);

      // ---- ranges of this tile: 2 slots per lane, exclusive prefix sum over the 64 slots
      int b0, l0, b1, l1;
      slot_range<DIM, W>(g, rc, lane, bin_start, &b0, &l0);
      slot_range<DIM, W>(g, rc, lane + 32, bin_start, &b1, &l1);
      int inc0 = l0, inc1 = l1;
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
      const int total = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
      __syncwarp();
      s_pre[lane] = inc0 - l0;
      s_pre[lane + 32] = tot0 + inc1 - l1;
      s_beg[lane] = b0;
      s_beg[lane + 32] = b1;
      __syncwarp();
      const int v_lo = chunk * CHUNK;
      const int v_hi = min(total, v_lo + CHUNK);
      const int nblk = (v_hi - v_lo + BLK - 1) / BLK;
      const int left_len = (rc.bx == 0) ? (nfx - (nbx - 1) * CX) : CX;

      // ---- accumulators: acc[r*16 + c*8 + j] = (cell 2j, cell 2j+1) of row r, c = re / im
      u64 acc[NACC];
      if (SPREAD) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0ull;
      } else {
        // load the tile: coalesced per coil (128 B per row) -> smem -> registers (lane = coil)
#pragma unroll 8
        for (int t = 0; t < 32; ++t) {
          u64 v = 0ull;
          if (t < T && x < nfx) v = __ldg(gtile + (long long)t * g.nftot);
          tre[t * TS + lane] = lo32(v);
          tim[t * TS + lane] = hi32(v);
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

      // ---- staging helpers
      auto stage_issue = [&](int blk, int buf, Staged& st) {
        const int v = v_lo + blk * BLK + lane;
        st.s = -1;
        st.left_len = 0;
        int dz = 0;
        if (lane < BLK && v < v_hi) {
          int pos = 0;
#pragma unroll
          for (int step = NSLOT / 2; step >= 1; step >>= 1)
            if (s_pre[pos + step] <= v) pos += step;
          st.s = s_beg[pos] + (v - s_pre[pos]);
          const int sub = (pos >> 1) % 3;
          dz = (pos >> 1) / 3;
          if (sub == 2) st.left_len = left_len;
          const float4* r = reinterpret_cast<const float4*>(rec + (long long)st.s * REC);
          st.a = __ldg(r);
          st.b = __ldg(r + 1);
          st.c = __ldg(r + 2);
          st.d = __ldg(r + 3);
          st.wz = (DIM == 3) ? __ldg(rec + (long long)st.s * REC + 16 + dz) : 1.f;
        }
        if (SPREAD) {
          // coil values of the block's points: 2 points per instruction, 16 bytes per lane
          u64* vb = vbuf + buf * (BLK * 32);
#pragma unroll
          for (int i = 0; i < BLK / 2; ++i) {
            const int kk = 2 * i + (lane >> 4);
            const int sk = __shfl_sync(0xffffffffu, st.s, kk);
            if (sk >= 0)
              cp_async16(vb + kk * 32 + (lane & 15) * 2, kt + (long long)sk * 32 + (lane & 15) * 2);
          }
          cp_async_commit();
        }
      };
      auto stage_finish = [&](int buf, const Staged& st) {
        if (st.s >= 0) {
          // row scales: row y takes wy[dy], row y+1 takes wy[dy+1]  (dy = y - y0 in [-1, W-1])
          int dy = rc.y - __float_as_int(st.d.w);
          if (dy < -1) dy += nfy;
          const float wy[8] = {st.c.x, st.c.y, st.c.z, st.c.w, st.d.x, st.d.y, st.d.z, 0.f};
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < W; ++i) {
            s0 = dy == i ? wy[i] : s0;
            s1 = dy + 1 == i ? wy[i] : s1;
          }
          s0 *= st.wz;
          s1 *= st.wz;
          // pair-packed x weights: P_q = (w[2q - par], w[2q + 1 - par])
          const int off = __float_as_int(st.b.w) - st.left_len;
          const int jb = off >> 1;
          const bool odd = off & 1;
          const float w[9] = {0.f, st.a.x, st.a.y, st.a.z, st.a.w, st.b.x, st.b.y, st.b.z, 0.f};
          float pw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // even: pw[i] = wx[i] ; odd: pw[i] = wx[i - 1]   (wx[i] = w[i + 1], zero outside [0, W))
            const float we = (i < W) ? w[i + 1] : 0.f;
            const float wo = (i >= 1 && i - 1 < W) ? w[i] : 0.f;
            pw[i] = odd ? wo : we;
          }
          uint4* m = meta + (buf * BLK + lane) * 3;
          m[0] = make_uint4(__float_as_uint(pw[0]), __float_as_uint(pw[1]), __float_as_uint(pw[2]),
                            __float_as_uint(pw[3]));
          m[1] = make_uint4(__float_as_uint(pw[4]), __float_as_uint(pw[5]), __float_as_uint(pw[6]),
                            __float_as_uint(pw[7]));
          m[2] = make_uint4(__float_as_uint(s0), __float_as_uint(s1), (unsigned)(jb + JB0), (unsigned)st.s);
        }
      };

      Staged st;
      if (nblk > 0) {
        stage_issue(0, 0, st);
        stage_finish(0, st);
      }
      for (int blk = 0; blk < nblk; ++blk) {
        const int cur = blk & 1;
        const bool more = blk + 1 < nblk;
        if (more) stage_issue(blk + 1, cur ^ 1, st);
        if (SPREAD) {
          if (more) cp_async_wait<1>();
          else cp_async_wait<0>();
        }
        __syncwarp();
        const int n = min(BLK, v_hi - v_lo - blk * BLK);
        const uint4* m = meta + cur * BLK * 3;
        const u64* vb = vbuf + cur * (BLK * 32) + lane;
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
          const uint4 m0 = m[3 * k], m1 = m[3 * k + 1], m2 = m[3 * k + 2];
          const u64 P[4] = {((u64)m0.y << 32) | m0.x, ((u64)m0.w << 32) | m0.z,
                            ((u64)m1.y << 32) | m1.x, ((u64)m1.w << 32) | m1.z};
          const float s0 = __uint_as_float(m2.x), s1 = __uint_as_float(m2.y);
          if (SPREAD) {
            const u64 v = vb[k * 32];
            const float vx = lo32(v), vy = hi32(v);
            const float a0x = vx * s0, a0y = vy * s0, a1x = vx * s1, a1y = vy * s1;
            const u64 A[4] = {pack2(a0x, a0x), pack2(a0y, a0y), pack2(a1x, a1x), pack2(a1y, a1y)};
            taps_spread<W>(acc, m2.z, P, A);
          } else {
            u64 S[4];
            taps_interp<W>(S, acc, m2.z, P);
            const float px = s0 * (lo32(S[0]) + hi32(S[0])) + s1 * (lo32(S[2]) + hi32(S[2]));
            const float py = s0 * (lo32(S[1]) + hi32(S[1])) + s1 * (lo32(S[3]) + hi32(S[3]));
            red_add_f32x2(kt + (long long)m2.w * 32 + lane, pack2(px, py), lane < T);
          }
        }
        if (more) stage_finish(cur ^ 1, st);
        __syncwarp();
      }

      if (SPREAD) {
        // flush: registers (lane = coil) -> smem transpose -> coalesced 128-byte rows per coil
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

template <int DIM, int W>
int build_items(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long nrows = num_rows<DIM>(p->g);
  if (nrows >= (1LL << 31) - 2) {
    b200_set_error("too many grid rows (%lld) for the row kernels", nrows);
    return B200_EINVAL;
  }
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  if (ts->nrows != nrows) {
    fr(ts->d_nchunks);
    fr(ts->d_item_start);
    ts->d_nchunks = ts->d_item_start = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_nchunks, (size_t)(nrows + 1) * 4));
    CUDA_TRY(cudaMalloc(&ts->d_item_start, (size_t)(nrows + 1) * 4));
    ts->nrows = nrows;
  }
  k_row_chunks<DIM, W><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(p->g, nrows, p->d_bin_start,
                                                                 ts->d_nchunks);
  CHECK_LAUNCH();
  // item_start = exclusive scan of |nchunks| (an empty row still owns one item)
  k_abs_chunks<<<ceil_div(nrows + 1, 256), 256, 0, st>>>(nrows + 1, ts->d_nchunks, ts->d_item_start);
  CHECK_LAUNCH();
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, ts->d_item_start, ts->d_item_start, (int)(nrows + 1), st);
  if (need > ts->scan_tmp_bytes) {
    fr(ts->d_scan_tmp);
    ts->d_scan_tmp = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_scan_tmp, need));
    ts->scan_tmp_bytes = need;
  }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ts->d_scan_tmp, need, ts->d_item_start, ts->d_item_start,
                                         (int)(nrows + 1), st));
  g_kernel_launches += 2;
  int32_t nitems = 0;
  CUDA_TRY(cudaMemcpyAsync(&nitems, ts->d_item_start + nrows, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  fr(ts->d_items);
  fr(ts->d_split_rows);
  ts->d_items = nullptr;
  ts->d_split_rows = nullptr;
  CUDA_TRY(cudaMalloc(&ts->d_items, (size_t)(nitems > 0 ? nitems : 1) * sizeof(int2)));
  CUDA_TRY(cudaMalloc(&ts->d_split_rows, (size_t)(nitems / 2 + 1) * 4));
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, 64, st));
  k_fill_items<<<ceil_div(nrows, 256), 256, 0, st>>>(nrows, ts->d_nchunks, ts->d_item_start,
                                                     ts->d_items, ts->d_split_rows,
                                                     ts->d_counters + 1);
  CHECK_LAUNCH();
  int nsplit = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsplit, ts->d_counters + 1, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  ts->nitems = nitems;
  ts->nsplit = nsplit;
  return B200_OK;
}

template <int DIM, int W>
int prepare(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long M = p->M;
  if (!ts->d_counters) CUDA_TRY(cudaMalloc(&ts->d_counters, 64));
  if (ts->d_rec) cudaFree(ts->d_rec);
  ts->d_rec = nullptr;
  CUDA_TRY(cudaMalloc(&ts->d_rec, (size_t)(M > 0 ? M : 1) * REC * sizeof(float)));
  if (M > 0) {
    k_point_records<W><<<ceil_div(M, 256), 256, 0, st>>>(
        p->g, M, p->d_poly, p->d_org_s[0], p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],
        p->d_x1_s[1], p->d_x1_s[2], ts->d_rec);
    CHECK_LAUNCH();
  }
  B200_TRY((build_items<DIM, W>(p, ts, st)));
  ts->M = M;
  ts->valid = true;
  return B200_OK;
}

template <int DIM, int W, bool SPREAD>
int launch_rows(b200_plan* p, RowsState* ts, float2* fw, int T, cudaStream_t st) {
  auto kern = k_rows<DIM, W, SPREAD>;
  const size_t smem = (size_t)WARPS * SM_WARP;
  static bool attr_done = false;
  static int ctas_per_sm = 1;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, THREADS, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    attr_done = true;
  }
  if (SPREAD && ts->nsplit > 0) {
    k_zero_split_rows<DIM><<<ceil_div(ts->nsplit * 32, 128), 128, 0, st>>>(p->g, T, ts->nsplit,
                                                                         ts->d_split_rows, fw);
    CHECK_LAUNCH();
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, sizeof(int), st));
  const long long want = (ts->nitems + GROUP * WARPS - 1) / (GROUP * WARPS);
  const long long cap = (long long)p->num_sms * ctas_per_sm;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  // slot 4 of the plan's timing events brackets the row kernel alone (bench.py roofline)
  const bool timed = p->timing && p->ev_ok;
  if (timed) cudaEventRecord(p->ev[8], st);
  kern<<<grid, THREADS, smem, st>>>(p->g, T, ts->nitems, ts->d_items, p->d_bin_start, ts->d_rec,
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
  fr(ts->d_nchunks);
  fr(ts->d_item_start);
  fr(ts->d_items);
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

int spread_tiled(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                 cudaStream_t st) {
  B200_TRY(ensure_state(p, st));
  RowsState* ts = state(p);
  const long long M = p->M;
  B200_TRY(ensure_kt(ts, M));
  if (M > 0) {
    k_gather_kspace<<<ceil_div(M * 32, 256), 256, 0, st>>>(M, T, p->d_perm, ksp, density, ts->d_kt);
    CHECK_LAUNCH();
  }
  DISPATCH_DWS(launch_rows, true, p, ts, fw, T, st);
}

int interp_tiled(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                 const float2* obs, cudaStream_t st) {
  B200_TRY(ensure_state(p, st));
  RowsState* ts = state(p);
  const long long M = p->M;
  if (M == 0) return B200_OK;
  B200_TRY(ensure_kt(ts, M));
  CUDA_TRY(cudaMemsetAsync(ts->d_kt, 0, (size_t)M * 32 * sizeof(float2), st));
  int rc = [&]() -> int { DISPATCH_DWS(launch_rows, false, p, ts, const_cast<float2*>(fw), T, st); }();
  if (rc != B200_OK) return rc;
  k_scatter_kspace<<<ceil_div(M * 32, 256), 256, 0, st>>>(M, T, p->d_perm, ts->d_kt, ksp, scale, obs);
  CHECK_LAUNCH();
  return B200_OK;
}
