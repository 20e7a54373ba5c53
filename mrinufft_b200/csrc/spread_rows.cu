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
#include <type_traits>

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
constexpr int VBLK = 16;        // visits per value block (cp.async granularity)
constexpr int MBLK = 32;        // visits per packet block (one lane stages one visit)
constexpr int WARPS = 4;        // warps per CTA
constexpr int THREADS = WARPS * 32;
constexpr int LCH = 2048;       // stream entries per chunk (= one unit of dynamically scheduled work)

// The visit stream: 16-byte entries {s0, s1, idx, s}, tile after tile in tile-id order.
//   visit entry  : row scales s0, s1, tap-kernel case idx (small), sorted point index s
//   header entry : idx = IDX_HDR, s = tile id -- "the following visits belong to this tile"
//   padding      : all ones (behind the end of the stream; reads as a header)
constexpr unsigned IDX_HDR = 0xffffffffu;
constexpr unsigned IDX_NONE = 0xfffffffeu;  // in registers only: lane beyond the end of the chunk

// per-warp shared memory (bytes)
constexpr int SM_VBUF = 2 * VBLK * 32 * 8;  // double-buffered coil values of VBLK points (spreader only)
constexpr int SM_META = 2 * MBLK * 48;      // double-buffered packets {P0..P3, s0, s1, idx, s}
constexpr int TBS = 10;                     // float stride of the transpose planes [32 coils][8 cells]
constexpr int SM_TBUF = 2 * 32 * TBS * 4;   // real + imaginary plane of half a grid row (8 cells), 32 coils
constexpr int TFS = 18;                     // spreader's flush: planes [16 coils][16 cells], float stride 18
static_assert(2 * 16 * TFS * 4 <= SM_TBUF, "flush planes must fit in the transpose buffer");
__host__ __device__ constexpr int smem_per_warp(bool spread) { return (spread ? SM_VBUF : 0) + SM_META + SM_TBUF; }

struct RowsState {
  float* d_rec = nullptr;        // [M][REC] per sorted point
  float2* d_kt = nullptr;        // [M][32] transposed (sorted point, coil) k-space batch
  size_t kt_bytes = 0;
  int32_t* d_iperm = nullptr;    // [M] point index -> sorted position
  int32_t* d_tot = nullptr;      // [nrows + 1] visits per tile (-1: tile id outside the grid)
  uint32_t* d_start = nullptr;   // [nrows + 1] stream position of a tile's header word
  uint4* d_ent = nullptr;        // [S + slack] the visit stream
  float* d_ptab = nullptr;       // [M][8] pair-packed x weights per sorted point
  uint32_t* d_empty = nullptr;   // empty-tile bit strings (k_mark_empty)
  size_t empty_cap = 0;
  int32_t* d_chunk_row = nullptr;  // [nchunks] tile owning the first word of a chunk
  int32_t* d_split_rows = nullptr; // tiles cut by a chunk boundary (accumulated with red.add)
  int* d_counters = nullptr;     // [0] work counter, [1] split-row counter, [2..3] total visits (u64)
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  size_t ent_cap = 0, chunk_cap = 0, pts_cap = 0;  // capacities (entries, chunks, points): grow-only
  long long nrows = 0, nsplit = 0, nvis = 0;
  unsigned S = 0;                // stream length in entries
  int nchunks = 0;
  int lch = LCH;  // entries per chunk of this trajectory's stream (smaller for short streams, see build_stream)
  long long M = -1;
  bool valid = false;
  bool unsupported = false;      // too many visits / points for the 32-bit stream words
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
                float* __restrict__ rec, float* __restrict__ ptab) {
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
  float4* pd = reinterpret_cast<float4*>(ptab + s * 8);
  pd[0] = make_float4(out[0], out[1], out[2], out[3]);
  pd[1] = make_float4(out[4], out[5], out[6], out[7]);
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

// stream words per tile: header + visits (0 for tile ids outside the grid)
__global__ void __launch_bounds__(256)
k_scan_inputs(long long n, const int32_t* __restrict__ tot, uint32_t* __restrict__ words) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tot[i];
  words[i] = t < 0 ? 0u : (uint32_t)t + 1u;
}

// One flag bit per tile, 1 = no visitors, laid out as a bit string per grid column of tiles along the
// slowest axis: 3-D word [(y / 2) nbx + bx][z / 32], bit z % 32; 2-D word [bx][(y / 2) / 32], bit
// (y / 2) % 32.  Lets the spreader skip the zero-fill of such tiles when the next consumer is the fused
// FFT, whose first pass (along that axis) then substitutes zeros instead of reading them (half the grid
// for a radial trajectory); a CTA of that pass needs one contiguous bit string.
template <int DIM>
__global__ void __launch_bounds__(256)
k_mark_empty(Geom g, long long nrows, const int32_t* __restrict__ tot, uint32_t* __restrict__ bits,
             int words_per_col) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  RowCoord rc;
  if (!decode_row<DIM>(g, row, &rc) || tot[row] != 0) return;
  const int nyh = g.nf[DIM - 2] / 2, nbx = num_xtiles<DIM>(g);
  const long long col = DIM == 3 ? (long long)(rc.y >> 1) * nbx + rc.bx : rc.bx;
  const int pos = DIM == 3 ? rc.z : (rc.y >> 1);
  (void)nyh;
  atomicOr(bits + col * words_per_col + (pos >> 5), 1u << (pos & 31));
}

// The visit stream, written once per trajectory: 16-byte entries {s0, s1, idx, s}, tile after tile.
// One warp per tile writes the tile's header entry {0, 0, IDX_HDR, tile id} and, behind it, one entry
// per visit with everything the row kernels need: the two row scales wy[dy] wz[dz], wy[dy+1] wz[dz],
// the tap-kernel case idx and the sorted point index.  It also records which tile owns the first
// entry of every chunk and which tiles are cut by a chunk boundary.
template <int DIM, int W>
__global__ void __launch_bounds__(256)
k_build_stream(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
               const int32_t* __restrict__ tot, const uint32_t* __restrict__ start,
               const float* __restrict__ rec, uint4* __restrict__ ent,
               int32_t* __restrict__ chunk_row, int32_t* __restrict__ split_rows,
               int* __restrict__ split_counter, uint32_t lch) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int t = tot[row];
  if (t < 0) return;
  const uint32_t hs = start[row], he = hs + 1u + (uint32_t)t;
  if (lane == 0) {
    ent[hs] = make_uint4(0u, 0u, IDX_HDR, (uint32_t)row);
    for (uint32_t c = (hs + lch - 1) / lch; c * lch < he; ++c) chunk_row[c] = (int32_t)row;
    if (hs / lch != (he - 1) / lch) split_rows[atomicAdd(split_counter, 1)] = (int32_t)row;
  }
  if (t == 0) return;
  RowCoord rc;
  decode_row<DIM>(g, row, &rc);
  const int nfx = g.nf[DIM - 1], nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  // a left neighbour's crossing point lands at x offset (xo - length of the left tile)
  const int lhalf = ((rc.bx == 0) ? (nfx - (nbx - 1) * CX) : CX) >> 1;
  uint4* out = ent + hs + 1;
  auto entry = [&](int s, int dz, int left) -> uint4 {
    const float* r = rec + (long long)s * REC;
    const int2 jy = *reinterpret_cast<const int2*>(r + R_JY);
    // row y takes wy[dy], row y+1 takes wy[dy+1]  (dy = y - y0 in [-1, W-1]; the record stores
    // 0, wy[0..6], 0 so that both loads are unconditional)
    int dy = rc.y - jy.y;
    if (dy < -1) dy += nfy;
    const float wz = (DIM == 3) ? r[R_WZ + dz] : 1.f;
    return make_uint4(__float_as_uint(r[R_WY + 1 + dy] * wz), __float_as_uint(r[R_WY + 2 + dy] * wz),
                      (unsigned)(jy.x - left * lhalf), (unsigned)s);
  };
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
  // Tiles made of short ranges only (all but the dense k-space centre): the visits are grouped by
  // their tap-kernel case `idx` with a counting sort, so that the consume loop runs through
  // straight-line code for whole runs of visits (tools/gen_taps.py).  Deterministic: inside a case
  // the order is (lane, slot, position).
  constexpr int NC = 8 + W / 2;  // number of cases
  __shared__ int s_cnt[8][NC][32];
  const int wib = threadIdx.x >> 5;
  const bool any_long = __any_sync(0xffffffffu, l[0] > 64 || l[1] > 64);
  if (!any_long) {
#pragma unroll
    for (int c = 0; c < NC; ++c) s_cnt[wib][c][lane] = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int slot = lane + 32 * h;
      const int dz = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
      for (int i = 0; i < l[h]; ++i) s_cnt[wib][entry(b[h] + i, dz, left).z][lane] += 1;
    }
    // offsets: case-major, lane-minor exclusive prefix
    int base = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int mine = s_cnt[wib][c][lane];
      int inc = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
      }
      s_cnt[wib][c][lane] = base + inc - mine;
      base += __shfl_sync(0xffffffffu, inc, 31);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int slot = lane + 32 * h;
      const int dz = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
      for (int i = 0; i < l[h]; ++i) {
        const uint4 e = entry(b[h] + i, dz, left);
        out[s_cnt[wib][e.z][lane]++] = e;
      }
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int slot = lane + 32 * h;
    const int dz = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
    // short ranges: written by the owning lane; long ranges (dense k-space centre): by the whole warp
    const bool is_long = l[h] > 64;
    if (!is_long)
      for (int i = 0; i < l[h]; ++i) out[pre[h] + i] = entry(b[h] + i, dz, left);
    unsigned longs = __ballot_sync(0xffffffffu, is_long);
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const int bb = __shfl_sync(0xffffffffu, b[h], src);
      const int ll = __shfl_sync(0xffffffffu, l[h], src);
      const int pp = __shfl_sync(0xffffffffu, pre[h], src);
      const int sdz = __shfl_sync(0xffffffffu, dz, src);
      const int sl = __shfl_sync(0xffffffffu, left, src);
      for (int i = lane; i < ll; i += 32) out[pp + i] = entry(bb + i, sdz, sl);
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
__device__ __forceinline__ void rows_loop_spread(u64 (&acc)[NACC], unsigned pk, int n, unsigned vb) {
  if (W == 7) rows_loop_spread_w7(acc, pk, n, vb);
  else if (W == 6) rows_loop_spread_w6(acc, pk, n, vb);
  else if (W == 5) rows_loop_spread_w5(acc, pk, n, vb);
  else rows_loop_spread_w4(acc, pk, n, vb);
}
template <int W>
__device__ __forceinline__ void rows_loop_interp(u64 (&acc)[NACC], unsigned pk, int n, const void* ktl) {
  if (W == 7) rows_loop_interp_w7(acc, pk, n, ktl);
  else if (W == 6) rows_loop_interp_w6(acc, pk, n, ktl);
  else if (W == 5) rows_loop_interp_w5(acc, pk, n, ktl);
  else rows_loop_interp_w4(acc, pk, n, ktl);
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
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// A staged visit is a 48-byte packet {P0..P3 | s0, s1, idx, s} (tools/gen_taps.py): the first 32
// bytes are the point's pair-packed x weights, copied from `ptab` with cp.async, the last 16 bytes
// are the visit's stream entry.  idx = IDX_HDR marks a tile header, with s = tile id.
__device__ __forceinline__ unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

// The row kernel.  A warp consumes one chunk of the visit stream at a time; the tile whose
// accumulators it holds changes whenever a header entry comes by.  Staging is pure data movement:
// entries are read one packet block ahead (coalesced), the x weights and the coil values of their
// points arrive through cp.async, and their cache lines are pulled into L2 another block earlier.
// FIXED: chunks of LCH entries, a compile-time constant (the long streams the kernel is tuned on: with the
// chunk length in a register the spreader spills three more words); otherwise 2^lch_log2 entries (short streams)
template <int DIM, int W, bool SPREAD, bool FIXED>
__global__ void __launch_bounds__(THREADS, 4)
k_rows(Geom g, int T, int nchunks, unsigned S, long long M, const uint4* __restrict__ ent,
       const int32_t* __restrict__ chunk_row, const float* __restrict__ ptab,
       float2* __restrict__ kt, float2* __restrict__ fw, int* __restrict__ counter, int dbg, int skip_empty,
       int lch_log2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int SMW = smem_per_warp(SPREAD);
  constexpr unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = lane >> 4, cl = lane & 15;
  unsigned char* wsm = smem_raw + (size_t)warp * SMW;
  // (made opaque: nvcc otherwise re-derives these addresses from %tid at every use)
  asm volatile("" : "+l"(wsm));
  const unsigned vbuf_a = smem_u32(wsm);                         // [2][VBLK][32] (re, im); spreader only
  unsigned char* metap = wsm + (SPREAD ? SM_VBUF : 0);
  const unsigned meta_a = smem_u32(metap);                       // [2][MBLK] packets of 48 bytes
  // transpose planes [32 coils][TBS]: lane = coil on the register side; on the grid side a warp
  // instruction moves 8 cells (64 contiguous bytes) of 4 coils
  float* tre = reinterpret_cast<float*>(metap + SM_META);
  float* tim = tre + 32 * TBS;  // (the flush's [16][TFS] planes start at the same offsets)
  const int g4 = lane >> 3, c8 = lane & 7;  // grid-side role: coil 4 i + g4, cell 8 h + c8
  const char* ktl = reinterpret_cast<const char*>(kt) + (SPREAD ? cl * 16 : lane * 8);
  asm volatile("" : "+l"(ktl));

  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  u64* fw64 = reinterpret_cast<u64*>(fw);

  for (;;) {
    int c = 0;
    if (lane == 0) c = atomicAdd(counter, 1);
    c = __shfl_sync(FULL, c, 0);
    if (c >= nchunks) break;
    const int lg = FIXED ? 11 : lch_log2;
    static_assert(LCH == 2048, "FIXED chunks are 2^11 entries");
    const unsigned base = (unsigned)c << lg;
    const uint4* v = ent + base;
    const int nw = (int)min(1u << lg, S - base);
    // consume blocks: 16 entries for the spreader (the granularity of its value copies), a whole packet
    // block of 32 for the interpolator (half the per-block bookkeeping)
    constexpr int SB = SPREAD ? VBLK : MBLK, SPB = MBLK / SB;
    const int nsub = (nw + SB - 1) / SB;
    // is the entry behind the chunk a header (or the end of the stream)?  Then the last tile ends here.
    const bool tail_whole = __ldg(reinterpret_cast<const unsigned*>(v + nw) + 2) == IDX_HDR;

    // ---- the tile in the accumulators
    u64* gbase = nullptr;  // grid-side role of this lane: cell c8 of coil g4, row 0 of the tile
    u64* fbase = nullptr;  // ... in the spreader's flush: cell cl of coil hl
    int xlim = 0;          // cells of this tile inside the grid (16, less for a short last tile)
    auto tile_setup = [&](int row) {
      // 32-bit version of decode_row (tile ids are below 2^30 here)
      const int nz = DIM == 3 ? g.nf[0] : 1;
      const int ps = row & 3;
      int r = row >> 2;
      const int bx = r % nbx;
      r /= nbx;
      const int pg = r % (YB / 8);
      r /= (YB / 8);
      const int z = r % nz;
      const int yb = r / nz;
      const int y = yb * YB + pg * 8 + ps * 2;
      xlim = nfx - bx * CX;
      gbase = fw64 + ((long long)z * nfy + y) * nfx + bx * CX + c8 + (long long)g4 * g.nftot;
      if (SPREAD) fbase = fw64 + ((long long)z * nfy + y) * nfx + bx * CX + cl + (long long)hl * g.nftot;
    };
    tile_setup(__ldg(chunk_row + c));
    bool started = false;  // the tile's header came by in this chunk
    bool loaded = false;   // interpolator: the tile is in the registers
    bool dirty = false;    // some visit was applied to the tile

    // accumulators: acc[r*16 + c*8 + j] = (cell 2j, cell 2j+1) of row r, c = re / im
    u64 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0ull;

    // registers (lane = coil) -> grid rows: 16 coils of one row at a time go through the transpose
    // planes; a store instruction writes one full 128-byte line of 2 coils.
    // MODE 0: plain stores (the tile is complete), 1: red.add (tile shared with other chunks)
    auto flush = [&](auto mode) {
      constexpr int MODE = decltype(mode)::value;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (hl == half) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              *reinterpret_cast<u64*>(tre + cl * TFS + 2 * j) = acc[r * 16 + j];
              *reinterpret_cast<u64*>(tim + cl * TFS + 2 * j) = acc[r * 16 + 8 + j];
            }
          }
          __syncwarp();
          if (cl < xlim) {
            u64* dst = fbase + (long long)r * nfx + (long long)(half * 16) * g.nftot;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int tl = 2 * i + hl;  // coil inside this half
              if (T == 32 || half * 16 + tl < T) {
                const u64 val = pack2(tre[tl * TFS + cl], tim[tl * TFS + cl]);
                u64* a = dst + (long long)(2 * i) * g.nftot;
                // streaming stores: the grid is written once and not read again by this kernel --
                // keep L2 for the point data (coil rows, x weights) that neighbouring tiles re-read
                if (MODE == 0) __stcs(a, val);
                else red_add_f32x2(reinterpret_cast<float2*>(a), val, 1);
              }
            }
          }
          __syncwarp();
        }
    };
    // a tile without visits: plain zero stores
    auto store_zero = [&]() {
      if (cl < xlim && !skip_empty) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
          for (int t = hl; t < T; t += 2) __stcs(fbase + (long long)r * nfx + (long long)(t - hl) * g.nftot, 0ull);
      }
    };
    // grid rows -> registers
    auto load_tile = [&]() {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        u64 q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int h = i >> 3, t = 4 * (i & 7) + g4;
          q[i] = (8 * h + c8 < xlim && t < T)
                     ? __ldg(gbase + (long long)r * nfx + 8 * h + (long long)(t - g4) * g.nftot)
                     : 0ull;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            tre[(4 * i + g4) * TBS + c8] = lo32(q[h * 8 + i]);
            tim[(4 * i + g4) * TBS + c8] = hi32(q[h * 8 + i]);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[r * 16 + 4 * h + j] = *reinterpret_cast<const u64*>(tre + lane * TBS + 2 * j);
            acc[r * 16 + 8 + 4 * h + j] = *reinterpret_cast<const u64*>(tim + lane * TBS + 2 * j);
          }
          __syncwarp();
        }
      }
    };

    // ---- staging: packets in blocks of 32 entries (one lane = one entry), values in blocks of 16.
    // Pure data movement: a lane only ever holds the {idx, s} half of its entry in registers.
    auto load_is = [&](int mb) -> uint2 {  // {idx, s} of this lane's entry of packet block mb
      const int i = mb * MBLK + lane;
      return i < nw ? __ldg(reinterpret_cast<const uint2*>(v + i) + 1) : make_uint2(IDX_NONE, 0u);
    };
    // pull the x-weight and coil-value lines of a block's points into L2 ahead of the cp.async copies
    auto prefetch_points = [&](const uint2& is) {
      if (is.x < IDX_NONE) {
        prefetch_l2(ptab + (long long)is.y * 8);
        if (SPREAD) {
          const char* q = reinterpret_cast<const char*>(kt + (long long)is.y * 32);
          prefetch_l2(q);
          prefetch_l2(q + 128);
        }
      }
    };
    // ... and the stream itself, three packet blocks (1.5 KB) ahead
    auto prefetch_stream = [&](int mb) {
      const int i = mb * MBLK + lane * 8;
      if (lane < 4 && i < nw) prefetch_l2(v + i);
    };
    // packet block mb: entry -> last 16 bytes of the packet, x weights -> first 32 bytes (cp.async);
    // returns the header mask of the block
    auto stage_block = [&](int mb, const uint2& is) -> unsigned {
      const unsigned row_a = meta_a + (unsigned)(((mb & 1) * MBLK + lane) * 48);
      if (is.x != IDX_NONE) cp_async16(row_a + 32u, v + mb * MBLK + lane);
      if (is.x < IDX_NONE) {
        const float* pw = ptab + (long long)is.y * 8;
        cp_async16(row_a, pw);
        cp_async16(row_a + 16u, pw + 4);
      }
      return __ballot_sync(FULL, is.x == IDX_HDR);
    };
    // coil values of value block j (entries 16 j .. 16 j + 15 sit in lanes 16 (j & 1) .. of `is`, the
    // registers of packet block j >> 1): 2 points per instruction, 16 bytes per lane
    // (destination and source-lane bases are formed once per call: inside the predicated copies the
    // compiler re-derived them per point)
    const unsigned vdst0 = vbuf_a + (unsigned)(hl * 32 + cl * 2) * 8u;
    auto values_issue = [&](int j, const uint2& is) {
      const int buf = j & 1;
      const unsigned sv = is.x < IDX_NONE ? is.y : IDX_NONE;
      const unsigned vdst = vdst0 + (unsigned)buf * (VBLK * 256u);
      const int lane0 = buf * VBLK + hl;
#pragma unroll
      for (int i = 0; i < VBLK / 2; ++i) {
        const unsigned sk = __shfl_sync(FULL, sv, lane0 + 2 * i);
        if (sk != IDX_NONE) cp_async16(vdst + (unsigned)i * 512u, ktl + (unsigned long long)sk * 256u);
      }
    };

    uint2 is_val = load_is(0);   // {idx, s} of the packet block being staged / whose values are fetched
    uint2 is_nxt = load_is(1);   // one block ahead (prefetched into L2 half a block before its staging)
    prefetch_stream(2);
    prefetch_stream(3);
    unsigned hb_cur = stage_block(0, is_val);
    unsigned hb_nxt = 0;
    if (SPREAD) values_issue(0, is_val);
    cp_async_commit();

#pragma unroll 1
    for (int j = 0; j < nsub; ++j) {
      const bool more = j + 1 < nsub;
      const bool new_block = more && ((j + 1) % SPB) == 0;
      if (new_block) {
        const int nb = (j + 1) / SPB;
        is_val = is_nxt;
        is_nxt = load_is(nb + 1);
        prefetch_stream(nb + 3);
        hb_nxt = stage_block(nb, is_val);
      } else {
        prefetch_points(is_nxt);
      }
      if (SPREAD && more && !(dbg & 2)) values_issue(j + 1, is_val);
      cp_async_commit();
      if (more) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncwarp();
      const int n = min(SB, nw - j * SB);
      const unsigned pk_a = meta_a + (unsigned)((((j / SPB) & 1) * MBLK + (j % SPB) * SB) * 48);
      const unsigned vb_a = vbuf_a + (unsigned)((j & 1) * VBLK * 32 + lane) * 8u;

      // the sub-block is a sequence of visit runs separated by header entries
      unsigned hm = SPB == 1 ? hb_cur : ((hb_cur >> ((j % SPB) * SB)) & 0xffffu);
      int k0 = 0;
      for (;;) {
        const int k1 = hm ? (__ffs(hm) - 1) : n;
        if (k1 > k0) {
          if (!SPREAD && !loaded) {
            load_tile();
            loaded = true;
          }
          dirty = true;
          if (SPREAD) rows_loop_spread<W>(acc, pk_a + (unsigned)k0 * 48u, k1 - k0, vb_a + (unsigned)k0 * 256u);
          else rows_loop_interp<W>(acc, pk_a + (unsigned)k0 * 48u, k1 - k0, ktl);
        }
        if (k1 >= n) break;
        hm &= hm - 1;
        // a header: the tile in the registers is complete if its own header came by in this chunk
        if (SPREAD && !(dbg & 1)) {
          if (dirty) {
            if (started) flush(std::integral_constant<int, 0>());
            else flush(std::integral_constant<int, 1>());
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = 0ull;
          } else if (started) {
            store_zero();
          }
        }
        tile_setup((int)lds32(pk_a + (unsigned)k1 * 48u + 44u));
        started = true;
        loaded = false;
        dirty = false;
        k0 = k1 + 1;
      }
      if (new_block) hb_cur = hb_nxt;
      __syncwarp();
    }

    if (SPREAD && !(dbg & 1)) {
      if (dirty) {
        if (started && tail_whole) flush(std::integral_constant<int, 0>());
        else flush(std::integral_constant<int, 1>());
      } else if (started && tail_whole) {
        store_zero();
      }
    }
  }
}

constexpr int EFALLBACK = 1;  // internal: the row kernels cannot serve this trajectory

template <int DIM, int W>
int build_stream(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long nrows = num_rows<DIM>(p->g);
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  ts->unsupported = false;
  if (nrows >= (1LL << 30) || p->M >= (1LL << 31) - 2 || (p->rows_dbg & 8)) {  // bit 3: test hook
    ts->unsupported = true;
    return B200_OK;
  }
  if (ts->nrows != nrows) {
    fr(ts->d_tot);
    fr(ts->d_start);
    ts->d_tot = nullptr;
    ts->d_start = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_tot, (size_t)(nrows + 1) * 4));
    CUDA_TRY(cudaMalloc(&ts->d_start, (size_t)(nrows + 1) * 4));
    ts->nrows = nrows;
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, 64, st));
  k_row_totals<DIM, W><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(
      p->g, nrows, p->d_bin_start, ts->d_tot, reinterpret_cast<unsigned long long*>(ts->d_counters + 2));
  CHECK_LAUNCH();
  unsigned long long grand = 0;
  CUDA_TRY(cudaMemcpyAsync(&grand, ts->d_counters + 2, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (grand + (unsigned long long)nrows >= (1ULL << 31)) {
    ts->unsupported = true;
    return B200_OK;
  }
  k_scan_inputs<<<ceil_div(nrows + 1, 256), 256, 0, st>>>(nrows + 1, ts->d_tot, ts->d_start);
  CHECK_LAUNCH();
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, ts->d_start, ts->d_start, (int)(nrows + 1), st);
  if (need > ts->scan_tmp_bytes) {
    fr(ts->d_scan_tmp);
    ts->d_scan_tmp = nullptr;
    CUDA_TRY(cudaMalloc(&ts->d_scan_tmp, need));
    ts->scan_tmp_bytes = need;
  }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ts->d_scan_tmp, need, ts->d_start, ts->d_start, (int)(nrows + 1), st));
  g_kernel_launches += 2;
  uint32_t S = 0;
  CUDA_TRY(cudaMemcpyAsync(&S, ts->d_start + nrows, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // Chunk = unit of dynamically scheduled work.  2048 entries amortise the per-chunk set-up on long streams;
  // a short stream (2-D, few samples) is cut finer so that every resident warp still gets several chunks
  // (cfg-B: 7.7e5 entries = 376 chunks of 2048 for 2368 warp slots).
  int lch = LCH;
  while (lch > 128 && (long long)S / lch < 8LL * p->num_sms * 4 * WARPS) lch >>= 1;
  ts->lch = lch;
  const int nchunks = (int)((S + lch - 1) / lch);
  // + 64 entries of all-ones padding: the kernel looks one entry past the last chunk.
  // Buffers only ever grow: update_samples in a trajectory-learning loop must not pay cudaFree's
  // device synchronisation and a multi-GB cudaMalloc per step.
  if ((size_t)S + 64 > ts->ent_cap) {
    fr(ts->d_ent);
    ts->d_ent = nullptr;
    ts->ent_cap = 0;
    const size_t cap = (size_t)S + 64 + (size_t)S / 16;
    if (cudaMalloc(&ts->d_ent, cap * 16) != cudaSuccess) {
      cudaGetLastError();
      ts->d_ent = nullptr;
      ts->unsupported = true;
      return B200_OK;
    }
    ts->ent_cap = cap;
  }
  if ((size_t)nchunks + 1 > ts->chunk_cap) {
    fr(ts->d_chunk_row);
    fr(ts->d_split_rows);
    ts->d_chunk_row = nullptr;
    ts->d_split_rows = nullptr;
    ts->chunk_cap = 0;
    const size_t cap = (size_t)nchunks + 1 + (size_t)nchunks / 16;
    CUDA_TRY(cudaMalloc(&ts->d_chunk_row, cap * 4));
    CUDA_TRY(cudaMalloc(&ts->d_split_rows, cap * 4));
    ts->chunk_cap = cap;
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_ent, 0xff, ((size_t)S + 64) * 16, st));
  CUDA_TRY(cudaMemsetAsync(ts->d_chunk_row, 0, (size_t)(nchunks + 1) * 4, st));
  {
    const int nyh = p->g.nf[DIM - 2] / 2, nbx = num_xtiles<DIM>(p->g);
    const int npos = DIM == 3 ? p->g.nf[0] : nyh;          // tiles per column
    const int wpc = (npos + 31) / 32;
    const size_t nwords = (size_t)(DIM == 3 ? (size_t)nyh * nbx : nbx) * wpc;
    if (nwords > ts->empty_cap) {
      fr(ts->d_empty);
      ts->d_empty = nullptr;
      CUDA_TRY(cudaMalloc(&ts->d_empty, nwords * 4));
      ts->empty_cap = nwords;
    }
    CUDA_TRY(cudaMemsetAsync(ts->d_empty, 0, nwords * 4, st));
    k_mark_empty<DIM><<<ceil_div(nrows, 256), 256, 0, st>>>(p->g, nrows, ts->d_tot, ts->d_empty, wpc);
    CHECK_LAUNCH();
  }
  k_build_stream<DIM, W><<<ceil_div(nrows * 32, 256), 256, 0, st>>>(
      p->g, nrows, p->d_bin_start, ts->d_tot, ts->d_start, ts->d_rec, ts->d_ent, ts->d_chunk_row,
      ts->d_split_rows, ts->d_counters + 1, (uint32_t)ts->lch);
  CHECK_LAUNCH();
  int nsplit = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsplit, ts->d_counters + 1, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  ts->S = S;
  ts->nchunks = nchunks;
  ts->nsplit = nsplit;
  ts->nvis = (long long)grand;
  return B200_OK;
}

template <int DIM, int W>
int prepare(b200_plan* p, RowsState* ts, cudaStream_t st) {
  const long long M = p->M;
  if (!ts->d_counters) CUDA_TRY(cudaMalloc(&ts->d_counters, 64));
  if ((size_t)M > ts->pts_cap || !ts->d_iperm) {
    if (ts->d_iperm) cudaFree(ts->d_iperm);
    if (ts->d_ptab) cudaFree(ts->d_ptab);
    if (ts->d_rec) cudaFree(ts->d_rec);
    ts->d_iperm = nullptr;
    ts->d_ptab = nullptr;
    ts->d_rec = nullptr;
    ts->pts_cap = 0;
    const size_t cap = (size_t)(M > 0 ? M : 1);
    CUDA_TRY(cudaMalloc(&ts->d_ptab, cap * 8 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&ts->d_iperm, cap * 4));
    // the records (112 B per point) only feed the stream builder, but stay allocated: a
    // cudaMalloc / cudaFree pair of that size per update_samples costs up to tens of ms
    CUDA_TRY(cudaMalloc(&ts->d_rec, cap * REC * sizeof(float)));
    ts->pts_cap = cap;
  }
  if (M > 0) {
    k_point_records<W><<<ceil_div(M, 256), 256, 0, st>>>(
        p->g, M, p->d_poly, p->d_org_s[0], p->d_org_s[1], p->d_org_s[2], p->d_x1_s[0],
        p->d_x1_s[1], p->d_x1_s[2], ts->d_rec, ts->d_ptab);
    CHECK_LAUNCH();
    k_invert_perm<<<ceil_div(M, 256), 256, 0, st>>>(M, p->d_perm, ts->d_iperm);
    CHECK_LAUNCH();
  }
  B200_TRY((build_stream<DIM, W>(p, ts, st)));
  ts->M = M;
  ts->valid = true;
  return B200_OK;
}

template <int DIM, int W, bool SPREAD, bool FIXED>
int launch_rows_v(b200_plan* p, RowsState* ts, float2* fw, int T, cudaStream_t st) {
  auto kern = k_rows<DIM, W, SPREAD, FIXED>;
  const size_t smem = (size_t)WARPS * smem_per_warp(SPREAD);
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
  const long long want = (ts->nchunks + WARPS - 1) / WARPS;
  const long long cap = (long long)p->num_sms * ctas_per_sm;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  // slot 4 of the plan's timing events brackets the row kernel alone (bench.py roofline)
  const bool timed = p->timing && p->ev_ok;
  if (timed) cudaEventRecord(p->ev[8], st);
  kern<<<grid, THREADS, smem, st>>>(p->g, T, ts->nchunks, ts->S, p->M, ts->d_ent, ts->d_chunk_row,
                                    ts->d_ptab, ts->d_kt, fw, ts->d_counters, p->rows_dbg,
                                    (SPREAD && p->spread_may_skip_empty) ? 1 : 0, 31 - __builtin_clz(ts->lch));
  if (SPREAD) {
    p->spread_empty = p->spread_may_skip_empty ? ts->d_empty : nullptr;
    p->empty_nyh = p->g.nf[DIM - 2] / 2;
    p->empty_nbx = num_xtiles<DIM>(p->g);
  }
  if (timed) {
    cudaEventRecord(p->ev[9], st);
    p->ev_used[4] = 1;
  }
  CHECK_LAUNCH();
  return B200_OK;
}

template <int DIM, int W, bool SPREAD>
int launch_rows(b200_plan* p, RowsState* ts, float2* fw, int T, cudaStream_t st) {
  if (ts->lch == LCH) return launch_rows_v<DIM, W, SPREAD, true>(p, ts, fw, T, st);
  return launch_rows_v<DIM, W, SPREAD, false>(p, ts, fw, T, st);
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

// new sample locations: keep the (grow-only) buffers, rebuild their contents on the next execute
void tiled_invalidate(b200_plan* p) {
  if (p->tiled) ((RowsState*)p->tiled)->valid = false;
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
  fr(ts->d_start);
  fr(ts->d_ent);
  fr(ts->d_chunk_row);
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

// Bit strings of the tiles that no point visits (nullptr if the row kernels do not serve this
// trajectory): the type-2 FFT may leave those tiles unwritten, the row interpolator never reads them.
const uint32_t* tiled_empty_bits(b200_plan* p, cudaStream_t st) {
  if (ensure_state(p, st) != B200_OK) return nullptr;
  RowsState* ts = state(p);
  if (ts->unsupported || !ts->d_empty) return nullptr;
  const int d = p->g.dim;
  p->empty_nyh = p->g.nf[d - 2] / 2;
  p->empty_nbx = (p->g.nf[d - 1] + CX - 1) / CX;
  return ts->d_empty;
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
