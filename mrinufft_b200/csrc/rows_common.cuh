// rows_common.cuh -- "method 2": output-owned, register-accumulating spreading (K2) and its transpose
// for interpolation (K3), written for sm_100a.  Shared by spread_rows.cu (coil class 32, host API) and
// spread_rows_cls.cu (one translation unit per smaller coil class).
//
// Why not the usual shared-memory sub-grid + atomicAdd design: a 3-D width-7 kernel needs 343
// complex accumulations per point per coil; with 32 coils batched that is 1.8e11 float atomics per
// transform, and on sm_100a a float atomicAdd on shared memory is an ATOMS.CAST.SPIN
// compare-and-swap loop (checked with cuobjdump).  Shared memory cannot feed the FMA pipe either
// (128 B/clk/SM against 128 FFMA/clk/SM).  The register file can.  So every fine-grid TILE is OWNED by
// one warp at a time and lives in REGISTERS: 32 64-bit registers per lane hold the (cell 2j, cell 2j+1)
// pairs of the real and imaginary parts of 2 grid rows x 16 consecutive cells along the fastest axis.
//
// COIL CLASSES.  What the 32 lanes of the warp are is decided by the number of coils T of the call:
//   class TC = 32 (17..32 coils): lane = coil; the warp's tile is 2 rows x 16 cells at fixed z;
//   class TC < 32 (T <= TC)     : lane = (row group g, coil t), g = lane / TC, t = lane % TC; the
//       G = 32 / TC groups hold GY row pairs x GZ planes of a larger tile (2 GY rows x GZ planes x 16
//       cells), so that a point visits (W + 2 GY - 1) / (2 GY) x (W + GZ - 1) / GZ tiles instead of
//       (W + 1) / 2 x W and the cost of a transform follows its coil count (one coil: 3.1 visits per
//       point instead of 28).  A lane forms its own two row scales from per-visit windows
//       wy[2 GY], wz[GZ] (zero outside the point's footprint); the interpolator adds the partial sums
//       of the row groups with butterfly shuffles before its one red.global per coil.
//
// The points whose footprint covers a tile ("visits") are found through the bin sort of K1 without any
// search:
//   * bins are pencils of 1 x 1 x 16 cells, split in two sub-bins: points whose footprint stays
//     inside the 16-cell tile ("interior") and points whose footprint crosses into the next tile
//     ("crossing"); key order (z0, x-tile, crossing, y0);
//   * the visits of a tile are therefore exactly 3 contiguous ranges of sorted points per origin plane
//     z0 (own interior, own crossing, left neighbour's crossing), each spanning the origin rows y0 whose
//     footprint reaches the tile (twice that when the range wraps periodically);
//   * once per trajectory and coil class the ranges are written out as the VISIT STREAM: tile after
//     tile, one header entry per tile followed by one entry per visit holding everything the inner
//     loop needs that is not a property of the point alone;
//   * the row kernel consumes fixed-size chunks of the stream handed out by one atomic counter; per
//     block of 32 entries it stages a packet per visit in shared memory with cp.async (entry + the
//     point's pair-packed x weights) and, for the spreader, the coil values of the points;
//   * the consume loop is generated PTX (tools/gen_taps.py): one indexed branch per run of visits,
//     packed fma.rn.f32x2 on statically indexed accumulators;
//   * a tile is written to HBM exactly once as coalesced stores (through a shared-memory transpose):
//     no memset of the oversampled grid, no halo flush, no atomics -- except for tiles cut by a chunk
//     boundary, which are accumulated with vector red.global.add on pre-zeroed rows.
//
// Interpolation is the exact transpose: the warp loads its tile into registers once (coalesced),
// every visiting point takes its tap dot products from registers and adds the partial sum into
// the (sorted point, coil) accumulator with one vector `red.global.add.v2.f32` per coil.
//
// Replaces finufft's spread/interp stage (call sites
// src/mrinufft/operators/interfaces/finufft.py:69,76; algorithm docs/explanations/nufft.rst:253-309).
#pragma once
#include <cub/device/device_scan.cuh>
#include <type_traits>

#include "common.cuh"
#include "device_utils.cuh"

namespace rows {

typedef unsigned long long u64;

constexpr int CX = 16;          // cells per tile row (== pencil-bin width, B200_BIN_X)
constexpr int NACC = 32;        // 64-bit accumulator registers per lane: [row 2][re/im 2][cell pair 8]
constexpr int REC = 28;         // floats per point record (112 B):
constexpr int R_WY = 8;         //   [0..7]  pair-packed x weights P | [8..16] 0, wy[0..6], 0
constexpr int R_WZ = 17;        //   [17..23] wz[0..6] | [24] (xo >> 1) + W/2 | [25] y0 | [26..27] pad
constexpr int R_JY = 24;
constexpr int MBLK = 32;        // visits per packet block (one lane stages one visit)
constexpr int WARPS = 4;        // warps per CTA
constexpr int THREADS = WARPS * 32;
constexpr int LCH = 2048;       // stream entries per chunk (= one unit of dynamically scheduled work)
constexpr int YB = 32;          // tile ids sweep z inside slabs of YB rows (see decode_tile)
constexpr int NCLS = 6;         // coil classes 1, 2, 4, 8, 16, 32 (index = log2)

// The visit stream: entries of ESZ bytes, tile after tile in tile-id order.
//   class 32 : {s0, s1, idx, s}                      visit: row scales, tap-kernel case, sorted point index
//   class <32: {idx, s, wy[2 GY], wz[GZ], padding}   visit: case, point, the tile's windows of the y / z weights
//   header entry : idx = IDX_HDR, s = tile id -- "the following visits belong to this tile"
//   padding      : all ones (behind the end of the stream; reads as a header)
constexpr unsigned IDX_HDR = 0xffffffffu;
constexpr unsigned IDX_NONE = 0xfffffffeu;  // in registers only: lane beyond the end of the chunk

// Geometry of a coil class.
template <int DIM, int TC>
struct Cls {
  static_assert(TC == 1 || TC == 2 || TC == 4 || TC == 8 || TC == 16 || TC == 32, "coil class");
  static constexpr int G = 32 / TC;                                                  // row groups per warp
  static constexpr int GZ = DIM == 3 ? (G >= 32 ? 8 : G >= 8 ? 4 : G >= 2 ? 2 : 1) : 1;  // planes per tile
  static constexpr int GY = G / GZ;                                                  // row pairs per tile
  static constexpr int TY = 2 * GY;                                                  // rows per tile
  static constexpr bool GENERIC = TC != 32;
  static constexpr int ESZ = GENERIC ? (8 + 4 * TY + 4 * GZ + 15) / 16 * 16 : 16;    // bytes per stream entry
  static constexpr int EU = ESZ / 16;                                                // ... in uint4 units
  static constexpr int IS_OFF = GENERIC ? 0 : 8;                                     // byte offset of {idx, s}
  static constexpr int PKT = 32 + ESZ;                                               // staged packet: P0..P3 | entry
  // The packets of a block of MBLK visits are staged PIECE-MAJOR: 16-byte piece q of visit k sits at
  // q * PSTR + 16 k, so that a cp.async instruction (one lane = one visit) writes 512 contiguous bytes = 4
  // shared-memory wavefronts.  (Packet-major, lane stride PKT = 48 .. 80 bytes, every 128-byte line the 32 copies
  // touched was a wavefront: 12 .. 32 per instruction, ncu `L1 Wavefronts Shared Excessive` -- in class 4 the
  // staging copies were 35 % of all shared-memory wavefronts.)  fa(b): address of packet byte b inside a block.
  static constexpr int PSTR = 16 * 32;
  __host__ __device__ static constexpr unsigned fa(int b) { return (unsigned)((b >> 4) * PSTR + (b & 15)); }
  // PROD: the entry has room for the 2 G row scales wy[2 gy + r] wz[gz] themselves (3-D class 16, the 2-D
  // classes, whose z weight is 1): a lane loads its pair instead of forming it (one load and two multiplies
  // less per visit); entry = {idx, s, p[gz GY + gy][r]}
  static constexpr bool PROD = GENERIC && 8 + 8 * G <= ESZ;
  // DIRECT: every lane sends its partial sum to k-space itself (one red per lane and visit, like class 32)
  // instead of parking it for a reduction over the row groups: with only two groups the second red is cheaper
  // than the store, the re-reads and the extra pass (15.5 -> 13.3 ms at cfg-C geometry; with the four groups of
  // class 8 the two ways cost the same, 8.9 ms, and parking keeps the L2 reduction traffic at a quarter)
  static constexpr bool DIRECT = GENERIC && TC == 16;
  static constexpr int WY_OFF = 32 + 8;                                              // packet offsets of the windows
  static constexpr int WZ_OFF = 32 + 8 + 4 * TY;
  static constexpr int SB = TC >= 16 ? 16 : 32;      // visits per value block of the spreader (cp.async granularity)
  static constexpr int LPP = TC >= 2 ? TC / 2 : 1;   // lanes that copy one point's coil row (16 bytes each; TC = 1: 8)
  static constexpr int PPI = 32 / LPP;               // points per cp.async instruction
  static constexpr int VCP = TC >= 2 ? 16 : 8;       // bytes per lane and copy
  static constexpr int NPB = YB / TY;                // tiles per y-block along y
  static constexpr int PG = NPB / 4;                 // groups of 4 vertically adjacent tiles per block
  // per-warp shared memory (bytes)
  static constexpr int SM_VBUF = 2 * SB * TC * 8;    // double-buffered coil values of SB points (spreader only)
  static constexpr int SM_META = 2 * MBLK * PKT;     // double-buffered packets
  static constexpr int TBS = 10;                     // class 32: float stride of the transpose planes [32 coils][8 cells]
  static constexpr int TFS = 18;                     // class 32: flush planes [16 coils][16 cells], float stride
  static constexpr int TCS = 144;                    // class < 32: byte stride of the transpose rows [32 lanes][16 cells] of
                                                     // interleaved (re, im): 16-byte aligned, conflict-free 128-bit column reads
  // interpolator, class < 32: partial sums of OBV visits, [visit][lane] (re, im) in rows of OBS bytes (the
  // skew keeps the column reads of the reduction -- half a warp = 16 / TC visits x TC coils -- off each
  // other's banks; tools/gen_taps.py `obuf_stride`); shares the transpose planes' memory (a tile is loaded
  // between runs of visits, never during one)
  static constexpr int OBV = 16, OBS = TC >= 16 ? 256 : 256 + 8 * TC;
  static constexpr int SM_TBUF = GENERIC ? (32 * TCS > OBV * OBS ? 32 * TCS : OBV * OBS) : 2 * 32 * TBS * 4;
  static_assert(!GENERIC || 2 * 32 * TFS * 4 <= SM_TBUF, "the interpolator's planes [2][32 lanes][TFS] share that memory");
  static_assert(GENERIC || 2 * 16 * TFS * 4 <= SM_TBUF, "flush planes must fit in the transpose buffer");
  __host__ __device__ static constexpr int smem_per_warp(bool spread) { return (spread ? SM_VBUF : 0) + SM_META + SM_TBUF; }
};

// Stream of one coil class (built lazily, on the first transform that needs it).
struct StreamState {
  int32_t* d_tot = nullptr;        // [nrows + 1] visits per tile (-1: tile id outside the grid)
  uint32_t* d_start = nullptr;     // [nrows + 1] stream position of a tile's header entry
  uint4* d_ent = nullptr;          // [(S + slack) * EU] the visit stream
  int32_t* d_chunk_row = nullptr;  // [nchunks] tile owning the first entry of a chunk
  int32_t* d_split_rows = nullptr; // tiles cut by a chunk boundary (accumulated with red.add)
  size_t ent_cap = 0, chunk_cap = 0;  // capacities (entries, chunks): grow-only
  long long nrows = 0, nsplit = 0, nvis = 0;
  unsigned S = 0;                  // stream length in entries
  int nchunks = 0;
  int lch = LCH;                   // entries per chunk of this stream (smaller for short streams, see build_stream)
  bool valid = false;
  bool unsupported = false;        // too many visits / points for the 32-bit stream words
};

struct RowsState {
  float* d_rec = nullptr;        // [M][REC] per sorted point
  float2* d_kt = nullptr;        // [M][TC] transposed (sorted point, coil) k-space batch of the running call
  size_t kt_bytes = 0;
  int32_t* d_iperm = nullptr;    // [M] point index -> sorted position
  float* d_ptab = nullptr;       // [M][8] pair-packed x weights per sorted point
  uint32_t* d_empty = nullptr;   // empty-tile bit strings of the class-32 tiling (k_mark_empty)
  size_t empty_cap = 0;
  bool empty_valid = false;
  int* d_counters = nullptr;     // [0] work counter, [1] split-row counter, [2..3] total visits (u64)
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  size_t pts_cap = 0;            // capacity in points: grow-only
  long long M = -1;
  bool valid = false;            // point records are those of the plan's current points
  StreamState cls[NCLS];
};

// ------------------------------------------------------------------------------ geometry helpers
template <int DIM>
__host__ __device__ __forceinline__ int num_xtiles(const Geom& g) {
  return (g.nf[DIM - 1] + CX - 1) / CX;
}

// Tile ids enumerate (y-block of YB rows, z-tile, group of 4 tiles inside the block, x-tile, tile
// inside the group): the 4 ids fetched together are 4 vertically adjacent tiles, and the sweep over z
// stays inside a slab of YB rows, so that the coil rows / records of the points (re-visited by the
// next planes) are still in L2: one z step streams YB * nfx * 8 B * T = 4 MB of grid, not a whole
// 67 MB plane.
template <int DIM, int TC>
__host__ __device__ __forceinline__ long long num_tiles(const Geom& g) {
  using C = Cls<DIM, TC>;
  const long long nyb = (g.nf[DIM - 2] + YB - 1) / YB;
  const long long nzt = DIM == 3 ? (g.nf[0] + C::GZ - 1) / C::GZ : 1;
  return nyb * nzt * C::PG * num_xtiles<DIM>(g) * 4;
}

struct TileCoord {
  int z, y, bx;  // first plane, first (even) row, x-tile
};

template <int DIM, int TC>
__device__ __forceinline__ bool decode_tile(const Geom& g, long long row, TileCoord* rc) {
  using C = Cls<DIM, TC>;
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int nzt = DIM == 3 ? (g.nf[0] + C::GZ - 1) / C::GZ : 1;
  const int ps = (int)(row & 3);
  long long r = row >> 2;
  rc->bx = (int)(r % nbx);
  r /= nbx;
  const int pg = (int)(r % C::PG);
  r /= C::PG;
  rc->z = (int)(r % nzt) * C::GZ;
  const int yb = (int)(r / nzt);
  rc->y = yb * YB + (pg * 4 + ps) * C::TY;
  return rc->y < nfy;  // nfy is even: both rows of a pair are valid or neither
}

// Range slot -> [begin, begin + len) in sorted point order.
//   slot = ((zs * 3) + sub) * 2 + part ; sub 0: own interior, 1: own crossing, 2: left crossing ;
//   zs: origin plane z0 = z + GZ - 1 - zs (periodic) ;
//   part 0: y0 in [max(y-w+1, 0), y+TY-1], part 1: the periodic wrap [y-w+1+nfy, nfy-1] (if any).
// (grids are at least W + tile extent - 1 cells along y and z, so no point is seen twice)
template <int DIM, int W, int TC>
__device__ __forceinline__ void slot_range(const Geom& g, const TileCoord& rc, int slot,
                                           const int32_t* __restrict__ bin_start, int* begin,
                                           int* len) {
  using C = Cls<DIM, TC>;
  constexpr int NZS = (DIM == 3) ? W + C::GZ - 1 : 1;
  *begin = 0;
  *len = 0;
  if (slot >= NZS * 6) return;
  const int part = slot & 1;
  const int sub = (slot >> 1) % 3;
  const int zs = (slot >> 1) / 3;
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  const int ylo = rc.y - (W - 1);
  int a, b;
  if (part == 0) {
    a = ylo > 0 ? ylo : 0;
    b = rc.y + C::TY - 1;
    if (b > nfy - 1) b = nfy - 1;
  } else {
    if (ylo >= 0) return;
    a = ylo + nfy;
    b = nfy - 1;
  }
  int z0 = 0;
  if (DIM == 3) {
    z0 = rc.z + (C::GZ - 1) - zs;
    if (z0 < 0) z0 += g.nf[0];
    else if (z0 >= g.nf[0]) return;  // planes of a short last tile that do not exist
  }
  const int bxx = (sub == 2) ? (rc.bx == 0 ? nbx - 1 : rc.bx - 1) : rc.bx;
  const int cross = sub != 0;
  const long long kb = (((long long)z0 * nbx + bxx) * 2 + cross) * nfy;
  const int s0 = __ldg(bin_start + kb + a);
  *begin = s0;
  *len = __ldg(bin_start + kb + b + 1) - s0;
}

// ------------------------------------------------------------------------------ stream construction
// visits per tile (-1 for ids outside the grid) and their grand total
template <int DIM, int W, int TC>
__global__ void __launch_bounds__(256)
k_row_totals(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
             int32_t* __restrict__ tot, unsigned long long* __restrict__ grand) {
  using C = Cls<DIM, TC>;
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = 0;
  TileCoord rc;
  const bool ok = row < nrows && decode_tile<DIM, TC>(g, row, &rc);
  if (ok) {
    constexpr int NZS = (DIM == 3) ? W + C::GZ - 1 : 1;
    for (int slot = 0; slot < NZS * 6; ++slot) {
      int b, l;
      slot_range<DIM, W, TC>(g, rc, slot, bin_start, &b, &l);
      total += l;
    }
  }
  if (row <= nrows) tot[row] = ok ? (int32_t)min(total, (long long)INT32_MAX) : -1;
  long long wsum = total;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, d);
  if ((threadIdx.x & 31) == 0 && wsum > 0) atomicAdd(grand, (unsigned long long)wsum);
}

// stream entries per tile: header + visits (0 for tile ids outside the grid)
static __global__ void __launch_bounds__(256)
k_scan_inputs(long long n, const int32_t* __restrict__ tot, uint32_t* __restrict__ words) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tot[i];
  words[i] = t < 0 ? 0u : (uint32_t)t + 1u;
}

// One stream entry.  `zs`: origin-plane slot of the range the point came from, `left`: it is a crossing
// point of the left neighbour tile.
template <int DIM, int W, int TC>
__device__ __forceinline__ unsigned entry_case(const float* __restrict__ rec, int s, int left, int lhalf) {
  const int jx = __float_as_int(__ldg(rec + (long long)s * REC + R_JY));
  return (unsigned)(jx - left * lhalf);
}

template <int DIM, int W, int TC>
__device__ __forceinline__ void write_entry(const Geom& g, const TileCoord& rc, const float* __restrict__ rec,
                                            int s, int zs, int left, int lhalf, uint4* __restrict__ dst) {
  using C = Cls<DIM, TC>;
  const float* r = rec + (long long)s * REC;
  const int2 jy = *reinterpret_cast<const int2*>(r + R_JY);
  const int nfy = g.nf[DIM - 2];
  const unsigned idx = (unsigned)(jy.x - left * lhalf);
  int dy = rc.y - jy.y;  // row offset of the tile's first row inside the footprint
  if (dy < -(C::TY - 1)) dy += nfy;
  if constexpr (!C::GENERIC) {
    // row y takes wy[dy], row y+1 takes wy[dy+1]  (dy in [-1, W-1]; the record stores 0, wy[0..6], 0 so
    // that both loads are unconditional)
    const float wz = (DIM == 3) ? r[R_WZ + zs] : 1.f;
    dst[0] = make_uint4(__float_as_uint(r[R_WY + 1 + dy] * wz), __float_as_uint(r[R_WY + 2 + dy] * wz), idx,
                        (unsigned)s);
  } else {
    unsigned wd[C::EU * 4];
#pragma unroll
    for (int i = 0; i < C::EU * 4; ++i) wd[i] = 0u;
    wd[0] = idx;
    wd[1] = (unsigned)s;
    float wyk[C::TY], wzk[C::GZ];
#pragma unroll
    for (int k = 0; k < C::TY; ++k) {
      const int d = dy + k;
      const bool in = d >= 0 && d < W && rc.y + k < nfy;
      wyk[k] = in ? r[R_WY + 1 + (in ? d : 0)] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < C::GZ; ++k) {
      if (DIM == 3) {
        const int d = zs - (C::GZ - 1) + k;
        const bool in = d >= 0 && d < W && rc.z + k < g.nf[0];
        wzk[k] = in ? r[R_WZ + (in ? d : 0)] : 0.f;
      } else {
        wzk[k] = 1.f;
      }
    }
    if constexpr (C::PROD) {
#pragma unroll
      for (int q = 0; q < C::G; ++q) {
        wd[2 + 2 * q] = __float_as_uint(wyk[2 * (q % C::GY)] * wzk[q / C::GY]);   // q = row group of a lane
        wd[3 + 2 * q] = __float_as_uint(wyk[2 * (q % C::GY) + 1] * wzk[q / C::GY]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < C::TY; ++k) wd[2 + k] = __float_as_uint(wyk[k]);
#pragma unroll
      for (int k = 0; k < C::GZ; ++k) wd[2 + C::TY + k] = __float_as_uint(wzk[k]);
    }
#pragma unroll
    for (int q = 0; q < C::EU; ++q) dst[q] = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
  }
}

// The visit stream, written once per trajectory and coil class.  One warp per tile writes the tile's
// header entry and, behind it, one entry per visit.  It also records which tile owns the first entry of
// every chunk and which tiles are cut by a chunk boundary.
template <int DIM, int W, int TC>
__global__ void __launch_bounds__(256)
k_build_stream(Geom g, long long nrows, const int32_t* __restrict__ bin_start,
               const int32_t* __restrict__ tot, const uint32_t* __restrict__ start,
               const float* __restrict__ rec, uint4* __restrict__ ent,
               int32_t* __restrict__ chunk_row, int32_t* __restrict__ split_rows,
               int* __restrict__ split_counter, uint32_t lch) {
  using C = Cls<DIM, TC>;
  constexpr int NZS = (DIM == 3) ? W + C::GZ - 1 : 1;
  constexpr int NR = (NZS * 6 + 31) / 32;  // rounds of 32 range slots
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const int t = tot[row];
  if (t < 0) return;
  const uint32_t hs = start[row], he = hs + 1u + (uint32_t)t;
  if (lane == 0) {
    uint4* h = ent + (size_t)hs * C::EU;
    if constexpr (!C::GENERIC) {
      h[0] = make_uint4(0u, 0u, IDX_HDR, (uint32_t)row);
    } else {
      h[0] = make_uint4(IDX_HDR, (uint32_t)row, 0u, 0u);
      for (int q = 1; q < C::EU; ++q) h[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (uint32_t c = (hs + lch - 1) / lch; c * lch < he; ++c) chunk_row[c] = (int32_t)row;
    if (hs / lch != (he - 1) / lch) split_rows[atomicAdd(split_counter, 1)] = (int32_t)row;
  }
  if (t == 0) return;
  TileCoord rc;
  decode_tile<DIM, TC>(g, row, &rc);
  const int nfx = g.nf[DIM - 1];
  const int nbx = num_xtiles<DIM>(g);
  // a left neighbour's crossing point lands at x offset (xo - length of the left tile)
  const int lhalf = ((rc.bx == 0) ? (nfx - (nbx - 1) * CX) : CX) >> 1;
  uint4* out = ent + ((size_t)hs + 1) * C::EU;
  int b[NR], l[NR], pre[NR];
  int run = 0;
  bool any_long = false;
#pragma unroll
  for (int h = 0; h < NR; ++h) {
    slot_range<DIM, W, TC>(g, rc, lane + 32 * h, bin_start, &b[h], &l[h]);
    int inc = l[h];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += u;
    }
    pre[h] = run + inc - l[h];
    run += __shfl_sync(0xffffffffu, inc, 31);
    any_long |= l[h] > 64;
  }
  any_long = __any_sync(0xffffffffu, any_long);
  // Tiles made of short ranges only (all but the dense k-space centre): the visits are grouped by
  // their tap-kernel case `idx` with a counting sort, so that the consume loop runs through
  // straight-line code for whole runs of visits (tools/gen_taps.py).  Deterministic: inside a case
  // the order is (lane, slot, position).
  constexpr int NC = 8 + W / 2;  // number of cases
  __shared__ int s_cnt[8][NC][32];
  const int wib = threadIdx.x >> 5;
  if (!any_long) {
#pragma unroll
    for (int c = 0; c < NC; ++c) s_cnt[wib][c][lane] = 0;
#pragma unroll
    for (int h = 0; h < NR; ++h) {
      const int slot = lane + 32 * h;
      const int left = ((slot >> 1) % 3) == 2 ? 1 : 0;
      for (int i = 0; i < l[h]; ++i) s_cnt[wib][entry_case<DIM, W, TC>(rec, b[h] + i, left, lhalf)][lane] += 1;
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
    for (int h = 0; h < NR; ++h) {
      const int slot = lane + 32 * h;
      const int zs = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
      for (int i = 0; i < l[h]; ++i) {
        const unsigned cs = entry_case<DIM, W, TC>(rec, b[h] + i, left, lhalf);
        const int pos = s_cnt[wib][cs][lane]++;
        write_entry<DIM, W, TC>(g, rc, rec, b[h] + i, zs, left, lhalf, out + (size_t)pos * C::EU);
      }
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < NR; ++h) {
    const int slot = lane + 32 * h;
    const int zs = (slot >> 1) / 3, left = ((slot >> 1) % 3) == 2 ? 1 : 0;
    // short ranges: written by the owning lane; long ranges (dense k-space centre): by the whole warp
    const bool is_long = l[h] > 64;
    if (!is_long)
      for (int i = 0; i < l[h]; ++i)
        write_entry<DIM, W, TC>(g, rc, rec, b[h] + i, zs, left, lhalf, out + (size_t)(pre[h] + i) * C::EU);
    unsigned longs = __ballot_sync(0xffffffffu, is_long);
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const int bb = __shfl_sync(0xffffffffu, b[h], src);
      const int ll = __shfl_sync(0xffffffffu, l[h], src);
      const int pp = __shfl_sync(0xffffffffu, pre[h], src);
      const int szs = __shfl_sync(0xffffffffu, zs, src);
      const int sl = __shfl_sync(0xffffffffu, left, src);
      for (int i = lane; i < ll; i += 32)
        write_entry<DIM, W, TC>(g, rc, rec, bb + i, szs, sl, lhalf, out + (size_t)(pp + i) * C::EU);
    }
  }
}

// rows shared by several work items are accumulated with red.add: zero them first
template <int DIM, int TC>
__global__ void __launch_bounds__(128)
k_zero_split_rows(Geom g, int T, long long nsplit, const int32_t* __restrict__ split_rows,
                  float2* __restrict__ fw) {
  using C = Cls<DIM, TC>;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nsplit) return;
  TileCoord rc;
  if (!decode_tile<DIM, TC>(g, split_rows[w], &rc)) return;
  const int nfx = g.nf[DIM - 1], nfy = g.nf[DIM - 2];
  const int nz = DIM == 3 ? g.nf[0] : 1;
  const int x = rc.bx * CX + (lane & 15);
  if (x >= nfx) return;
  for (int k = 0; k < C::GZ; ++k) {
    if (rc.z + k >= nz) break;
    for (int r = lane >> 4; r < C::TY; r += 2) {
      if (rc.y + r >= nfy) break;
      float2* dst = fw + ((long long)(rc.z + k) * nfy + rc.y + r) * nfx + x;
      for (int t = 0; t < T; ++t) dst[(long long)t * g.nftot] = make_float2(0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------ row kernels
// Per-visit tap kernels: generated inline PTX (tools/gen_taps.py), defined by the including translation
// unit for its coil class.
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_spread(u64 (&acc)[NACC], unsigned pk, int n, unsigned vb, unsigned yo,
                                                 unsigned zo);
// (class 32 adds to k-space through `ktl`; the smaller classes park partial sums at `ob`, see OBUF below)
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_interp(u64 (&acc)[NACC], unsigned pk, int n, const void* ktl, unsigned ob,
                                                 unsigned yo, unsigned zo);

template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>());
    static_for<N, I + 1>(f);
  }
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
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ u64 lds64(unsigned addr) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];\n" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// 4-byte copy, `n` (0 or 4) bytes read from global memory and the rest zero-filled
__device__ __forceinline__ void cp_async4_zfill(unsigned smem_dst, const void* gsrc, int n) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ float2 lds64f(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
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
// `unread` (interpolator, class < 32): bit strings of the class-32 tiles that the producer of the grid left
// unwritten (they have no visitors); their share of a larger tile is taken as zero instead of being read.
template <int DIM, int W, bool SPREAD, bool FIXED, int TC>
__global__ void __launch_bounds__(THREADS, 4)
k_rows(Geom g, int T, int nchunks, unsigned S, long long M, const uint4* __restrict__ ent,
       const int32_t* __restrict__ chunk_row, const float* __restrict__ ptab,
       float2* __restrict__ kt, float2* __restrict__ fw, int* __restrict__ counter, int dbg, int skip_empty,
       int lch_log2, const uint32_t* __restrict__ unread) {
  using C = Cls<DIM, TC>;
  constexpr bool GEN = C::GENERIC;
  constexpr int EU = C::EU, PKT = C::PKT, TBS = C::TBS, TFS = C::TFS, TCS = C::TCS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int SMW = C::smem_per_warp(SPREAD);
  constexpr unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = lane >> 4, cl = lane & 15;
  unsigned char* wsm = smem_raw + (size_t)warp * SMW;
  // (made opaque: nvcc otherwise re-derives these addresses from %tid at every use)
  asm volatile("" : "+l"(wsm));
  const unsigned vbuf_a = smem_u32(wsm);                         // [2][SB][TC] (re, im); spreader only
  unsigned char* metap = wsm + (SPREAD ? C::SM_VBUF : 0);
  const unsigned meta_a = smem_u32(metap);                       // [2] blocks of MBLK packets of PKT bytes, piece-major
  // transpose planes.  Class 32: [32 coils][TBS], lane = coil on the register side; on the grid side a warp
  // instruction moves 8 cells (64 contiguous bytes) of 4 coils.  Smaller classes: one array [32 lanes][16
  // cells] of (re, im) in rows of TCS bytes.
  float* tre = reinterpret_cast<float*>(metap + C::SM_META);
  float* tim = tre + 32 * TBS;                     // (class 32: the flush's [16][TFS] planes start at the same offsets)
  const unsigned tc_a = smem_u32(tre);             // class < 32: the transpose rows ...
  const unsigned ob_a = tc_a;                      // ... which the interpolator's OBUF shares (see Cls)
  const int g4 = lane >> 3, c8 = lane & 7;  // class 32, grid-side role: coil 4 i + g4, cell 8 h + c8
  const int tl = lane & (TC - 1);           // this lane's coil
  const char* ktl = reinterpret_cast<const char*>(kt) +
                    (SPREAD ? (lane & (C::LPP - 1)) * C::VCP : tl * 8);
  asm volatile("" : "+l"(ktl));
  // class < 32: this lane's row group and the packet offsets of its window entries
  const int grp = lane / TC, gyi = grp % C::GY, gzi = grp / C::GY;
  // (block-relative addresses of the lane's window entries in the piece-major staging, see Cls::fa)
  const unsigned yo = C::fa(C::WY_OFF + 8 * (C::PROD ? grp : gyi)), zo = C::fa(C::WZ_OFF + 4 * gzi);

  const int nfx = g.nf[DIM - 1];
  const int nfy = g.nf[DIM - 2];
  const int nbx = num_xtiles<DIM>(g);
  u64* fw64 = reinterpret_cast<u64*>(fw);

  // accumulators: acc[r*16 + c*8 + j] = (cell 2j, cell 2j+1) of row r, c = re / im
  // (spreader: zero here and again after every flush; interpolator: loaded per tile)
  u64 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0ull;

  for (;;) {
    int c = 0;
    if (lane == 0) c = atomicAdd(counter, 1);
    c = __shfl_sync(FULL, c, 0);
    if (c >= nchunks) break;
    const int lg = FIXED ? 11 : lch_log2;
    static_assert(LCH == 2048, "FIXED chunks are 2^11 entries");
    const unsigned base = (unsigned)c << lg;
    const uint4* v = ent + (size_t)base * EU;
    const int nw = (int)min(1u << lg, S - base);
    // consume blocks: SB entries for the spreader (the granularity of its value copies), a whole packet
    // block of 32 for the interpolator (half the per-block bookkeeping)
    constexpr int SB = SPREAD ? C::SB : MBLK, SPB = MBLK / SB;
    const int nsub = (nw + SB - 1) / SB;
    // is the entry behind the chunk a header (or the end of the stream)?  Then the last tile ends here.
    const bool tail_whole = __ldg(reinterpret_cast<const unsigned*>(v + (size_t)nw * EU) + C::IS_OFF / 4) == IDX_HDR;

    // ---- the tile in the accumulators
    u64* gbase = nullptr;  // class 32, grid-side role of this lane: cell c8 of coil g4, row 0 of the tile
    u64* fbase = nullptr;  // class 32: ... in the spreader's flush: cell cl of coil hl; class < 32: cell cl, coil 0
    u64* fb_hl = nullptr;  // class < 32: fbase + this half-warp's share of a transposed access (coil hl; one coil: row group hl)
    int xlim = 0;          // cells of this tile inside the grid (16, less for a short last tile)
    unsigned gmask = 0;      // class < 32: row groups of this tile that lie inside the grid
    int ylim = 0, zlim = 0;  // class < 32: rows / planes of this tile inside the grid
    int ucol = 0, uz = 0;    // class < 32: first tile column / plane of this tile in the `unread` bit strings
    auto tile_setup = [&](int row) {
      // 32-bit version of decode_tile (tile ids are below 2^30 here)
      const int nz = DIM == 3 ? g.nf[0] : 1;
      const int nzt = DIM == 3 ? (nz + C::GZ - 1) / C::GZ : 1;
      const int ps = row & 3;
      int r = row >> 2;
      const int bx = r % nbx;
      r /= nbx;
      const int pg = r % C::PG;
      r /= C::PG;
      const int z = (r % nzt) * C::GZ;
      const int yb = r / nzt;
      const int y = yb * YB + (pg * 4 + ps) * C::TY;
      xlim = nfx - bx * CX;
      if constexpr (!GEN) {
        gbase = fw64 + ((long long)z * nfy + y) * nfx + bx * CX + c8 + (long long)g4 * g.nftot;
        if (SPREAD) fbase = fw64 + ((long long)z * nfy + y) * nfx + bx * CX + cl + (long long)hl * g.nftot;
      } else {
        fbase = fw64 + ((long long)z * nfy + y) * nfx + bx * CX + cl;
        fb_hl = fbase + (TC >= 2 ? (long long)hl * g.nftot : (long long)hl * 2 * nfx);
        ylim = nfy - y;
        zlim = nz - z;
        ucol = (y >> 1) * nbx + bx;
        uz = z;
        gmask = __ballot_sync(FULL, lane < C::G && 2 * (lane % C::GY) < ylim && lane / C::GY < zlim);
      }
    };
    tile_setup(__ldg(chunk_row + c));
    bool started = false;  // the tile's header came by in this chunk
    bool loaded = false;   // interpolator: the tile is in the registers
    bool dirty = false;    // some visit was applied to the tile


    // registers -> grid rows.  Class 32 (lane = coil): 16 coils of one row at a time go through the
    // transpose planes; a store instruction writes one full 128-byte line of 2 coils.  Smaller classes:
    // all 32 lanes park one of their two rows, a store instruction writes the 128-byte lines of 2 lanes.
    // shared = false: plain stores (the tile is complete); true: red.add (tile shared with other chunks).
    // (one copy of this code, with the mode as a run-time flag: the kernel has to fit the instruction cache)
    auto flush = [&](bool shared) {
      if constexpr (!GEN) {
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
                const int tl2 = 2 * i + hl;  // coil inside this half
                if (T == 32 || half * 16 + tl2 < T) {
                  const u64 val = pack2(tre[tl2 * TFS + cl], tim[tl2 * TFS + cl]);
                  u64* a = dst + (long long)(2 * i) * g.nftot;
                  // streaming stores: the grid is written once and not read again by this kernel --
                  // keep L2 for the point data (coil rows, x weights) that neighbouring tiles re-read
                  if (!shared) __stcs(a, val);
                  else red_add_f32x2(reinterpret_cast<float2*>(a), val, 1);
                }
              }
            }
            __syncwarp();
          }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const u64 re = acc[r * 16 + j], im = acc[r * 16 + 8 + j];
            sts128(tc_a + (unsigned)(lane * TCS + j * 16),
                   make_uint4((unsigned)re, (unsigned)im, (unsigned)(re >> 32), (unsigned)(im >> 32)));
          }
          __syncwarp();
          if (cl < xlim) {
            static_for<16>([&](auto I) {
              constexpr int i = decltype(I)::value;
              // source lane L = 2 i + hl: coil L % TC of row group L / TC.  Everything but the hl term is a
              // compile-time or warp-uniform quantity (fb_hl carries the hl term, see tile_setup)
              const int L = 2 * i + hl;
              constexpr int T0 = TC >= 2 ? ((2 * i) & (TC - 1)) : 0, G0 = TC >= 2 ? (2 * i) / TC : 2 * i;
              constexpr int Y0 = 2 * (G0 % C::GY), Z0 = G0 / C::GY;
              const int t = T0 + (TC >= 2 ? hl : 0), gg = G0 + (TC >= 2 ? 0 : hl);
              if (t < T && ((gmask >> gg) & 1u)) {
                const u64 val = lds64(tc_a + (unsigned)(L * TCS + cl * 8));
                u64* a = fb_hl + ((long long)T0 * g.nftot + ((long long)Z0 * nfy + Y0 + r) * nfx);
                if (!shared) __stcs(a, val);
                else red_add_f32x2(reinterpret_cast<float2*>(a), val, 1);
              }
            });
          }
          __syncwarp();
        }
      }
    };
    // a tile without visits: plain zero stores
    auto store_zero = [&]() {
      if (cl < xlim && !skip_empty) {
        if constexpr (!GEN) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
            for (int t = hl; t < T; t += 2) __stcs(fbase + (long long)r * nfx + (long long)(t - hl) * g.nftot, 0ull);
        } else {
          for (int zz = 0; zz < C::GZ && zz < zlim; ++zz)
            for (int yy = hl; yy < C::TY && yy < ylim; yy += 2)
              for (int t = 0; t < T; ++t)
                __stcs(fbase + (long long)t * g.nftot + ((long long)zz * nfy + yy) * nfx, 0ull);
        }
      }
    };
    // grid rows -> registers
    auto load_tile = [&]() {
      if constexpr (!GEN) {
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
      } else {
        // row groups of the tile that exist and whose class-32 tile the producer of the grid wrote
        unsigned um = gmask;
        if (unread != nullptr && DIM == 3) {
          bool ok = lane < C::G && ((gmask >> lane) & 1u);
          if (ok) {
            const int wpc = (g.nf[0] + 31) >> 5;
            const int pz = uz + lane / C::GY;
            ok = !((__ldg(unread + (long long)(ucol + (lane % C::GY) * nbx) * wpc + (pz >> 5)) >> (pz & 31)) & 1u);
          }
          um = __ballot_sync(FULL, ok);
        }
        // grid -> planes [32 lanes][TFS] of real and of imaginary parts with 4-byte cp.async (zero fill for
        // what is missing), one of the two rows at a time.  Planar, because the register side must arrive as
        // 64-bit loads of (cell 2j, cell 2j+1): assembled from the halves of interleaved 128-bit loads, ptxas
        // keeps the halves where they landed and pays two moves per FFMA2 operand in the visit loops.
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          static_for<16>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const int L = 2 * i + hl;
            constexpr int T0 = TC >= 2 ? ((2 * i) & (TC - 1)) : 0, G0 = TC >= 2 ? (2 * i) / TC : 2 * i;
            constexpr int Y0 = 2 * (G0 % C::GY), Z0 = G0 / C::GY;
            const int t = T0 + (TC >= 2 ? hl : 0), gg = G0 + (TC >= 2 ? 0 : hl);
            const bool ok = cl < xlim && t < T && ((um >> gg) & 1u);
            const u64* src = ok ? fb_hl + ((long long)T0 * g.nftot + ((long long)Z0 * nfy + Y0 + r) * nfx) : fw64;
            const unsigned dst = tc_a + (unsigned)((L * TFS + cl) * 4);
            cp_async4_zfill(dst, src, ok ? 4 : 0);
            cp_async4_zfill(dst + 32u * TFS * 4u, reinterpret_cast<const char*>(src) + 4, ok ? 4 : 0);
          });
          cp_async_commit();
          cp_async_wait<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[r * 16 + j] = lds64(tc_a + (unsigned)((lane * TFS + 2 * j) * 4));
            acc[r * 16 + 8 + j] = lds64(tc_a + (unsigned)((32 * TFS + lane * TFS + 2 * j) * 4));
          }
          __syncwarp();
        }
      }
    };

    // ---- staging: packets in blocks of 32 entries (one lane = one entry), values in blocks of SB.
    // Pure data movement: a lane only ever holds the {idx, s} half of its entry in registers.
    auto load_is = [&](int mb) -> uint2 {  // {idx, s} of this lane's entry of packet block mb
      const int i = mb * MBLK + lane;
      return i < nw ? __ldg(reinterpret_cast<const uint2*>(v + (size_t)i * EU) + C::IS_OFF / 8)
                    : make_uint2(IDX_NONE, 0u);
    };
    // pull the x-weight and coil-value lines of a block's points into L2 ahead of the cp.async copies
    auto prefetch_points = [&](const uint2& is) {
      if (is.x < IDX_NONE) {
        prefetch_l2(ptab + (long long)is.y * 8);
        if (SPREAD) {
          const char* q = reinterpret_cast<const char*>(kt + (long long)is.y * TC);
          prefetch_l2(q);
          if (TC == 32) prefetch_l2(q + 128);
        }
      }
    };
    // ... and the stream itself, three packet blocks ahead
    auto prefetch_stream = [&](int mb) {
      if constexpr (!GEN) {
        const int i = mb * MBLK + lane * 8;
        if (lane < 4 && i < nw) prefetch_l2(v + i);
      } else {
        constexpr int LINES = MBLK * C::ESZ / 128;  // 128-byte lines per packet block
        const int i = mb * MBLK + (lane * 128) / C::ESZ;
        if (lane < LINES && i < nw) prefetch_l2(reinterpret_cast<const char*>(v) + (size_t)mb * MBLK * C::ESZ + lane * 128);
      }
    };
    // packet block mb: entry -> behind the first 32 bytes of the packet, x weights -> first 32 bytes
    // (cp.async); returns the header mask of the block
    auto stage_block = [&](int mb, const uint2& is) -> unsigned {
      const unsigned row_a = meta_a + (unsigned)((mb & 1) * MBLK * PKT + lane * 16);
      if (is.x != IDX_NONE) {
        const uint4* e = v + (size_t)(mb * MBLK + lane) * EU;
#pragma unroll
        for (int q = 0; q < EU; ++q) cp_async16(row_a + (unsigned)((2 + q) * C::PSTR), e + q);
      }
      if (is.x < IDX_NONE) {
        const float* pw = ptab + (long long)is.y * 8;
        cp_async16(row_a, pw);
        cp_async16(row_a + (unsigned)C::PSTR, pw + 4);
      }
      return __ballot_sync(FULL, is.x == IDX_HDR);
    };
    // coil values of value block j (its entries sit in lanes SB (j % SPB) .. of `is`, the registers of
    // packet block j / SPB): PPI points per instruction, VCP bytes per lane
    // (destination and source-lane bases are formed once per call: inside the predicated copies the
    // compiler re-derived them per point)
    const unsigned vdst0 = vbuf_a + (unsigned)lane * (unsigned)C::VCP;
    auto values_issue = [&](int j, const uint2& is) {
      const int buf = j & 1;
      const unsigned sv = is.x < IDX_NONE ? is.y : IDX_NONE;
      const unsigned vdst = vdst0 + (unsigned)buf * (unsigned)(C::SB * TC * 8);
      const int lane0 = (j % (MBLK / C::SB)) * C::SB + lane / C::LPP;
#pragma unroll
      for (int i = 0; i < C::SB / C::PPI; ++i) {
        const unsigned sk = __shfl_sync(FULL, sv, lane0 + C::PPI * i);
        if (sk != IDX_NONE) {
          if constexpr (TC >= 2) cp_async16(vdst + (unsigned)i * 512u, ktl + (unsigned long long)sk * (unsigned)(TC * 8));
          else cp_async8(vdst + (unsigned)i * 256u, ktl + (unsigned long long)sk * 8u);
        }
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
        if (GEN && SPB == 1) prefetch_points(is_nxt);
      } else {
        prefetch_points(is_nxt);
      }
      if (SPREAD && more && !(dbg & 2)) values_issue(j + 1, is_val);
      cp_async_commit();
      if (more) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncwarp();
      const int n = min(SB, nw - j * SB);
      const unsigned pk_a = meta_a + (unsigned)(((j / SPB) & 1) * MBLK * PKT + (j % SPB) * SB * 16);
      const unsigned vb_a = vbuf_a + (unsigned)((j & 1) * C::SB * TC + tl) * 8u;

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
          if (SPREAD) {
            rows_loop_spread<W, DIM, TC>(acc, pk_a + (unsigned)(k0 * 16), k1 - k0, vb_a + (unsigned)(k0 * TC * 8), yo, zo);
          } else if constexpr (!GEN || C::DIRECT) {
            rows_loop_interp<W, DIM, TC>(acc, pk_a + (unsigned)(k0 * 16), k1 - k0, ktl, 0u, yo, zo);
          } else {
            // class < 32: pieces of OBV visits; every lane parks its partial sums (this lane's coil, its
            // row group) in OBUF[visit][lane], then lane (v, t) adds the G row groups of coil t of visit v
            // and sends the sum to k-space: one coalesced red per TC coils and visit
            for (int kk = k0; kk < k1; kk += C::OBV) {
              const int nv = min(C::OBV, k1 - kk);
              const unsigned pkk = pk_a + (unsigned)(kk * 16);
              rows_loop_interp<W, DIM, TC>(acc, pkk, nv, nullptr, ob_a + (unsigned)lane * 8u, yo, zo);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < (C::OBV * TC + 31) / 32; ++it) {
                const int o = it * 32 + lane, vv = o / TC;
                const unsigned a = ob_a + (unsigned)(vv * C::OBS + (o & (TC - 1)) * 8);
                if (vv < nv) {
                  float2 sum = lds64f(a);
#pragma unroll
                  for (int gg = 1; gg < C::G; ++gg) {
                    const float2 q = lds64f(a + (unsigned)(gg * TC * 8));
                    sum.x += q.x;
                    sum.y += q.y;
                  }
                  const unsigned sp = lds32(pkk + (unsigned)(vv * 16) + C::fa(32 + C::IS_OFF + 4));
                  red_add_f32x2(reinterpret_cast<float2*>(const_cast<char*>(ktl)) + (size_t)sp * TC, pack2(sum.x, sum.y), 1);
                }
              }
              __syncwarp();
            }
          }
        }
        // a header entry, or the end of the chunk: the tile in the registers is finished as far as this chunk
        // goes.  It is complete if its own header came by in this chunk and it does not go on behind it.
        // (the end of the chunk shares the header's code path: one copy of the flush in the kernel)
        const bool at_end = k1 >= n;
        if (at_end && more) break;
        if (SPREAD && !(dbg & 1)) {
          const bool whole = started && (!at_end || tail_whole);
          if (dirty) {
            flush(!whole);
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = 0ull;
          } else if (whole) {
            store_zero();
          }
        }
        if (at_end) break;
        hm &= hm - 1;
        tile_setup((int)lds32(pk_a + (unsigned)(k1 * 16) + C::fa(32 + C::IS_OFF + 4)));
        started = true;
        loaded = false;
        dirty = false;
        k0 = k1 + 1;
      }
      if (new_block) hb_cur = hb_nxt;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------ host side, per class
inline int class_index(int tc) { return 31 - __builtin_clz((unsigned)tc); }

// per-tile arrays of a class (visit totals, stream positions)
inline int ensure_tile_arrays(StreamState* ss, long long nrows) {
  if (ss->nrows == nrows && ss->d_tot && ss->d_start) return B200_OK;
  if (ss->d_tot) cudaFree(ss->d_tot);
  if (ss->d_start) cudaFree(ss->d_start);
  ss->d_tot = nullptr;
  ss->d_start = nullptr;
  ss->nrows = 0;
  CUDA_TRY(cudaMalloc(&ss->d_tot, (size_t)(nrows + 1) * 4));
  CUDA_TRY(cudaMalloc(&ss->d_start, (size_t)(nrows + 1) * 4));
  ss->nrows = nrows;
  return B200_OK;
}

template <int DIM, int W, int TC>
int build_stream(b200_plan* p, RowsState* ts, cudaStream_t st) {
  using C = Cls<DIM, TC>;
  StreamState* ss = &ts->cls[class_index(TC)];
  const long long nrows = num_tiles<DIM, TC>(p->g);
  auto fr = [](void* q) {
    if (q) cudaFree(q);
  };
  ss->unsupported = false;
  ss->valid = false;
  if (nrows >= (1LL << 30) || p->M >= (1LL << 31) - 2 || (p->rows_dbg & 8)) {  // bit 3: test hook
    ss->unsupported = true;
    ss->valid = true;
    return B200_OK;
  }
  B200_TRY(ensure_tile_arrays(ss, nrows));
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, 64, st));
  k_row_totals<DIM, W, TC><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(
      p->g, nrows, p->d_bin_start, ss->d_tot, reinterpret_cast<unsigned long long*>(ts->d_counters + 2));
  CHECK_LAUNCH();
  unsigned long long grand = 0;
  CUDA_TRY(cudaMemcpyAsync(&grand, ts->d_counters + 2, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (grand + (unsigned long long)nrows >= (1ULL << 31)) {
    ss->unsupported = true;
    ss->valid = true;
    return B200_OK;
  }
  k_scan_inputs<<<ceil_div(nrows + 1, 256), 256, 0, st>>>(nrows + 1, ss->d_tot, ss->d_start);
  CHECK_LAUNCH();
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, ss->d_start, ss->d_start, (int)(nrows + 1), st);
  if (need > ts->scan_tmp_bytes) {
    fr(ts->d_scan_tmp);
    ts->d_scan_tmp = nullptr;
    ts->scan_tmp_bytes = 0;
    CUDA_TRY(cudaMalloc(&ts->d_scan_tmp, need));
    ts->scan_tmp_bytes = need;
  }
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(ts->d_scan_tmp, need, ss->d_start, ss->d_start, (int)(nrows + 1), st));
  g_kernel_launches += 2;
  uint32_t S = 0;
  CUDA_TRY(cudaMemcpyAsync(&S, ss->d_start + nrows, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // Chunk = unit of dynamically scheduled work.  2048 entries amortise the per-chunk set-up on long streams;
  // a short stream (2-D, few samples, few coils) is cut finer so that every resident warp still gets several
  // chunks (cfg-B: 7.7e5 entries = 376 chunks of 2048 for 2368 warp slots).
  int lch = LCH;
  while (lch > 128 && (long long)S / lch < 8LL * p->num_sms * 4 * WARPS) lch >>= 1;
  ss->lch = lch;
  const int nchunks = (int)((S + lch - 1) / lch);
  // + 64 entries of all-ones padding: the kernel looks one entry past the last chunk.
  // Buffers only ever grow: update_samples in a trajectory-learning loop must not pay cudaFree's
  // device synchronisation and a multi-GB cudaMalloc per step.
  if ((size_t)S + 64 > ss->ent_cap) {
    fr(ss->d_ent);
    ss->d_ent = nullptr;
    ss->ent_cap = 0;
    const size_t cap = (size_t)S + 64 + (size_t)S / 16;
    if (cudaMalloc(&ss->d_ent, cap * C::ESZ) != cudaSuccess) {
      cudaGetLastError();
      ss->d_ent = nullptr;
      ss->unsupported = true;
      ss->valid = true;
      return B200_OK;
    }
    ss->ent_cap = cap;
  }
  if ((size_t)nchunks + 1 > ss->chunk_cap) {
    fr(ss->d_chunk_row);
    fr(ss->d_split_rows);
    ss->d_chunk_row = nullptr;
    ss->d_split_rows = nullptr;
    ss->chunk_cap = 0;
    const size_t cap = (size_t)nchunks + 1 + (size_t)nchunks / 16;
    CUDA_TRY(cudaMalloc(&ss->d_chunk_row, cap * 4));
    CUDA_TRY(cudaMalloc(&ss->d_split_rows, cap * 4));
    ss->chunk_cap = cap;
  }
  CUDA_TRY(cudaMemsetAsync(ss->d_ent, 0xff, ((size_t)S + 64) * C::ESZ, st));
  CUDA_TRY(cudaMemsetAsync(ss->d_chunk_row, 0, (size_t)(nchunks + 1) * 4, st));
  k_build_stream<DIM, W, TC><<<ceil_div(nrows * 32, 256), 256, 0, st>>>(
      p->g, nrows, p->d_bin_start, ss->d_tot, ss->d_start, ts->d_rec, ss->d_ent, ss->d_chunk_row,
      ss->d_split_rows, ts->d_counters + 1, (uint32_t)ss->lch);
  CHECK_LAUNCH();
  int nsplit = 0;
  CUDA_TRY(cudaMemcpyAsync(&nsplit, ts->d_counters + 1, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  ss->S = S;
  ss->nchunks = nchunks;
  ss->nsplit = nsplit;
  ss->nvis = (long long)grand;
  ss->valid = true;
  return B200_OK;
}

template <int DIM, int W, bool SPREAD, bool FIXED, int TC>
int launch_rows_v(b200_plan* p, RowsState* ts, float2* fw, int T, const uint32_t* unread, cudaStream_t st) {
  using C = Cls<DIM, TC>;
  StreamState* ss = &ts->cls[class_index(TC)];
  auto kern = k_rows<DIM, W, SPREAD, FIXED, TC>;
  const size_t smem = (size_t)WARPS * C::smem_per_warp(SPREAD);
  static PerDeviceOnce once;
  static int ctas_per_sm = 1;
  if (once.first()) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, THREADS, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  if (SPREAD && ss->nsplit > 0) {
    k_zero_split_rows<DIM, TC><<<ceil_div(ss->nsplit * 32, 128), 128, 0, st>>>(p->g, T, ss->nsplit,
                                                                             ss->d_split_rows, fw);
    CHECK_LAUNCH();
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, sizeof(int), st));
  const long long want = (ss->nchunks + WARPS - 1) / WARPS;
  const long long cap = (long long)p->num_sms * ctas_per_sm;
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  // slot 4 of the plan's timing events brackets the row kernel alone (bench.py roofline)
  const bool timed = p->timing && p->ev_ok;
  if (timed) cudaEventRecord(p->ev[8], st);
  kern<<<grid, THREADS, smem, st>>>(p->g, T, ss->nchunks, ss->S, p->M, ss->d_ent, ss->d_chunk_row,
                                    ts->d_ptab, ts->d_kt, fw, ts->d_counters, p->rows_dbg,
                                    (SPREAD && p->spread_may_skip_empty) ? 1 : 0, 31 - __builtin_clz(ss->lch),
                                    unread);
  if (timed) {
    cudaEventRecord(p->ev[9], st);
    p->ev_used[4] = 1;
  }
  CHECK_LAUNCH();
  return B200_OK;
}

template <int DIM, int W, int TC>
int launch_rows(b200_plan* p, RowsState* ts, float2* fw, int T, bool spread, const uint32_t* unread,
                cudaStream_t st) {
  StreamState* ss = &ts->cls[class_index(TC)];
  if (TC == 32 && ss->lch == LCH) {
    // (only the class the kernels were tuned on keeps a compile-time chunk length)
    if (spread) return launch_rows_v<DIM, W, true, TC == 32, TC>(p, ts, fw, T, unread, st);
    return launch_rows_v<DIM, W, false, TC == 32, TC>(p, ts, fw, T, unread, st);
  }
  if (spread) return launch_rows_v<DIM, W, true, false, TC>(p, ts, fw, T, unread, st);
  return launch_rows_v<DIM, W, false, false, TC>(p, ts, fw, T, unread, st);
}

#define ROWS_DISPATCH_W(FN, DIMV, TCV, ...)                     \
  do {                                                          \
    const int w_ = p->g.w;                                      \
    if (w_ == 7) return FN<DIMV, 7, TCV>(__VA_ARGS__);          \
    if (w_ == 6) return FN<DIMV, 6, TCV>(__VA_ARGS__);          \
    if (w_ == 5) return FN<DIMV, 5, TCV>(__VA_ARGS__);          \
    return FN<DIMV, 4, TCV>(__VA_ARGS__);                       \
  } while (0)

}  // namespace rows

// Entry points of the per-class translation units (spread_rows_cls.cu, compiled once per class).
#define ROWS_DECLARE_CLASS(D, TCV)                                                                       \
  int rows_build_d##D##c##TCV(b200_plan* p, rows::RowsState* ts, cudaStream_t st);                        \
  int rows_launch_d##D##c##TCV(b200_plan* p, rows::RowsState* ts, float2* fw, int T, bool spread,         \
                               const uint32_t* unread, cudaStream_t st);
ROWS_DECLARE_CLASS(3, 16)
ROWS_DECLARE_CLASS(3, 8)
ROWS_DECLARE_CLASS(3, 4)
ROWS_DECLARE_CLASS(3, 2)
ROWS_DECLARE_CLASS(3, 1)
ROWS_DECLARE_CLASS(2, 16)
ROWS_DECLARE_CLASS(2, 8)
