// spread_rows.cu -- host side of the row kernels ("method 2" spreading / interpolation, see
// rows_common.cuh for the design), the per-trajectory pre-passes, the k-space transposes, and the
// kernels of coil class 32.  The smaller coil classes are compiled from spread_rows_cls.cu.
//
// Replaces finufft's spread/interp stage (call sites
// src/mrinufft/operators/interfaces/finufft.py:69,76; algorithm docs/explanations/nufft.rst:253-309).
#include "rows_common.cuh"

using namespace rows;

namespace {
// Per-visit tap kernels of coil class 32: generated inline PTX (tools/gen_taps.py) -- one `brx.idx` on the
// x offset, then w packed `fma.rn.f32x2` (SASS FFMA2) on statically indexed 64-bit (re, im) accumulators.
// A C++ `switch` is lowered by nvcc to a compare/branch tree that cost 17 issue slots per visit.
#include "taps_generated.inc"
}  // namespace

namespace rows {
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_spread(u64 (&acc)[NACC], unsigned pk, int n, unsigned vb, unsigned,
                                                 unsigned) {
  static_assert(TC == 32, "this translation unit holds coil class 32");
  if (W == 7) rows_loop_spread_w7(acc, pk, n, vb);
  else if (W == 6) rows_loop_spread_w6(acc, pk, n, vb);
  else if (W == 5) rows_loop_spread_w5(acc, pk, n, vb);
  else rows_loop_spread_w4(acc, pk, n, vb);
}
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_interp(u64 (&acc)[NACC], unsigned pk, int n, const void* ktl, unsigned,
                                                 unsigned, unsigned) {
  static_assert(TC == 32, "this translation unit holds coil class 32");
  if (W == 7) rows_loop_interp_w7(acc, pk, n, ktl);
  else if (W == 6) rows_loop_interp_w6(acc, pk, n, ktl);
  else if (W == 5) rows_loop_interp_w5(acc, pk, n, ktl);
  else rows_loop_interp_w4(acc, pk, n, ktl);
}
}  // namespace rows

namespace {

RowsState* state(b200_plan* p) {
  if (!p->tiled) p->tiled = new RowsState();
  return (RowsState*)p->tiled;
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

// One flag bit per class-32 tile (2 rows x 16 cells), 1 = no visitors, laid out as a bit string per grid
// column of tiles along the slowest axis: 3-D word [(y / 2) nbx + bx][z / 32], bit z % 32; 2-D word
// [bx][(y / 2) / 32], bit (y / 2) % 32.  Lets the spreader skip the zero-fill of such tiles when the next
// consumer is the fused FFT, whose first pass (along that axis) then substitutes zeros instead of reading
// them (half the grid for a radial trajectory); a CTA of that pass needs one contiguous bit string.  The
// larger tiles of the smaller coil classes are unions of these: an empty one has all its bits set.
template <int DIM>
__global__ void __launch_bounds__(256)
k_mark_empty(Geom g, long long nrows, const int32_t* __restrict__ tot, uint32_t* __restrict__ bits,
             int words_per_col) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  TileCoord rc;
  if (!decode_tile<DIM, 32>(g, row, &rc) || tot[row] != 0) return;
  const int nbx = num_xtiles<DIM>(g);
  const long long col = DIM == 3 ? (long long)(rc.y >> 1) * nbx + rc.bx : rc.bx;
  const int pos = DIM == 3 ? rc.z : (rc.y >> 1);
  atomicOr(bits + col * words_per_col + (pos >> 5), 1u << (pos & 31));
}

// Transposes between the caller's k-space batch ksp[t][j] and the row kernels' kt[s][t], t < TC
// (s = sorted position of sample j, TC = coil class of the call).  A warp moves 32 consecutive samples
// x TC coils: coalesced 256-byte rows per coil on the ksp side, runs of TC coils per sample on the kt side.
constexpr int KT_WARPS = 4;
constexpr int KT_STRIDE = 33;  // u64 row stride of the transpose tile

// kt[iperm[j]][t] = ksp[t][j] * density[j]   (t < T; coils T <= t < TC are zero-filled)
__global__ void __launch_bounds__(KT_WARPS * 32)
k_gather_kspace(long long M, int T, int TC, const int32_t* __restrict__ iperm,
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
  for (int t = 0; t < TC; ++t) {
    float2 v = make_float2(0.f, 0.f);
    if (t < T && live) {
      const u64 raw = __ldg(src + (long long)t * M + j);
      v = make_float2(__uint_as_float((unsigned)raw) * d, __uint_as_float((unsigned)(raw >> 32)) * d);
    }
    tl[lane * KT_STRIDE + t] = ((u64)__float_as_uint(v.y) << 32) | (u64)__float_as_uint(v.x);
  }
  __syncwarp();
  u64* dst = reinterpret_cast<u64*>(kt);
  const int lg = 31 - __clz(TC);
#pragma unroll 8
  for (int it = 0; it < TC; ++it) {
    const int e = it * 32 + lane, i = e >> lg, t = e & (TC - 1);
    const int si = __shfl_sync(0xffffffffu, s, i);
    if (si >= 0) dst[(long long)si * TC + t] = tl[i * KT_STRIDE + t];
  }
}

// ksp[t][j] = scale * kt[iperm[j]][t] (- obs[t][j])
__global__ void __launch_bounds__(KT_WARPS * 32)
k_scatter_kspace(long long M, int T, int TC, const int32_t* __restrict__ iperm,
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
  const int lg = 31 - __clz(TC);
#pragma unroll 8
  for (int it = 0; it < TC; ++it) {
    const int e = it * 32 + lane, i = e >> lg, t = e & (TC - 1);
    const int si = __shfl_sync(0xffffffffu, s, i);
    tl[i * KT_STRIDE + t] = si >= 0 ? __ldg(src + (long long)si * TC + t) : 0ull;
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

constexpr int EFALLBACK = 1;  // internal: the row kernels cannot serve this trajectory

// per-trajectory point data shared by every coil class: weight records, x-weight table, inverse permutation
template <int W>
int prepare_points(b200_plan* p, RowsState* ts, cudaStream_t st) {
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
    // the records (112 B per point) only feed the stream builders, but stay allocated: a
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
  for (int c = 0; c < NCLS; ++c) ts->cls[c].valid = false;
  ts->empty_valid = false;
  ts->M = M;
  ts->valid = true;
  return B200_OK;
}

int ensure_state(b200_plan* p, cudaStream_t st) {
  RowsState* ts = state(p);
  if (ts->valid && ts->M == p->M) return B200_OK;
  switch (p->g.w) {
    case 7: return prepare_points<7>(p, ts, st);
    case 6: return prepare_points<6>(p, ts, st);
    case 5: return prepare_points<5>(p, ts, st);
    default: return prepare_points<4>(p, ts, st);
  }
}

// class-32 visit totals -> empty-tile bit strings (also needed when only a smaller class has a stream)
template <int DIM, int W, int TC>
int mark_empty(b200_plan* p, RowsState* ts, cudaStream_t st) {
  static_assert(TC == 32, "the bit strings describe the class-32 tiling");
  StreamState* s32 = &ts->cls[class_index(32)];
  const long long nrows = num_tiles<DIM, 32>(p->g);
  if (nrows >= (1LL << 30)) return EFALLBACK;
  if (!s32->valid || s32->unsupported || s32->nrows != nrows) {
    B200_TRY(ensure_tile_arrays(s32, nrows));
    s32->valid = false;
    CUDA_TRY(cudaMemsetAsync(ts->d_counters, 0, 64, st));
    k_row_totals<DIM, W, 32><<<ceil_div(nrows + 1, 256), 256, 0, st>>>(
        p->g, nrows, p->d_bin_start, s32->d_tot, reinterpret_cast<unsigned long long*>(ts->d_counters + 2));
    CHECK_LAUNCH();
  }
  const int nyh = p->g.nf[DIM - 2] / 2, nbx = num_xtiles<DIM>(p->g);
  const int npos = DIM == 3 ? p->g.nf[0] : nyh;          // tiles per column
  const int wpc = (npos + 31) / 32;
  const size_t nwords = (size_t)(DIM == 3 ? (size_t)nyh * nbx : nbx) * wpc;
  if (nwords > ts->empty_cap) {
    if (ts->d_empty) cudaFree(ts->d_empty);
    ts->d_empty = nullptr;
    ts->empty_cap = 0;
    CUDA_TRY(cudaMalloc(&ts->d_empty, nwords * 4));
    ts->empty_cap = nwords;
  }
  CUDA_TRY(cudaMemsetAsync(ts->d_empty, 0, nwords * 4, st));
  k_mark_empty<DIM><<<ceil_div(nrows, 256), 256, 0, st>>>(p->g, nrows, s32->d_tot, ts->d_empty, wpc);
  CHECK_LAUNCH();
  ts->empty_valid = true;
  return B200_OK;
}

int ensure_empty_bits(b200_plan* p, RowsState* ts, cudaStream_t st) {
  if (ts->empty_valid) return B200_OK;
  if (p->g.dim == 3) ROWS_DISPATCH_W(mark_empty, 3, 32, p, ts, st);
  ROWS_DISPATCH_W(mark_empty, 2, 32, p, ts, st);
}

// geometry of a coil class as run-time numbers
struct ClassDims {
  int ty, gz;
};
ClassDims class_dims(int dim, int tc) {
  switch (dim * 100 + tc) {
    case 316: return {Cls<3, 16>::TY, Cls<3, 16>::GZ};
    case 308: return {Cls<3, 8>::TY, Cls<3, 8>::GZ};
    case 304: return {Cls<3, 4>::TY, Cls<3, 4>::GZ};
    case 302: return {Cls<3, 2>::TY, Cls<3, 2>::GZ};
    case 301: return {Cls<3, 1>::TY, Cls<3, 1>::GZ};
    case 216: return {Cls<2, 16>::TY, 1};
    case 208: return {Cls<2, 8>::TY, 1};
    default: return {2, 1};
  }
}

// Coil class of a call with T coils: the smallest class that holds them (2-D: 8 at least -- its tiles grow
// along y only) and whose tile fits the grid's periodic wrap; option 4 of b200_plan_set_option forces a
// larger class (tests, timing experiments).
int pick_class(const b200_plan* p, int T) {
  const Geom& g = p->g;
  const int lowest = g.dim == 3 ? 1 : 8;
  for (int tc = lowest; tc < 32; tc <<= 1) {
    if (tc < T || tc < p->rows_class) continue;
    const ClassDims cd = class_dims(g.dim, tc);
    if (g.nf[g.dim - 2] < g.w + cd.ty - 1) continue;
    if (g.dim == 3 && g.nf[0] < g.w + cd.gz - 1) continue;
    return tc;
  }
  return 32;
}

int build_class(b200_plan* p, RowsState* ts, int tc, cudaStream_t st) {
  if (p->g.dim == 3) {
    switch (tc) {
      case 16: return rows_build_d3c16(p, ts, st);
      case 8: return rows_build_d3c8(p, ts, st);
      case 4: return rows_build_d3c4(p, ts, st);
      case 2: return rows_build_d3c2(p, ts, st);
      case 1: return rows_build_d3c1(p, ts, st);
      default: ROWS_DISPATCH_W(build_stream, 3, 32, p, ts, st);
    }
  }
  switch (tc) {
    case 16: return rows_build_d2c16(p, ts, st);
    case 8: return rows_build_d2c8(p, ts, st);
    default: ROWS_DISPATCH_W(build_stream, 2, 32, p, ts, st);
  }
}

int launch_class(b200_plan* p, RowsState* ts, int tc, float2* fw, int T, bool spread, const uint32_t* unread,
                 cudaStream_t st) {
  if (p->g.dim == 3) {
    switch (tc) {
      case 16: return rows_launch_d3c16(p, ts, fw, T, spread, unread, st);
      case 8: return rows_launch_d3c8(p, ts, fw, T, spread, unread, st);
      case 4: return rows_launch_d3c4(p, ts, fw, T, spread, unread, st);
      case 2: return rows_launch_d3c2(p, ts, fw, T, spread, unread, st);
      case 1: return rows_launch_d3c1(p, ts, fw, T, spread, unread, st);
      default: ROWS_DISPATCH_W(launch_rows, 3, 32, p, ts, fw, T, spread, unread, st);
    }
  }
  switch (tc) {
    case 16: return rows_launch_d2c16(p, ts, fw, T, spread, unread, st);
    case 8: return rows_launch_d2c8(p, ts, fw, T, spread, unread, st);
    default: ROWS_DISPATCH_W(launch_rows, 2, 32, p, ts, fw, T, spread, unread, st);
  }
}

// stream of the class that serves T coils, built on first use; *tc_out = the class, EFALLBACK if the row
// kernels cannot serve this trajectory
int ensure_stream(b200_plan* p, int T, int* tc_out, cudaStream_t st) {
  B200_TRY(ensure_state(p, st));
  RowsState* ts = state(p);
  const int tc = pick_class(p, T);
  StreamState* ss = &ts->cls[class_index(tc)];
  if (!ss->valid) B200_TRY(build_class(p, ts, tc, st));
  *tc_out = tc;
  return ss->unsupported ? EFALLBACK : B200_OK;
}

int ensure_kt(RowsState* ts, long long M, int tc) {
  const size_t need = (size_t)(M > 0 ? M : 1) * tc * sizeof(float2);
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
  if (p->tiled) {
    const RowsState* ts = (const RowsState*)p->tiled;
    const StreamState* ss = &ts->cls[class_index(pick_class(p, T))];
    if (ts->valid && ts->M == p->M && ss->valid && ss->unsupported) return false;
  }
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
  fr(ts->d_ptab);
  fr(ts->d_empty);
  for (int c = 0; c < NCLS; ++c) {
    fr(ts->cls[c].d_tot);
    fr(ts->cls[c].d_start);
    fr(ts->cls[c].d_ent);
    fr(ts->cls[c].d_chunk_row);
    fr(ts->cls[c].d_split_rows);
  }
  fr(ts->d_counters);
  fr(ts->d_scan_tmp);
  delete ts;
  p->tiled = nullptr;
}

// Coil class that a call with T coils runs in, and the number of (point, tile) visits of its stream
// (0 before the stream exists): diagnostics for tests and bench.py.
void tiled_class_info(b200_plan* p, int T, int64_t out[4]) {
  out[0] = pick_class(p, T);
  out[1] = out[2] = out[3] = 0;
  if (!p->tiled) return;
  const RowsState* ts = (const RowsState*)p->tiled;
  const StreamState* ss = &ts->cls[class_index((int)out[0])];
  if (!ts->valid || !ss->valid) return;
  out[1] = ss->nvis;
  out[2] = ss->S;
  out[3] = ss->unsupported ? 1 : 0;
}

// Bit strings of the class-32 tiles that no point visits (nullptr if the row kernels do not serve a
// T-coil transform of this trajectory): the type-2 FFT may leave those tiles unwritten, the row
// interpolator never reads them (its larger-tile classes substitute zeros).
const uint32_t* tiled_empty_bits(b200_plan* p, int T, cudaStream_t st) {
  int tc = 32;
  if (ensure_stream(p, T, &tc, st) != B200_OK) return nullptr;
  RowsState* ts = state(p);
  if (ensure_empty_bits(p, ts, st) != B200_OK || !ts->d_empty) return nullptr;
  const int d = p->g.dim;
  p->empty_nyh = p->g.nf[d - 2] / 2;
  p->empty_nbx = (p->g.nf[d - 1] + CX - 1) / CX;
  return ts->d_empty;
}

// Both entry points return 1 (not an error) when the row kernels cannot serve this trajectory
// (more than 2^31 visits or 2^27 points): the caller then uses the point-driven kernels.
int spread_tiled(b200_plan* p, const float2* ksp, const float* density, float2* fw, int T,
                 cudaStream_t st) {
  int tc = 32;
  const int rc = ensure_stream(p, T, &tc, st);
  if (rc != B200_OK) return rc;
  RowsState* ts = state(p);
  const long long M = p->M;
  // tiles without visitors are left unwritten only if the consumer can be told which ones they are
  if (p->spread_may_skip_empty && ensure_empty_bits(p, ts, st) != B200_OK) p->spread_may_skip_empty = false;
  B200_TRY(ensure_kt(ts, M, tc));
  if (M > 0) {
    k_gather_kspace<<<ceil_div(M, 32 * KT_WARPS), KT_WARPS * 32, 0, st>>>(M, T, tc, ts->d_iperm, ksp, density,
                                                                          ts->d_kt);
    CHECK_LAUNCH();
  }
  B200_TRY(launch_class(p, ts, tc, fw, T, true, nullptr, st));
  p->spread_empty = p->spread_may_skip_empty ? ts->d_empty : nullptr;
  p->empty_nyh = p->g.nf[p->g.dim - 2] / 2;
  p->empty_nbx = (p->g.nf[p->g.dim - 1] + CX - 1) / CX;
  return B200_OK;
}

// `unread`: the producer of `fw` left the class-32 tiles flagged in these bit strings unwritten
int interp_tiled(b200_plan* p, const float2* fw, float2* ksp, int T, float scale,
                 const float2* obs, const uint32_t* unread, cudaStream_t st) {
  int tc = 32;
  const int rc = ensure_stream(p, T, &tc, st);
  if (rc != B200_OK) return rc;
  RowsState* ts = state(p);
  const long long M = p->M;
  if (M == 0) return B200_OK;
  B200_TRY(ensure_kt(ts, M, tc));
  CUDA_TRY(cudaMemsetAsync(ts->d_kt, 0, (size_t)M * tc * sizeof(float2), st));
  B200_TRY(launch_class(p, ts, tc, const_cast<float2*>(fw), T, false, unread, st));
  k_scatter_kspace<<<ceil_div(M, 32 * KT_WARPS), KT_WARPS * 32, 0, st>>>(M, T, tc, ts->d_iperm, ts->d_kt, ksp,
                                                                         scale, obs);
  CHECK_LAUNCH();
  return B200_OK;
}
