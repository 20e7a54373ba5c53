// spread_rows_cls.cu -- the row kernels of ONE coil class below 32 (see rows_common.cuh), compiled once per
// class with -DROWS_DIM=<2|3> -DROWS_TC=<16|8|4|2|1>: each class has its own generated visit loops
// (tools/gen_taps.py -> taps_generated_d<dim>c<tc>.inc) and its own translation unit so that `make -j`
// builds them side by side.
#include "rows_common.cuh"

#ifndef ROWS_DIM
#error "compile with -DROWS_DIM=... -DROWS_TC=..."
#endif

#define ROWS_CAT_(a, b, c, d) a##b##c##d
#define ROWS_CAT(a, b, c, d) ROWS_CAT_(a, b, c, d)
#define ROWS_STR_(x) #x
#define ROWS_STR(x) ROWS_STR_(x)
// rows_loop_spread_w7_d3c16 etc.
#define ROWS_LOOP(kind, w) ROWS_CAT(rows_loop_##kind##_w##w##_d, ROWS_DIM, c, ROWS_TC)

using namespace rows;

namespace {
#include ROWS_STR(ROWS_CAT(taps_generated_d, ROWS_DIM, c, ROWS_TC).inc)
}  // namespace

namespace rows {
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_spread(u64 (&acc)[NACC], unsigned pk, int n, unsigned vb, unsigned yo,
                                                 unsigned zo) {
  static_assert(DIM == ROWS_DIM && TC == ROWS_TC, "one coil class per translation unit");
  if (W == 7) ROWS_LOOP(spread, 7)(acc, pk, n, vb, yo, zo);
  else if (W == 6) ROWS_LOOP(spread, 6)(acc, pk, n, vb, yo, zo);
  else if (W == 5) ROWS_LOOP(spread, 5)(acc, pk, n, vb, yo, zo);
  else ROWS_LOOP(spread, 4)(acc, pk, n, vb, yo, zo);
}
template <int W, int DIM, int TC>
__device__ __forceinline__ void rows_loop_interp(u64 (&acc)[NACC], unsigned pk, int n, const void* ktl, unsigned ob,
                                                 unsigned yo, unsigned zo) {
  static_assert(DIM == ROWS_DIM && TC == ROWS_TC, "one coil class per translation unit");
  // class 16 adds to k-space through `ktl` like class 32 (Cls::DIRECT), the others park partial sums at `ob`
#if ROWS_TC == 16
#define ROWS_SINK ktl
#else
#define ROWS_SINK ob
#endif
  if (W == 7) ROWS_LOOP(interp, 7)(acc, pk, n, ROWS_SINK, yo, zo);
  else if (W == 6) ROWS_LOOP(interp, 6)(acc, pk, n, ROWS_SINK, yo, zo);
  else if (W == 5) ROWS_LOOP(interp, 5)(acc, pk, n, ROWS_SINK, yo, zo);
  else ROWS_LOOP(interp, 4)(acc, pk, n, ROWS_SINK, yo, zo);
  (void)ktl;
  (void)ob;
}
}  // namespace rows

int ROWS_CAT(rows_build_d, ROWS_DIM, c, ROWS_TC)(b200_plan* p, RowsState* ts, cudaStream_t st) {
  ROWS_DISPATCH_W(build_stream, ROWS_DIM, ROWS_TC, p, ts, st);
}

int ROWS_CAT(rows_launch_d, ROWS_DIM, c, ROWS_TC)(b200_plan* p, RowsState* ts, float2* fw, int T, bool spread,
                                                 const uint32_t* unread, cudaStream_t st) {
  ROWS_DISPATCH_W(launch_rows, ROWS_DIM, ROWS_TC, p, ts, fw, T, spread, unread, st);
}
