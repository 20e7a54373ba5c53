/*
 * es_spread.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain C + OpenMP restatement of the spread / interpolate stage of the finufft algorithm that
 * the reference reaches through `finufft.Plan.execute_adjoint` / `Plan.execute`
 * (src/mrinufft/operators/interfaces/finufft.py:69, :76; finufft >= 2.5.0 per pyproject.toml:23,
 * not vendored, not installable offline).  Algorithm as published (Barnett, Magland, af Klinteberg,
 * SISC 2019) and summarised in the reference's docs/explanations/nufft.rst:253-309:
 *   exponential-of-semicircle kernel phi(x) = exp(beta (sqrt(1 - (2x/w)^2) - 1)) evaluated directly,
 *   points processed in bin-sorted order; type 1 spreads chunks of sorted points into private
 *   sub-grids that are then added (with periodic wrapping) into the shared fine grid; type 2 is an
 *   embarrassingly parallel gather.
 *
 * Compiled twice by oracle/Makefile: REAL=double (parity checks) and REAL=float (host-core timing
 * baseline: finufft's single-precision build computes in float).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL double
#endif

#define MAXW 16

static inline REAL es_phi(REAL x, int w, REAL beta) {
  REAL a = (REAL)1 - ((REAL)2 * x / w) * ((REAL)2 * x / w);
  if (a < 0) return 0;
  return (REAL)exp((double)(beta * ((REAL)sqrt((double)a) - (REAL)1)));
}

/* K1 spec (oracle/es_nufft.py::fold_points): IEEE double, one rounding per operation. */
void oracle_fold(const float* x, int64_t M, int64_t stride, int nf, int w, int32_t* origin,
                 float* x1, double* gout) {
  const double INV_2PI = 0.15915494309189535;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < M; ++j) {
    volatile double xd = (double)x[j * stride];
    volatile double t = xd * INV_2PI;
    volatile double tf = t + 0.5;
    t = t - floor(tf);
    volatile double th = t + 0.5;
    volatile double g = th * (double)nf;
    if (g >= (double)nf) g = g - (double)nf;
    if (!(g >= 0.0)) g = 0.0;
    volatile double gm = g - 0.5 * (double)w;
    double i1 = ceil(gm);
    volatile double d = i1 - g;
    x1[j] = (float)d;
    int o = (int)i1;
    if (o < 0) o += nf;
    if (o >= nf) o -= nf;
    origin[j] = o;
    if (gout) gout[j] = g;
  }
}

/* weights of point j along one axis: wts[i] = phi(x1 + i) */
static inline void tap_weights(REAL x1, int w, REAL beta, REAL* wts) {
  for (int i = 0; i < w; ++i) wts[i] = es_phi(x1 + (REAL)i, w, beta);
}

/*
 * Interpolation (type 2 stage): c[t][j] = sum_l fw[t][l] phi(l - g_j).
 *   org: (dim, M) footprint origins in [0, nf), x1: (dim, M) first-tap offsets (float32),
 *   perm: processing order (bin-sorted) or NULL, fw: (T, nf0, nf1, nf2) interleaved complex,
 *   c: (T, M) interleaved complex.
 */
void oracle_interp(int dim, const int32_t* nf, int w, double beta_, int64_t M, const int32_t* org,
                   const float* x1, const int32_t* perm, int T, const REAL* fw, REAL* c) {
  const REAL beta = (REAL)beta_;
  const int nf0 = nf[0], nf1 = dim > 1 ? nf[1] : 1, nf2 = dim > 2 ? nf[2] : 1;
  const int64_t nftot = (int64_t)nf0 * nf1 * nf2;
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t s = 0; s < M; ++s) {
    const int64_t j = perm ? perm[s] : s;
    REAL wa[3][MAXW];
    int ia[3][MAXW];
    int wd[3] = {1, 1, 1};
    for (int a = 0; a < 3; ++a) {
      if (a < dim) {
        tap_weights((REAL)x1[(int64_t)a * M + j], w, beta, wa[a]);
        const int n = nf[a];
        for (int i = 0; i < w; ++i) {
          int l = org[(int64_t)a * M + j] + i;
          ia[a][i] = l >= n ? l - n : l;
        }
        wd[a] = w;
      } else {
        wa[a][0] = 1;
        ia[a][0] = 0;
      }
    }
    for (int t = 0; t < T; ++t) {
      const REAL* g = fw + 2 * nftot * t;
      REAL ar = 0, ai = 0;
      if (dim == 1) {
        for (int i = 0; i < w; ++i) {
          ar += g[2 * ia[0][i]] * wa[0][i];
          ai += g[2 * ia[0][i] + 1] * wa[0][i];
        }
      } else if (dim == 2) {
        for (int i0 = 0; i0 < w; ++i0) {
          const REAL* row = g + 2 * (int64_t)ia[0][i0] * nf1;
          REAL rr = 0, ri = 0;
          for (int i1 = 0; i1 < w; ++i1) {
            rr += row[2 * ia[1][i1]] * wa[1][i1];
            ri += row[2 * ia[1][i1] + 1] * wa[1][i1];
          }
          ar += rr * wa[0][i0];
          ai += ri * wa[0][i0];
        }
      } else {
        for (int i0 = 0; i0 < w; ++i0) {
          REAL pr = 0, pi = 0;
          for (int i1 = 0; i1 < w; ++i1) {
            const REAL* row = g + 2 * ((int64_t)ia[0][i0] * nf1 + ia[1][i1]) * nf2;
            REAL rr = 0, ri = 0;
            for (int i2 = 0; i2 < w; ++i2) {
              rr += row[2 * ia[2][i2]] * wa[2][i2];
              ri += row[2 * ia[2][i2] + 1] * wa[2][i2];
            }
            pr += rr * wa[1][i1];
            pi += ri * wa[1][i1];
          }
          ar += pr * wa[0][i0];
          ai += pi * wa[0][i0];
        }
      }
      c[2 * ((int64_t)t * M + j)] = ar;
      c[2 * ((int64_t)t * M + j) + 1] = ai;
      (void)wd;
    }
  }
}

/*
 * Spreading (type 1 stage): fw[t][l] += sum_j c[t][j] phi(l - g_j); fw must be zeroed by the
 * caller.  Sorted points are cut into chunks; each chunk is spread into a private sub-grid
 * (bounding box of its footprints in unwrapped coordinates) and then added into fw with atomics.
 */
void oracle_spread(int dim, const int32_t* nf, int w, double beta_, int64_t M, const int32_t* org,
                   const float* x1, const int32_t* perm, int T, const REAL* c, REAL* fw,
                   int64_t chunk) {
  const REAL beta = (REAL)beta_;
  const int nfa[3] = {nf[0], dim > 1 ? nf[1] : 1, dim > 2 ? nf[2] : 1};
  const int64_t nftot = (int64_t)nfa[0] * nfa[1] * nfa[2];
  if (chunk <= 0) chunk = 16384;
  const int64_t nchunks = (M + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ch = 0; ch < nchunks; ++ch) {
    const int64_t s0 = ch * chunk, s1 = (s0 + chunk < M) ? s0 + chunk : M;
    /* unwrapped origin: the stored origin is i1 mod nf with i1 in [-w/2-1, nf); un-wrap values in
       the top w cells to negative so that boxes stay compact when points straddle the seam */
    int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    for (int a = 0; a < dim; ++a) {
      int mn = 1 << 30, mx = -(1 << 30);
      for (int64_t s = s0; s < s1; ++s) {
        const int64_t j = perm ? perm[s] : s;
        int o = org[(int64_t)a * M + j];
        if (o < mn) mn = o;
        if (o > mx) mx = o;
      }
      lo[a] = mn;
      hi[a] = mx + w; /* exclusive */
    }
    const int sz[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const int64_t sub = (int64_t)sz[0] * sz[1] * sz[2];
    REAL* sg = (REAL*)calloc((size_t)(2 * sub), sizeof(REAL));
    for (int t = 0; t < T; ++t) {
      if (t > 0) memset(sg, 0, (size_t)(2 * sub) * sizeof(REAL));
      for (int64_t s = s0; s < s1; ++s) {
        const int64_t j = perm ? perm[s] : s;
        REAL wa[3][MAXW];
        int o[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) {
          if (a < dim) {
            tap_weights((REAL)x1[(int64_t)a * M + j], w, beta, wa[a]);
            o[a] = org[(int64_t)a * M + j] - lo[a];
          } else {
            wa[a][0] = 1;
          }
        }
        const REAL cr = c[2 * ((int64_t)t * M + j)], ci = c[2 * ((int64_t)t * M + j) + 1];
        const int w0 = w, w1 = dim > 1 ? w : 1, w2 = dim > 2 ? w : 1;
        for (int i0 = 0; i0 < w0; ++i0) {
          for (int i1 = 0; i1 < w1; ++i1) {
            const REAL w01 = wa[0][i0] * wa[1][i1];
            REAL* row = sg + 2 * (((int64_t)(o[0] + i0) * sz[1] + (o[1] + i1)) * sz[2] + o[2]);
            const REAL vr = cr * w01, vi = ci * w01;
            for (int i2 = 0; i2 < w2; ++i2) {
              row[2 * i2] += vr * wa[2][i2];
              row[2 * i2 + 1] += vi * wa[2][i2];
            }
          }
        }
      }
      /* add the sub-grid into the shared grid with periodic wrapping */
      REAL* g = fw + 2 * nftot * t;
      for (int a0 = 0; a0 < sz[0]; ++a0) {
        int l0 = (lo[0] + a0) % nfa[0];
        for (int a1 = 0; a1 < sz[1]; ++a1) {
          int l1 = (lo[1] + a1) % nfa[1];
          const REAL* srow = sg + 2 * (((int64_t)a0 * sz[1] + a1) * sz[2]);
          REAL* grow = g + 2 * (((int64_t)l0 * nfa[1] + l1) * nfa[2]);
          for (int a2 = 0; a2 < sz[2]; ++a2) {
            int l2 = (lo[2] + a2) % nfa[2];
            const REAL vr = srow[2 * a2], vi = srow[2 * a2 + 1];
            if (vr != 0 || vi != 0) {
#pragma omp atomic
              grow[2 * l2] += vr;
#pragma omp atomic
              grow[2 * l2 + 1] += vi;
            }
          }
        }
      }
    }
    free(sg);
  }
}

int oracle_real_bytes(void) { return (int)sizeof(REAL); }
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* torchrun exports OMP_NUM_THREADS=1 to its ranks: the timing legs set the thread count explicitly */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
