/*
 * b200nufft.h -- C ABI of libb200nufft.so, the sm_100a NUFFT engine behind
 * mri-nufft's `get_operator("b200")` backend.
 *
 * This is the drop-in boundary.  Every entry point names the
 * reference interface it replaces (paths relative to the mri-nufft tree).
 * In the reference the arithmetic of this path is delegated to the un-vendored
 * third-party library finufft (pyproject.toml:23) through `finufft.Plan`; the
 * functions below are what a binding for that path has to call instead.
 *
 * Conventions
 *   - plain C, opaque plan handle, caller owns every data buffer, the plan owns
 *     its workspace (oversampled grids, sorted points, cuFFT plan);
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default
 *     stream); every call is asynchronous on that stream;
 *   - complex data is interleaved float (re, im) = complex64, C-order, the last
 *     image axis is the fastest one; `samples[:, i]` pairs with image axis `i`
 *     (src/mrinufft/operators/interfaces/finufft.py:55-62);
 *   - return value 0 = success, negative = error (see B200_E*), message through
 *     b200_last_error() (thread local);
 *   - a plan is bound to one device and is not thread-safe.
 *
 * Math (docs/explanations/nufft.rst:253-309, verified against
 * src/mrinufft/operators/interfaces/nudft_numpy.py:12-55):
 *   type 2:  c_j = scale * sum_n  f_n * exp(isign * i * x_j . (n - N/2))   (isign = -1)
 *   type 1:  f_n = scale * sum_j  c_j * exp(isign * i * x_j . (n - N/2))   (isign = +1)
 *   x_j in radians, folded periodically into [-pi, pi).
 */
#ifndef B200NUFFT_H
#define B200NUFFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_plan b200_plan;

#define B200_OK          0
#define B200_EINVAL     -1   /* bad argument                                  */
#define B200_ECUDA      -2   /* CUDA runtime error                            */
#define B200_ECUFFT     -3   /* cuFFT error                                   */
#define B200_ENOMEM     -4   /* allocation failed                             */
#define B200_ESTATE     -5   /* call order (e.g. execute before setpts)       */

/* plan flags */
#define B200_SPREAD_ONLY 1   /* finufft `spreadinterponly=1` (finufft.py:225-232) */
#define B200_DOUBLE      2   /* float64 / complex128 plan: `dtype = samples.dtype` (base.py:934), finufft's
                                double instantiation.  Every data pointer of the plan's calls is then
                                complex128 / float64 (sample coordinates and density included).
                                Correctness-first kernels; no spread-only mode, sort read-back or Toeplitz. */

#define B200_EXACT_GRID  4   /* keep the oversampled grid at next235even(sigma N) per axis.  Without it the fastest
                                axis steps past sizes whose last 16-cell tile is narrower than w - 1 cells (the
                                tiled kernels do not take those: 450 -> 480), and a 3-D single-precision plan
                                whose grid has factors 3 / 5 takes the next power of two when that is at most a
                                third larger on every axis (the fused FFT passes need powers of two; the kernel is
                                only more accurate on a finer grid).  b200_toeplitz_apply needs exactly 2 N. */

/* Version of this ABI (bumped on any signature change). */
int b200_abi_version(void);

/* Thread-local message of the last failing call on this thread. */
const char* b200_last_error(void);

/*
 * Create a plan: replaces `finufft.Plan(2, shape, n_trans, eps, dtype=..., **kw)`
 * (src/mrinufft/operators/interfaces/finufft.py:43-50) and the second type-1
 * plan of the cufinufft interface (cufinufft.py:83-92).  One plan serves both
 * transform types and both signs.
 *   dim          2 or 3 (1 is accepted)
 *   n_modes      image shape, n_modes[i] pairs with samples[:, i]
 *   n_trans_max  largest number of transforms batched in one execute call
 *   eps          requested tolerance (kernel width w = ceil(log10(10/eps)) at sigma = 2)
 *   upsampfac    sigma; 0 -> 2.0
 *   flags        B200_SPREAD_ONLY: no FFT, no deapodisation, grid size = n_modes; B200_DOUBLE, B200_EXACT_GRID: above
 *   device       CUDA device ordinal
 */
int b200_plan_create(b200_plan** plan, int dim, const int64_t* n_modes,
                     int n_trans_max, double eps, double upsampfac,
                     int flags, int device);

int b200_plan_destroy(b200_plan* plan);

/*
 * Plan geometry, for the host side and the bit-exact sort test.
 * info[0..2]  = fine grid size per axis (nf)
 * info[3]     = kernel width w
 * info[4..6]  = bin size per axis (cells)
 * info[7..9]  = number of bins per axis
 * info[10]    = polynomial degree of the kernel evaluator
 * info[11]    = number of points M (0 before setpts)
 * info[12]    = workspace bytes held by the plan
 */
int b200_plan_info(const b200_plan* plan, int64_t info[16]);

/* Kernel shape parameters: out[0] = beta, out[1] = c (= 4 / w^2), out[2] = sigma. */
int b200_plan_kernel_params(const b200_plan* plan, double out[4]);

/*
 * Set the non-uniform points: replaces `Plan.setpts(x, y, z)`
 * (finufft.py:55-62, called from MRIfinufft.__init__ L130-135 and
 * update_samples L150-181).  This is kernel K1: fold to fine-grid units,
 * footprint origin + in-cell offset, bin key, stable sort by bin key.
 *   xyz   device pointer, float32, shape (M, dim), C-order, radians
 */
int b200_plan_setpts(b200_plan* plan, int64_t M, const float* xyz, void* stream);

/*
 * Read back the sort for the bit-exact test (no reference counterpart: finufft
 * keeps its sort private).  Any pointer may be NULL.
 *   origin  int32 (dim, M)  footprint origin cell per axis, UNSORTED point order
 *   x1      float32 (dim, M) offset of the first tap, UNSORTED point order
 *   key     int32 (M)       bin key, UNSORTED point order
 *   perm    int32 (M)       perm[s] = index of the point at sorted position s
 */
int b200_plan_get_sort(b200_plan* plan, int32_t* origin, float* x1,
                       int32_t* key, int32_t* perm, void* stream);

/*
 * Type 2 (uniform -> non-uniform) for T <= n_trans_max transforms: replaces
 * `RawFinufftPlan.op` / `Plan.execute` (finufft.py:71-76) together with the
 * per-chunk sensitivity-map multiply and the `ret *= inv_norm_factor` pass of
 * `FourierOperatorSimple._op_sense/_op_calibless/op`
 * (src/mrinufft/operators/base.py:949-1013).
 *   img    complex64; (T, *n_modes) if smaps == NULL, else (*n_modes) (one image)
 *   smaps  NULL or complex64 (T, *n_modes); coil images img*smaps[t] are formed on the fly
 *   ksp    complex64 (T, M) output, natural (unsorted) point order
 *   isign  -1 (forward model) or +1 (the `grad` plan of finufft.py:183-193)
 *   scale  multiplied into the output (1/norm_factor)
 *   conj_smaps  use conj(smaps) (toggle_grad_traj, base.py:1234-1238)
 */
int b200_type2(b200_plan* plan, const void* img, const void* smaps, void* ksp,
               int T, int isign, float scale, int conj_smaps, void* stream);

/*
 * Type 1 (non-uniform -> uniform) for T transforms: replaces
 * `RawFinufftPlan.adj_op` / `Plan.execute_adjoint` (finufft.py:64-69), the density
 * multiply of `_adj_op` (base.py:1068-1073), the conj-smaps multiply + coil
 * accumulation of `_adj_op_sense` (base.py:1037-1052; CUDA twin
 * `_coil_combine_kernel`, src/mrinufft/operators/gpu_utils.py:12-24) and the final
 * `ret *= inv_norm_factor`.
 *   ksp      complex64 (T, M)
 *   density  NULL or float32 (M): multiplies ksp before spreading
 *   smaps    NULL or complex64 (T, *n_modes): img = sum_t conj(smaps[t]) * f_t
 *   img      complex64 (*n_modes) if smaps else (T, *n_modes)
 *   accumulate  0: overwrite img, 1: img += result (next coil chunk)
 *   conj_smaps  0: multiply by conj(smaps) (adjoint), 1: by smaps (toggled plan)
 */
int b200_type1(b200_plan* plan, const void* ksp, const float* density,
               const void* smaps, void* img, int T, int accumulate, int isign,
               float scale, int conj_smaps, void* stream);

/*
 * Fused data-consistency gradient chunk  A^H (A x - y): replaces one iteration of the
 * loop in `_grad_sense/_grad_calibless` (base.py:1092-1139; GPU twin
 * cufinufft.py:992-1178).  The k-space residual never leaves the device and is never
 * materialised in caller memory.
 *   img   as in b200_type2, obs complex64 (T, M), grad as `img` of b200_type1
 *   The residual is r = scale * A x - obs, then (density .* r), then grad (+)= scale * A^H r.
 */
int b200_data_consistency(b200_plan* plan, const void* img, const void* smaps,
                          const void* obs, const float* density, void* grad,
                          int T, int accumulate, float scale, void* stream);

/*
 * Toeplitz (Gram) operator chunk  x -> sum_t conj(S_t) . crop(IFFT(K . FFT(pad(S_t . x)))): replaces
 * `apply_toeplitz_kernel` (src/mrinufft/operators/toeplitz.py:203-269) together with the SENSE wrap
 * of `_gram_op_sense` / `_GramOpGpuMixin` (base.py:344-356, toeplitz.py:374-430).  No trajectory is
 * needed (the kernel `kern` carries it); the plan's oversampled grid must be exactly 2 N per axis.
 *   img    as in b200_type2;  smaps NULL or complex64 (T, *n_modes)
 *   kern   float32 (*2 n_modes): real spectrum of the Toeplitz embedding, FFT index order
 *          (what `compute_toeplitz_kernel`, toeplitz.py:35-95, returns)
 *   out    as `img` of b200_type1;  scale is multiplied into the result (1 / prod(2 N) for the
 *          unnormalised FFT pair)
 */
int b200_toeplitz_apply(b200_plan* plan, const void* img, const void* smaps, const float* kern,
                        void* out, int T, int accumulate, float scale, void* stream);

/*
 * Spread / interpolate only (plans created with B200_SPREAD_ONLY): replaces the
 * `spreadinterponly=1` plan used by `MRIfinufft.pipe` (finufft.py:225-239).
 *   grid  complex64 (T, *n_modes)
 */
int b200_spread(b200_plan* plan, const void* ksp, void* grid, int T, void* stream);
int b200_interp(b200_plan* plan, const void* grid, void* ksp, int T, void* stream);

/*
 * One fixed-point iteration of Pipe's density compensation,  d <- d / |G G^H d|
 * (finufft.py:233-239), device resident.  `d` float32 (M), updated in place.
 */
int b200_pipe_iteration(b200_plan* plan, float* d, void* stream);

/*
 * z transform of the stacked (2.5-D) operator, fused with the sensitivity-map multiply, both centring
 * shifts, the kz-plane selection and the (coil, stack)-major plane layout of the 2-D operator's coil axis:
 * replaces `MRIStackedNUFFT._fftz` / `_ifftz` and the array shuffling around them
 * (src/mrinufft/operators/stacked.py:178-195, 203-240, 254-300).  No plan is needed; any length Z.
 *   forward: planes[c NZ + j, x, y] = scale * fftshift(fft(ifftshift(img[c | 0, x, y, :] * smaps[c, x, y, :])))[zsel[j]]
 *   adjoint: out[c | 0, x, y, :]    = (sum over c) conj(smaps[c]) * scale * fftshift(ifft_unnormalised(ifftshift(K_c)))
 *            with K_c[x, y, zsel[j]] = planes[c NZ + j, x, y], zero elsewhere (zsel without repeats)
 *   img / out  complex64 (C, X, Y, Z), or (X, Y, Z) when smaps is given;  smaps NULL or complex64 (C, X, Y, Z)
 *   planes     complex64 (C * NZ, X, Y);  zsel int32 (NZ) on the device, values in [0, Z)
 *   scale      multiplied into the result (the reference's 1 / sqrt(2 Z))
 */
int b200_stack_fftz_forward(const void* img, const void* smaps, void* planes, const int32_t* zsel,
                            int C, int X, int Y, int Z, int NZ, float scale, void* stream);
int b200_stack_fftz_adjoint(const void* planes, const void* smaps, void* out, const int32_t* zsel,
                            int C, int X, int Y, int Z, int NZ, float scale, void* stream);

/*
 * Temporal weights of the off-resonance-corrected operator whose L interpolators ride the coil batch as virtual
 * coils l C + c: replaces the per-interpolator `B[l] * ...` accumulation of `MRIFourierCorrected.op` / `adj_op`
 * (src/mrinufft/operators/off_resonance.py:232-332).  complex64; K = samples per coil, sample k belongs to
 * readout position k % NK.
 *   expand = 0:  y[b, c, k]     = sum_l kv[b, l, c, k] * bw[k % NK, l]
 *   expand = 1:  kv[b, l, c, k] = conj(bw[k % NK, l]) * y[b, c, k]
 *   kv (B, L * C, K), bw (NK, L), y (B, C, K)
 */
int b200_orc_weights(void* kv, const void* bw, void* y, int B, int L, int C, int64_t K, int NK, int expand,
                     void* stream);

/*
 * In-place FFT along every axis of T contiguous C-order arrays (dim = 1..3 axes of ANY length, complex64 or
 * complex128; unnormalised, sign < 0: exp(-i ...)): the library's own any-length passes (shared-memory Stockham,
 * csrc/fft_any.cu).  No plan.  Replaces the `fftn` of the Toeplitz kernel assembly
 * (src/mrinufft/operators/toeplitz.py:89-93); option key 2 = 4 routes the plans' non-power-of-two and
 * complex128 grids through the same passes.
 */
int b200_fft_c2c(void* data, int T, int dim, const int64_t* n, int sign, int dbl, void* stream);

/*
 * Vector updates of the iterative solvers, one pass over memory each: replace the array expressions of
 * `lsqr` / `lsmr` / `cg` (src/mrinufft/extras/optim.py:402-446, 669-724, 866-883).  Vectors are device arrays
 * of B x n complex64 (dbl = 0) or complex128 (dbl = 1) elements, batch-major; scalars are HOST arrays of B
 * complex doubles (re, im interleaved; the imaginary part is ignored where the update is real); sums are
 * DEVICE arrays of doubles, overwritten.  `out` may alias an input.
 *   axpby      out = a x + b y  (y NULL: out = a x);  sumsq (NULL or [B]) = ||out_b||^2
 *   cg_dots    out5 [B][5] = { ||gnew||^2, Re, Im sum gnew (gnew - gold), Re, Im sum gold gold }  (no conjugates,
 *              like `xp.dot`, optim.py:872-878)
 *   cg_step    v = g + beta v ;  x = x + minus_inv_l v
 *   lsqr_step  sumsq_w [B] = ||w_b||^2 of the incoming w ;  x += t1 w ;  w = v + t2 w
 *   lsmr_step  hbar = h + a hbar ;  x += b hbar ;  h = v + c h ;  sumsq_x [B] = ||x_b||^2 of the new x
 */
int b200_vec_axpby(void* out, const void* x, const void* y, const double* a, const double* b, int B,
                   int64_t n, double* sumsq, int dbl, void* stream);
int b200_vec_cg_dots(const void* gnew, const void* gold, int B, int64_t n, double* out5, int dbl,
                     void* stream);
int b200_vec_cg_step(void* x, void* v, const void* g, const double* beta, const double* minus_inv_l,
                     int B, int64_t n, int dbl, void* stream);
int b200_vec_lsqr_step(void* x, void* w, const void* v, const double* t1, const double* t2, int B,
                       int64_t n, double* sumsq_w, int dbl, void* stream);
int b200_vec_lsmr_step(void* x, void* hbar, void* h, const void* v, const double* a, const double* b,
                       const double* c, int B, int64_t n, double* sumsq_x, int dbl, void* stream);

/*
 * Counters for bench.py: number of kernels this library launched and number of cuFFT
 * executions since the last reset (process wide).
 */
int b200_launch_count(int64_t* kernels, int64_t* ffts, int reset);

/*
 * Select kernel variants at run time (A/B measurements; 0 = default for every key).
 *   key 0: spread method   (0 auto, 1 global-atomic point driven, 2 tiled)
 *   key 1: interp method   (0 auto, 1 point driven, 2 tiled)
 *   key 2: FFT method      (0 auto, 1 cuFFT + pad/crop kernels, 2 fused zero-padding-aware passes,
 *          3 the same with the tiles of the strided passes loaded by TMA bulk tensor copies,
 *          4 pad/crop kernels + the library's own any-length FFT passes instead of cuFFT: every grid size,
 *          complex64 and complex128)
 *   key 3: timing experiments on the tiled spreader (bit 0: skip the tile flush, bit 1: skip the
 *          coil-value copies; results are then wrong -- never set outside a profiling session;
 *          bit 3: pretend the visit stream does not fit 32-bit indices, which exercises the
 *          hand-over to the point-driven kernels -- set before b200_plan_setpts)
 *   key 4: smallest coil class of the tiled kernels (0 = by the call's coil count T, else 1, 2, 4, 8,
 *          16 or 32: a call with T coils runs in the smallest class >= max(T, value); tests use it to
 *          push a small batch through every class)
 *   key 5: look-ahead, in CTAs, of the L2 prefetch of the strided FFT passes (0 = off)
 */
int b200_plan_set_option(b200_plan* plan, int key, int64_t value);

/*
 * Which coil class of the tiled spread / interp kernels a call with T coils runs in (the reference
 * has no counterpart: finufft loops over `n_trans`, cost proportional to coils,
 * src/mrinufft/operators/base.py:980-1010; here a warp's 32 lanes are T coils x 32/T row groups).
 *   out[0] = class (1, 2, 4, 8, 16, 32; 0 if the tiled kernels do not serve this plan)
 *   out[1] = (point, tile) visits of that class's visit stream, out[2] = stream entries
 *            (both 0 until a transform has built the stream), out[3] = 1 if the stream does not fit
 * B200_DOUBLE plans: out[0] = 16 (coils per launch of the complex128 spreader's row kernel) once setpts has
 * built its visit stream, 0 when the point-driven spreader serves the plan.
 */
int b200_plan_rows_class(b200_plan* plan, int T, int64_t out[4]);

/*
 * Timing hooks for the roofline report: the plan records CUDA events around its
 * dominant kernels.  out[0] = spread ms, out[1] = interp ms, out[2] = fft ms,
 * out[3] = pad/crop ms, out[4] = the row kernel alone (inside spread or interp) of the LAST execute
 * call (synchronises the stream).  With the fused FFT passes out[2] includes pad/crop and out[3] = 0.
 */
int b200_plan_last_timings(b200_plan* plan, float out[8]);
int b200_plan_enable_timing(b200_plan* plan, int on);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H */
