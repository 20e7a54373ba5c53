"""ctypes binding of ``libb200nufft.so`` (C ABI declared in ``include/b200nufft.h``).

The product path has NO fallback: if the shared library is missing or a call fails the
error is raised to the caller.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libb200nufft.so"

B200_SPREAD_ONLY = 1
B200_DOUBLE = 2
B200_EXACT_GRID = 4

# every symbol declared in include/b200nufft.h (tests check the export list against the header)
_SIGNATURES = {
    "b200_abi_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_plan_create": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_double, C.c_double,
         C.c_int, C.c_int],
    ),
    "b200_plan_destroy": (C.c_int, [C.c_void_p]),
    "b200_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "b200_plan_kernel_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "b200_plan_setpts": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "b200_plan_get_sort": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "b200_type2": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int,
         C.c_void_p],
    ),
    "b200_type1": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
         C.c_float, C.c_int, C.c_void_p],
    ),
    "b200_data_consistency": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
         C.c_float, C.c_void_p],
    ),
    "b200_toeplitz_apply": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
         C.c_void_p],
    ),
    "b200_spread": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200_interp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200_pipe_iteration": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200_launch_count": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "b200_plan_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "b200_plan_rows_class": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "b200_plan_last_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "b200_plan_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "b200_orc_weights": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p],
    ),
    "b200_fft_c2c": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_void_p]),
    "b200_vec_axpby": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int,
         C.c_void_p],
    ),
    "b200_vec_cg_dots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "b200_vec_cg_step": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p],
    ),
    "b200_vec_lsqr_step": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int,
         C.c_void_p],
    ),
    "b200_vec_lsmr_step": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
         C.c_void_p, C.c_int, C.c_void_p],
    ),
    "b200_stack_fftz_forward": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
         C.c_float, C.c_void_p],
    ),
    "b200_stack_fftz_adjoint": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
         C.c_float, C.c_void_p],
    ),
}

_lib = None


class B200Error(RuntimeError):
    """A call into libb200nufft.so failed."""


def library_built() -> bool:
    return LIB_PATH.exists()


def load():
    """Load the shared library (once) and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(
            f"{LIB_PATH} is missing: build it with `make -C {LIB_PATH.parent}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "The b200 backend has no CPU or library fallback."
        )
    lib = C.CDLL(os.fspath(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().b200_last_error()
        raise B200Error(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def launch_count(reset: bool = False):
    k, f = C.c_int64(0), C.c_int64(0)
    check(load().b200_launch_count(C.byref(k), C.byref(f), int(reset)), "b200_launch_count")
    return int(k.value), int(f.value)


class Plan:
    """Thin RAII wrapper of ``b200_plan`` (one device, one trajectory, both transform types)."""

    def __init__(self, shape, n_trans_max=1, eps=1e-6, upsampfac=2.0, spread_only=False, device=0,
                 double=False, exact_grid=False):
        self._lib = load()
        self._h = C.c_void_p(None)
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.device = int(device)
        self.n_trans_max = int(n_trans_max)
        n_modes = (C.c_int64 * 3)(*self.shape, *([1] * (3 - self.dim)))
        check(
            self._lib.b200_plan_create(
                C.byref(self._h), self.dim, n_modes, self.n_trans_max, float(eps), float(upsampfac),
                (B200_SPREAD_ONLY if spread_only else 0) | (B200_DOUBLE if double else 0)
                | (B200_EXACT_GRID if exact_grid else 0), self.device,
            ),
            "b200_plan_create",
        )
        self.n_samples = 0
        self._refresh_info()

    def _refresh_info(self):
        info = (C.c_int64 * 16)()
        check(self._lib.b200_plan_info(self._h, info), "b200_plan_info")
        d = self.dim
        self.nf = tuple(int(info[a]) for a in range(d))
        self.w = int(info[3])
        self.bins = tuple(int(info[4 + a]) for a in range(d))
        self.nbins = tuple(int(info[7 + a]) for a in range(d))
        self.poly_degree = int(info[10])
        self.n_samples = int(info[11])
        self.workspace_bytes = int(info[12])
        kp = (C.c_double * 4)()
        check(self._lib.b200_plan_kernel_params(self._h, kp), "b200_plan_kernel_params")
        self.beta, self.sigma = float(kp[0]), float(kp[2])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- calls (pointers are raw device addresses, stream a raw cudaStream_t) --
    def setpts(self, xyz_ptr, M, stream=0):
        check(self._lib.b200_plan_setpts(self._h, int(M), xyz_ptr, stream), "b200_plan_setpts")
        self.n_samples = int(M)

    def get_sort(self, origin_ptr, x1_ptr, key_ptr, perm_ptr, stream=0):
        check(
            self._lib.b200_plan_get_sort(self._h, origin_ptr, x1_ptr, key_ptr, perm_ptr, stream),
            "b200_plan_get_sort",
        )

    def type2(self, img, smaps, ksp, T, isign=-1, scale=1.0, conj_smaps=0, stream=0):
        check(
            self._lib.b200_type2(self._h, img, smaps, ksp, int(T), int(isign), float(scale),
                                 int(conj_smaps), stream),
            "b200_type2",
        )

    def type1(self, ksp, density, smaps, img, T, accumulate=0, isign=1, scale=1.0, conj_smaps=0,
              stream=0):
        check(
            self._lib.b200_type1(self._h, ksp, density, smaps, img, int(T), int(accumulate),
                                 int(isign), float(scale), int(conj_smaps), stream),
            "b200_type1",
        )

    def data_consistency(self, img, smaps, obs, density, grad, T, accumulate=0, scale=1.0, stream=0):
        check(
            self._lib.b200_data_consistency(self._h, img, smaps, obs, density, grad, int(T),
                                            int(accumulate), float(scale), stream),
            "b200_data_consistency",
        )

    def toeplitz_apply(self, img, smaps, kern, out, T, accumulate=0, scale=1.0, stream=0):
        check(
            self._lib.b200_toeplitz_apply(self._h, img, smaps, kern, out, int(T), int(accumulate),
                                          float(scale), stream),
            "b200_toeplitz_apply",
        )

    def spread(self, ksp, grid, T, stream=0):
        check(self._lib.b200_spread(self._h, ksp, grid, int(T), stream), "b200_spread")

    def interp(self, grid, ksp, T, stream=0):
        check(self._lib.b200_interp(self._h, grid, ksp, int(T), stream), "b200_interp")

    def pipe_iteration(self, d, stream=0):
        check(self._lib.b200_pipe_iteration(self._h, d, stream), "b200_pipe_iteration")

    def set_option(self, key, value):
        check(self._lib.b200_plan_set_option(self._h, int(key), int(value)), "b200_plan_set_option")

    def rows_class(self, n_trans):
        """Coil class the tiled kernels run a call with ``n_trans`` coils in, and the size of its stream."""
        out = (C.c_int64 * 4)()
        check(self._lib.b200_plan_rows_class(self._h, int(n_trans), out), "b200_plan_rows_class")
        return {"class": int(out[0]), "visits": int(out[1]), "entries": int(out[2]), "unsupported": bool(out[3])}

    def enable_timing(self, on=True):
        check(self._lib.b200_plan_enable_timing(self._h, int(on)), "b200_plan_enable_timing")

    def last_timings(self):
        out = (C.c_float * 8)()
        check(self._lib.b200_plan_last_timings(self._h, out), "b200_plan_last_timings")
        return {"spread_ms": out[0], "interp_ms": out[1], "fft_ms": out[2], "grid_ms": out[3],
                "rows_ms": out[4]}


def stack_fftz(adjoint, src, smaps, dst, zsel, C_, X, Y, Z, NZ, scale, stream=0):
    """z transform of the stacked operator (raw device pointers; see include/b200nufft.h)."""
    fn = load().b200_stack_fftz_adjoint if adjoint else load().b200_stack_fftz_forward
    check(fn(src, smaps, dst, zsel, int(C_), int(X), int(Y), int(Z), int(NZ), float(scale), stream),
          "b200_stack_fftz_adjoint" if adjoint else "b200_stack_fftz_forward")


def fft_c2c(ptr, T, shape, sign, double=False, stream=0):
    """In-place own FFT of T contiguous arrays of `shape` (raw device pointer; see include/b200nufft.h)."""
    n = (C.c_int64 * len(shape))(*[int(v) for v in shape])
    check(load().b200_fft_c2c(ptr, int(T), len(shape), n, int(sign), int(double), stream), "b200_fft_c2c")


def _next235even(n: int) -> int:
    n = max(int(n), 2)
    n += n & 1
    while True:
        m = n
        for f in (2, 3, 5):
            while m % f == 0:
                m //= f
        if m == 1:
            return n
        n += 2


def grid_size(shape, eps=1e-6, upsampfac=2.0, double=False, exact_grid=False):
    """The oversampled grid `b200_plan_create` chooses (mirror of csrc/api.cu, checked against `Plan.nf` by the
    GPU tests): `next235even(sigma N)` per axis; on the fastest axis the next size whose last 16-cell tile is whole
    or at least w - 1 cells wide; in 3-D single precision the next power of two when that is at most a third
    larger on every axis.  Used to size the workspace before a plan exists."""
    sigma = float(upsampfac) if upsampfac else 2.0
    if sigma == 2.0:
        w = int(np.ceil(np.log10(10.0 / eps)))
    else:
        w = int(np.ceil(-np.log(eps) / (np.pi * np.sqrt(1.0 - 1.0 / max(sigma, 1.001)))))
    w = min(max(w, 2), 16)
    dim = len(shape)
    nf = []
    for a, n in enumerate(shape):
        v = _next235even(max(int(np.ceil(sigma * int(n))), 2 * w))
        if a == dim - 1 and dim >= 2 and not exact_grid:
            while v % 16 != 0 and v % 16 < w - 1:
                v = _next235even(v + 2)
        nf.append(v)
    if dim == 3 and not double and not exact_grid:
        p2 = []
        for v in nf:
            q = 32
            while q < v:
                q <<= 1
            p2.append(q)
        if all(q <= 1.34 * v for q, v in zip(p2, nf)) and p2 != nf:
            nf = p2
    return tuple(nf)


def header_symbols(header: Path | None = None):
    """Names of the functions declared in include/b200nufft.h (used by the CPU tests)."""
    import re

    header = header or (_HERE.parent / "include" / "b200nufft.h")
    txt = header.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


__all__ = ["Plan", "B200Error", "load", "library_built", "launch_count", "header_symbols", "LIB_PATH"]
_ = np  # numpy is part of the public typing surface of this module
