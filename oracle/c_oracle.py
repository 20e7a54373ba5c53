"""CPU ORACLE (test infrastructure, NOT product code) -- ctypes wrapper of ``es_spread.c``.

``CpuNufft`` is the finufft-algorithm CPU restatement used (a) as the fast checker for problems
too large for the numpy loops of ``es_nufft.py`` and (b) as ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg (the reference's own CPU backend, finufft, is not installable offline).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import scipy.fft as sfft

from . import es_nufft as E

_HERE = Path(__file__).resolve().parent
_libs = {}


def build():
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


def _load(precision: str):
    if precision in _libs:
        return _libs[precision]
    path = _HERE / f"liboracle_{precision}.so"
    if not path.exists():
        build()
    lib = C.CDLL(os.fspath(path))
    lib.oracle_fold.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p]
    lib.oracle_interp.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_spread.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    lib.oracle_max_threads.restype = C.c_int
    _libs[precision] = lib
    return lib


def max_threads() -> int:
    return int(_load("f64").oracle_max_threads())


def set_threads(n: int | None = None) -> int:
    """Use ``n`` OpenMP threads (default: every CPU this process may run on) in both builds of the
    library -- ``torch.distributed.run`` exports ``OMP_NUM_THREADS=1`` to its ranks.  Returns ``n``."""
    n = int(n) if n else len(os.sched_getaffinity(0))
    for prec in ("f64", "f32"):
        lib = _load(prec)
        lib.oracle_set_threads.argtypes = [C.c_int]
        lib.oracle_set_threads(n)
    return n


class PortRawOp:
    """The port behind the reference's ``raw_op`` contract (``FourierOperatorSimple._op/_adj_op``,
    base.py:1013, 1068-1073; finufft.py:64-76): ``op(coeffs_out, image_in)``, ``adj_op(coeffs_in,
    image_out)`` on ``(T, *shape)`` / ``(T, M)`` arrays, no smaps, no normalisation.  Lets the timing legs
    run the reference's own ``FourierOperatorCPU`` coil loop around the finufft-algorithm port."""

    def __init__(self, samples, shape, eps=1e-6, precision="f32", workers=None):
        self.cpu = CpuNufft(samples, shape, eps=eps, precision=precision, workers=workers)
        self.shape, self.n_samples = self.cpu.shape, self.cpu.M

    def op(self, coeffs, image):
        np.copyto(coeffs.reshape(-1, self.n_samples), self.cpu.type2(image.reshape(-1, *self.shape)))
        return coeffs

    def adj_op(self, coeffs, image):
        np.copyto(image.reshape(-1, *self.shape), self.cpu.type1(coeffs.reshape(-1, self.n_samples)))
        return image


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def fold(x, nf, w):
    """C version of ``es_nufft.fold_points`` (same bits): origin int32, x1 float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    M = x.shape[0]
    origin = np.empty(M, np.int32)
    x1 = np.empty(M, np.float32)
    _load("f64").oracle_fold(_ptr(x), M, 1, int(nf), int(w), _ptr(origin), _ptr(x1), None)
    return origin, x1


class CpuNufft:
    """finufft-algorithm NUFFT on the host cores (OpenMP spread/interp + pocketfft via scipy).

    precision "f64": complex128 checker; "f32": complex64, the timing baseline.
    """

    def __init__(self, samples, shape, eps=1e-6, sigma=2.0, precision="f64", bins=None, workers=None):
        self.lib = _load(precision)
        self.workers = -1 if workers is None else int(workers)  # scipy.fft threads
        self.rdt = np.float64 if precision == "f64" else np.float32
        self.cdt = np.complex128 if precision == "f64" else np.complex64
        self.samples = np.ascontiguousarray(samples, dtype=np.float32)
        self.shape = tuple(int(s) for s in shape)
        self.d = len(self.shape)
        self.M = self.samples.shape[0]
        self.w, self.beta = E.kernel_params(eps, sigma)
        self.nfs = tuple(E.fine_grid_size(n, self.w, sigma) for n in self.shape)
        self.nf_arr = np.asarray(self.nfs, np.int32)
        self.deapod = [E.deapod_vector(n, nf, self.w, self.beta).astype(self.rdt)
                       for n, nf in zip(self.shape, self.nfs)]
        # setpts: fold + bin sort (pencil bins, see es_nufft.bin_sort)
        self.origin = np.empty((self.d, self.M), np.int32)
        self.x1 = np.empty((self.d, self.M), np.float32)
        for a in range(self.d):
            self.origin[a], self.x1[a] = fold(self.samples[:, a], self.nfs[a], self.w)
        bins = bins or E.default_bins(self.d)
        key = E.make_key(self.origin, self.nfs, bins, self.w)
        self.perm = np.argsort(key, kind="stable").astype(np.int32)
        self._mode_idx = [(np.arange(n) - n // 2) % nf for n, nf in zip(self.shape, self.nfs)]

    def _dgrid(self):
        dg = self.deapod[0]
        for a in range(1, self.d):
            dg = np.multiply.outer(dg, self.deapod[a])
        return dg

    def spread(self, c):
        c = np.ascontiguousarray(c, dtype=self.cdt).reshape(-1, self.M)
        T = c.shape[0]
        fw = np.zeros((T, *self.nfs), self.cdt)
        self.lib.oracle_spread(self.d, _ptr(self.nf_arr), self.w, self.beta, self.M,
                               _ptr(self.origin), _ptr(self.x1), _ptr(self.perm), T, _ptr(c),
                               _ptr(fw), 0)
        return fw

    def interp(self, fw):
        fw = np.ascontiguousarray(fw, dtype=self.cdt).reshape(-1, *self.nfs)
        T = fw.shape[0]
        c = np.empty((T, self.M), self.cdt)
        self.lib.oracle_interp(self.d, _ptr(self.nf_arr), self.w, self.beta, self.M,
                               _ptr(self.origin), _ptr(self.x1), _ptr(self.perm), T, _ptr(fw),
                               _ptr(c))
        return c

    def type2(self, img, isign=-1):
        img = np.asarray(img, dtype=self.cdt).reshape((-1, *self.shape))
        T = img.shape[0]
        fw_hat = np.zeros((T, *self.nfs), self.cdt)
        fw_hat[(slice(None), *np.ix_(*self._mode_idx))] = img * self._dgrid()[None]
        axes = tuple(range(1, self.d + 1))
        if isign < 0:
            fw = sfft.fftn(fw_hat, axes=axes, workers=self.workers, overwrite_x=True)
        else:
            fw = sfft.ifftn(fw_hat, axes=axes, norm="forward", workers=self.workers, overwrite_x=True)
        return self.interp(fw)

    def type1(self, c, isign=+1):
        fw = self.spread(c)
        axes = tuple(range(1, self.d + 1))
        if isign > 0:
            F = sfft.ifftn(fw, axes=axes, norm="forward", workers=self.workers, overwrite_x=True)
        else:
            F = sfft.fftn(fw, axes=axes, workers=self.workers, overwrite_x=True)
        return F[(slice(None), *np.ix_(*self._mode_idx))] * self._dgrid()[None]

    # the reference's operator semantics (base.py:949-1073): smaps, density, 1/norm on both sides
    def op(self, image, smaps=None):
        norm = np.sqrt(np.prod(self.shape) * 2.0 ** self.d)
        image = np.asarray(image, dtype=self.cdt)
        if smaps is not None:
            out = np.empty((smaps.shape[0], self.M), self.cdt)
            for c in range(smaps.shape[0]):
                out[c] = self.type2(image * smaps[c])[0]
        else:
            out = self.type2(image)
        return out / norm

    def adj_op(self, ksp, smaps=None, density=None):
        norm = np.sqrt(np.prod(self.shape) * 2.0 ** self.d)
        ksp = np.asarray(ksp, dtype=self.cdt).reshape(-1, self.M)
        if density is not None:
            ksp = ksp * density
        if smaps is not None:
            img = np.zeros(self.shape, self.cdt)
            for c in range(smaps.shape[0]):
                img += np.conj(smaps[c]) * self.type1(ksp[c])[0]
        else:
            img = self.type1(ksp)
        return img / norm
