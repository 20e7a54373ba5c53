"""Toeplitz embedding of the Gram operator ``A^H W A`` -- kernel assembly (plan time).

Role of ``compute_toeplitz_kernel`` (``src/mrinufft/operators/toeplitz.py:35-200``).  The Gram operator of a
NUFFT is a convolution with the point-spread function of the trajectory,

    c[l] = sum_j w_j exp(i omega_j . l),      l in [-(N-1), N-1]^d,

and applying it to an image of ``N`` voxels per axis only ever needs those ``2N - 1`` lags per axis: zero-padded
to ``2N`` it is a circular convolution, i.e. a multiply by the (real, because ``c[-l] = conj(c[l])`` for real
weights) DFT of the lag array.  A raw adjoint NUFFT of the weights gives ``c`` on the ``N`` lags
``[-N/2, N/2)``; modulating the weights by ``exp(i omega . s N/2)``, ``s = +-1`` per axis, shifts that window to
``[0, N)`` or ``[-N, 0)``.  The ``2^d`` sign combinations tile ``[-N, N)^d``: every adjoint result is dropped
into its octant of the lag array (lag ``l`` at index ``l mod 2N``), the lag ``-N`` of every axis -- which a
zero-padded convolution never reads -- is cleared, and one FFT gives the spectrum.  (The reference halves the
number of adjoints with the Hermitian symmetry and a real inverse FFT of a half array; with single-coil
adjoints at 3 ms apiece -- coil class 1 of the row kernels -- the plain tiling is not worth complicating.)

The adjoints are the expensive part and run in ``libb200nufft.so``, and so does the final FFT of the lag
array (``b200_fft_c2c``); what is here is bookkeeping on torch tensors (any device, any dimension).  The spectrum is applied by ``b200_toeplitz_apply``
(``include/b200nufft.h``).
"""

from __future__ import annotations

import itertools
import math

import torch


def assemble_toeplitz_kernel(adj, shape, scale: float) -> torch.Tensor:
    """Real spectrum, of shape ``2N``, of the Toeplitz embedding.

    ``adj(signs)`` must return the raw (un-normalised, no smaps) adjoint NUFFT, a complex tensor of ``shape``,
    of the weights modulated by ``exp(i omega . (signs * N/2))`` (``modulated_weights``).  ``scale`` is
    ``1 / norm_factor``; the result carries ``scale / sqrt(prod(2N))``, the convention of the reference's
    ``irfftn(..., norm="ortho")`` kernel (toeplitz.py:89-93).
    """
    shape = tuple(int(n) for n in shape)
    if any(n % 2 for n in shape):
        raise ValueError(f"Toeplitz kernel computation only supports even grid sizes, got {shape}.")
    full = tuple(2 * n for n in shape)
    lags = None
    for signs in itertools.product((1, -1), repeat=len(shape)):
        window = adj(signs)  # c[n + (s - 1) N / 2]: lags [0, N) for s = +1, [-N, 0) for s = -1
        if lags is None:
            lags = torch.zeros(full, dtype=window.dtype, device=window.device)
        lags[tuple(slice(0, n) if s > 0 else slice(n, 2 * n) for s, n in zip(signs, shape))] = window
    for axis, n in enumerate(shape):
        lags.select(axis, n).zero_()
    return _fftn(lags).real * (scale / math.sqrt(math.prod(full)))


def _fftn(lags: torch.Tensor) -> torch.Tensor:
    """Forward DFT of the lag array: the library's own any-length FFT on the device (``b200_fft_c2c``).  CPU
    tensors only occur in the host-logic test (tests/test_toeplitz_cpu.py, reference NDFT, no GPU)."""
    if not lags.is_cuda:
        return torch.fft.fftn(lags)
    from . import _lib

    out = lags.contiguous().clone()
    with torch.cuda.device(out.device):
        _lib.fft_c2c(out.data_ptr(), 1, out.shape, -1, out.dtype == torch.complex128,
                     torch.cuda.current_stream(out.device).cuda_stream)
    return out


def modulated_weights(weights: torch.Tensor, omega: torch.Tensor, signs, shape) -> torch.Tensor:
    """``weights * exp(i omega . (signs * N/2))`` as complex64; the phase is formed in float64."""
    shift = torch.tensor([sg * (n // 2) for sg, n in zip(signs, shape)], dtype=torch.float64,
                         device=omega.device)
    ph = omega.to(torch.float64) @ shift
    w = weights.to(torch.float64)
    return torch.complex((w * torch.cos(ph)).to(torch.float32), (w * torch.sin(ph)).to(torch.float32))
