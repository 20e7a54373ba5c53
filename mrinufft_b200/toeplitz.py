"""Toeplitz embedding of the Gram operator ``A^H W A`` -- kernel assembly (plan time).

Host-side mirror of ``compute_toeplitz_kernel`` / ``_compute_toep_2d`` / ``_compute_toep_3d``
(``src/mrinufft/operators/toeplitz.py:35-200``): the point-spread function of the trajectory is
needed on the ``2N - 1`` lags of every axis; ``2^(d-1)`` raw adjoints of phase-modulated weights
give the lag windows ``[-N/2, N/2)`` shifted onto the outer lags, Hermitian symmetry (real weights)
gives the other half, and one real inverse FFT turns the lag array into the real spectrum of the
circulant embedding on the ``2N`` grid.  The adjoints are the expensive part and run in
``libb200nufft.so``; what is here is index bookkeeping on torch tensors (any device).

The spectrum is applied by ``b200_toeplitz_apply`` (``include/b200nufft.h``).
"""

from __future__ import annotations

import torch


def _rev(t: torch.Tensor) -> torch.Tensor:
    """``t[N-1:0:-1]`` along every axis (drop index 0, reverse the rest)."""
    idx = tuple(slice(1, None) for _ in range(t.ndim))
    return torch.flip(t[idx], dims=tuple(range(t.ndim)))


def assemble_toeplitz_kernel(adj, shape, scale: float) -> torch.Tensor:
    """Real spectrum ``(2N0, 2N1[, 2N2])`` of the Toeplitz embedding.

    ``adj(signs)`` must return the raw (un-normalised, no smaps) adjoint NUFFT, a complex tensor
    of ``shape``, of the weights modulated by ``exp(i omega . (signs * N/2))``
    (``_modulated_weights``, toeplitz.py:98-111).  ``scale`` is ``1 / norm_factor``.
    """
    shape = tuple(int(s) for s in shape)
    if any(s % 2 for s in shape):
        raise ValueError(f"Toeplitz kernel computation only supports even grid sizes, got {shape}.")
    if len(shape) == 2:
        N0, N1 = shape
        A = adj((1, 1))
        kernel = torch.zeros((2 * N0, N1 + 1), dtype=A.dtype, device=A.device)
        kernel[:N0, :N1] = A
        B = adj((1, -1))
        kernel[N0 + 1:, 0] = torch.conj(_rev(A[:, 0]))
        kernel[N0 + 1:, 1:N1] = torch.conj(_rev(B))
    elif len(shape) == 3:
        N0, N1, N2 = shape
        A = adj((1, 1, 1))
        kernel = torch.zeros((2 * N0, 2 * N1, N2 + 1), dtype=A.dtype, device=A.device)
        kernel[:N0, :N1, :N2] = A
        C = adj((1, -1, 1))
        kernel[:N0, N1 + 1:, :N2] = C[:, 1:, :]
        B = adj((1, 1, -1))
        D = adj((1, -1, -1))
        kernel[N0 + 1:, 0, 0] = torch.conj(_rev(A[:, 0, 0]))
        kernel[N0 + 1:, 0, 1:N2] = torch.conj(_rev(B[:, 0, :]))
        kernel[N0 + 1:, 1:N1, 0] = torch.conj(_rev(C[:, :, 0]))
        kernel[N0 + 1:, 1:N1, 1:N2] = torch.conj(_rev(D))
        kernel[N0 + 1:, N1 + 1:, 0] = torch.conj(_rev(A[:, :, 0]))
        kernel[N0 + 1:, N1 + 1:, 1:N2] = torch.conj(_rev(B))
    else:
        raise ValueError(f"Toeplitz kernel calculation not implemented for ndim={len(shape)}")
    full_shape = tuple(2 * s for s in shape)
    return torch.fft.irfftn(torch.conj(kernel) * scale, s=full_shape, norm="ortho")


def modulated_weights(weights: torch.Tensor, omega: torch.Tensor, signs, shape) -> torch.Tensor:
    """``weights * exp(i omega . (signs * N/2))`` as complex64; the phase is formed in float64."""
    shift = torch.tensor([sg * (n // 2) for sg, n in zip(signs, shape)], dtype=torch.float64,
                         device=omega.device)
    ph = omega.to(torch.float64) @ shift
    w = weights.to(torch.float64)
    return torch.complex((w * torch.cos(ph)).to(torch.float32), (w * torch.sin(ph)).to(torch.float32))
