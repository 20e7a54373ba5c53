"""CPU checks of the Toeplitz kernel assembly (mrinufft_b200/toeplitz.py) against the reference's own
``compute_toeplitz_kernel`` / ``apply_toeplitz_kernel`` (src/mrinufft/operators/toeplitz.py:35-269)
run on its exact-NDFT ``numpy`` backend.  No GPU, no CUDA library involved."""

import numpy as np
import pytest
import torch

import mrinufft
from mrinufft.operators.toeplitz import apply_toeplitz_kernel, compute_toeplitz_kernel

from mrinufft_b200.toeplitz import assemble_toeplitz_kernel, modulated_weights


@pytest.mark.parametrize("shape", [(12, 16), (8, 6, 10)])
@pytest.mark.parametrize("with_density", [False, True])
def test_kernel_assembly_matches_reference(shape, with_density):
    rng = np.random.default_rng(3)
    M = 300
    samples = rng.uniform(-0.5, 0.5, (M, len(shape))).astype(np.float32)
    ref_op = mrinufft.get_operator("numpy")(samples, shape)
    if with_density:
        ref_op.density = rng.uniform(0.5, 1.5, M).astype(np.float32)
    ref_kernel = compute_toeplitz_kernel(ref_op, ref_op.density)

    omega = torch.from_numpy(np.ascontiguousarray(ref_op.samples)).to(torch.float32)
    if np.abs(ref_op.samples).max() <= 0.5 + 1e-4:
        omega = omega * (2 * np.pi)
    w = torch.from_numpy(ref_op.density.astype(np.float32)) if with_density else torch.ones(M)
    raw = mrinufft.get_operator("numpy")(samples, shape)  # no density: raw adjoint

    def adj(signs):
        ksp = modulated_weights(w, omega, signs, shape).numpy()
        out = np.empty(shape, dtype=np.complex64)
        raw._adj_op(ksp, out)
        return torch.from_numpy(out)

    kern = assemble_toeplitz_kernel(adj, shape, 1.0 / float(ref_op.norm_factor)).numpy()
    assert kern.shape == tuple(2 * s for s in shape)
    assert np.linalg.norm(kern - ref_kernel) <= 2e-5 * np.linalg.norm(ref_kernel)

    # ... and the embedding reproduces adj_op(op(x)) of the exact NDFT (circular convolution on 2N)
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    direct = ref_op.adj_op(ref_op.op(x))
    toep = apply_toeplitz_kernel(x, kern.astype(np.float32))
    assert np.linalg.norm(toep - direct) <= 5e-5 * np.linalg.norm(direct)


def test_centred_padding_is_equivalent_to_corner_padding():
    """b200_toeplitz_apply pads / crops at the mode-centred position of the NUFFT grid, the reference
    at the top-left corner: a circular convolution commutes with the shift."""
    rng = np.random.default_rng(0)
    N = (8, 6)
    kern = rng.uniform(0.5, 2.0, tuple(2 * n for n in N))
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    corner = np.zeros(kern.shape, complex)
    corner[: N[0], : N[1]] = x
    ref = np.fft.ifftn(np.fft.fftn(corner) * kern)[: N[0], : N[1]]
    centred = np.zeros(kern.shape, complex)
    idx = [(np.arange(n) - n // 2) % (2 * n) for n in N]
    centred[np.ix_(*idx)] = x
    out = np.fft.ifftn(np.fft.fftn(centred) * kern)[np.ix_(*idx)]
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-12)
