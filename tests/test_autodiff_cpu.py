"""CPU checks of the autograd wrapper (mrinufft_b200/autodiff.py) -- the reference's own formulas
(src/mrinufft/operators/autodiff.py:14-154) without its hard ``deepinv`` import -- on the reference's
exact-NDFT backend, against torch's autograd through the dense model, as the reference's
tests/operators/test_autodiff.py does (``ndft_matrix`` there: NDFT x exp(field_map t) for
off-resonance-corrected operators).  No GPU and no CUDA library involved."""

import numpy as np
import pytest
import torch

import mrinufft
from mrinufft.operators.off_resonance import MRIFourierCorrected

from conftest import check_orc_autograd

from mrinufft_b200.autodiff import MRINufftAutoGrad


def _setup(n_coils, sense, n_shots, seed=0):
    rng = np.random.default_rng(seed)
    shape, NK = (8, 10), 48
    samples = rng.uniform(-0.5, 0.5, (n_shots * NK, 2)).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((n_coils, *shape)) + 1j * rng.standard_normal((n_coils, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("numpy")(samples, shape, n_coils=n_coils, smaps=smaps)
    op.squeeze_dims = False
    return rng, shape, NK, samples, smaps, op


@pytest.mark.parametrize("n_coils,sense", [(1, False), (3, True)])
def test_field_map_and_data_gradients_match_dense_model(n_coils, sense):
    # one shot: the reference's `full_readout_time` repeats each time point n_shots times
    # (off_resonance.py:218-222) while its op / adj_op lay k-space out shot-major (off_resonance.py:247-270);
    # the two agree for a single shot, which keeps the dense model exact
    rng, shape, NK, samples, smaps, op = _setup(n_coils, sense, n_shots=1)
    b0 = (30 * rng.standard_normal(shape)).astype(np.float32)
    t = np.linspace(0, 4e-3, NK).astype(np.float32)
    orc = MRIFourierCorrected(op, b0, t, interpolator={"name": "mti", "L": 24})
    orc.squeeze_dims = False
    ag = MRINufftAutoGrad(orc, wrt_data=True, wrt_traj=False, wrt_field_map=True)
    check_orc_autograd(ag, orc, samples, shape, t, smaps, n_coils, "cpu", rng)


def _c(rng, *s):
    return torch.from_numpy((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64))


def test_field_map_plumbing_and_errors():
    rng, shape, NK, samples, _, op = _setup(1, False, n_shots=2)
    plain = MRINufftAutoGrad(op, wrt_data=True)
    with pytest.raises(ValueError):
        plain.op(_c(rng, 1, 1, *shape), field_map=torch.zeros(shape))
    with pytest.raises(ValueError):
        _ = plain.field_map
    b0 = (10 * rng.standard_normal(shape)).astype(np.float32)
    t = np.linspace(0, 2e-3, NK).astype(np.float32)
    orc = MRIFourierCorrected(op, b0, t, interpolator={"name": "mti", "L": 6})
    orc.squeeze_dims = False
    ag = MRINufftAutoGrad(orc, wrt_data=True, wrt_field_map=True)
    B_before = np.array(orc.B, copy=True)
    ag.field_map = torch.from_numpy(np.asarray(orc.field_map) * 2.0)   # setter -> update_field_map
    assert not np.allclose(np.asarray(orc.B), B_before)                # interpolators were recomputed
    y = ag.op(_c(rng, 1, 1, *shape))
    assert tuple(y.shape) == (1, 1, 2 * NK)
    with pytest.raises(ValueError):
        op.squeeze_dims = True
        MRINufftAutoGrad(op, wrt_data=True)


def test_paired_batch_runs_item_by_item_with_its_own_maps_and_checks_shapes():
    """``paired_batch`` (autodiff.py:291-360, 408-456): item i uses smaps[i]; data gradients flow per item."""
    rng, shape, NK, samples, smaps, op = _setup(3, True, n_shots=2)
    K, D = 2 * NK, 3
    maps = np.stack([smaps * np.exp(1j * 0.3 * d) * (1 + 0.1 * d) for d in range(D)]).astype(np.complex64)
    ag = MRINufftAutoGrad(op, wrt_data=True, paired_batch=D)
    x = _c(rng, D, 1, 1, *shape).requires_grad_(True)
    y = ag.op(x, smaps=maps)
    assert tuple(y.shape) == (D, 1, 3, K)
    for d in range(D):
        op.smaps = maps[d]
        want = op.op(x[d].detach().numpy())
        assert np.allclose(y[d].detach().numpy(), want, rtol=1e-5, atol=1e-6)
    w = _c(rng, D, 1, 3, K)
    torch.sum(torch.abs(y - w) ** 2).backward()
    # torch's convention: the gradient of ||A_d x_d - w_d||^2 is 2 A_d^H (A_d x_d - w_d) -- with item d's OWN
    # maps (the reference's shared-operator design would use the last item's for every item)
    for d in range(D):
        op.smaps = maps[d]
        want = 2 * op.adj_op((y[d] - w[d]).detach().numpy())
        assert np.allclose(x.grad[d].numpy(), want.reshape(x.grad[d].shape), rtol=1e-4, atol=1e-5)
    k = _c(rng, D, 1, 3, K)
    img = ag.adj_op(k, smaps=maps)
    assert tuple(img.shape) == (D, 1, 1, *shape)
    with pytest.raises(ValueError, match="smaps and image"):
        ag.op(x, smaps=maps[:2])
    with pytest.raises(ValueError, match="smaps and k-space"):
        ag.adj_op(k, smaps=maps[:2])
    with pytest.raises(ValueError, match="k-space and samples"):
        ag.adj_op(k, samples=torch.zeros(D, K + 1, 2))
    with pytest.raises(ValueError, match="samples loc and image"):
        ag.op(x, samples=torch.zeros(D, K, 3))
