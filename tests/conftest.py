"""pytest configuration: import paths, the ``gpu`` marker, shared helpers."""

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "baseline" / "_ref"):
    if p.exists() and str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        d = {k: z[k] for k in z.files}
    d["shape"] = tuple(int(s) for s in d["shape"])
    if "n_coils" in d:
        d["n_coils"] = int(d["n_coils"])
    return d


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


GOLDEN_CASES = ["random2D", "random2D_sense", "random3D", "random3D_sense", "nyquist_radial2D",
                "spiral2D_sense", "grid2D", "cones3D"]


_NDFT_FULL = None


def ndft_full(samples, shape, **kw):
    """The reference's exact NDFT (``RawNDFT``, nudft_numpy.py:82-130) behind its complete
    ``FourierOperatorCPU`` surface -- ``MRInumpy`` itself takes neither ``n_batchs`` nor ``density``
    (nudft_numpy.py:141-150).  Registers a test-only backend name."""
    global _NDFT_FULL
    from mrinufft.operators.base import FourierOperatorCPU
    from mrinufft.operators.interfaces.nudft_numpy import RawNDFT

    if _NDFT_FULL is None:
        class NDFTFull(FourierOperatorCPU):
            backend = "ndft-full-test"
            available = True

        _NDFT_FULL = NDFTFull
    return _NDFT_FULL(samples, shape, raw_op=RawNDFT(samples, shape), **kw)


def check_orc_autograd(ag, orc, samples_unit, shape, t, smaps, n_coils, device, rng):
    """Forward / adjoint values and the data and field-map gradients of an autograd-wrapped
    off-resonance-corrected operator against torch's autograd through the dense model
    ``NDFT x exp(field_map t)`` -- the model of the reference's tests/operators/test_autodiff.py:109-120.
    Tolerances: the 24-segment time-segmentation interpolator itself is ~1e-3 from the dense model."""
    import torch

    def c(*s):
        return torch.from_numpy((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)).to(device)

    def rel(a, b):
        a, b = a.detach().cpu().numpy().ravel(), b.detach().cpu().numpy().ravel()
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    NK = len(t)
    assert ag.field_map.requires_grad and tuple(ag.field_map.shape) == tuple(shape)
    r = torch.stack(torch.meshgrid(*[torch.linspace(-s / 2, s / 2 - 1, s, dtype=torch.float64) for s in shape],
                                   indexing="ij"), 0).reshape(len(shape), -1)
    A = torch.exp(-2j * np.pi * (torch.from_numpy(samples_unit.astype(np.float64)) @ r))
    A = A / np.sqrt(np.prod(shape) * 2.0 ** len(shape))
    tt = torch.from_numpy(t.astype(np.float64)).reshape(-1, 1)
    fm0 = orc.field_map
    fm0 = fm0.detach().cpu() if torch.is_tensor(fm0) else torch.from_numpy(np.asarray(fm0))
    S = torch.ones((1, *shape), dtype=torch.complex128) if smaps is None else torch.from_numpy(smaps).to(torch.complex128)

    # forward direction
    fm = fm0.to(torch.complex128).requires_grad_(True)
    E = A * torch.exp(fm.reshape(1, -1) * tt)
    x = c(1, 1, *shape).requires_grad_(True)
    y_ref = c(1, n_coils, NK)
    y = ag.op(x)
    xd = x.detach().cpu().to(torch.complex128).requires_grad_(True)
    yd = torch.einsum("mn,cn->cm", E, (S * xd.reshape(1, *shape)).reshape(S.shape[0], -1))[None]
    assert rel(y, yd) < 3e-3
    torch.mean(torch.abs(y - y_ref) ** 2).backward()
    torch.mean(torch.abs(yd - y_ref.cpu()) ** 2).backward()
    assert rel(x.grad, xd.grad) < 5e-3
    assert rel(ag.field_map.grad, fm.grad) < 5e-3

    # adjoint direction
    ag.field_map.grad = None
    fm = fm0.to(torch.complex128).requires_grad_(True)
    E = A * torch.exp(fm.reshape(1, -1) * tt)
    k = c(1, n_coils, NK).requires_grad_(True)
    img_ref = c(1, 1, *shape)
    img = ag.adj_op(k)
    kd = k.detach().cpu().to(torch.complex128).requires_grad_(True)
    imgd = torch.sum(S.conj().reshape(S.shape[0], -1) * torch.einsum("mn,cm->cn", E.conj(), kd[0]), 0)
    imgd = imgd.reshape(1, 1, *shape)
    assert rel(img, imgd) < 3e-3
    torch.mean(torch.abs(img - img_ref) ** 2).backward()
    torch.mean(torch.abs(imgd - img_ref.cpu()) ** 2).backward()
    assert rel(k.grad, kd.grad) < 5e-3
    assert rel(ag.field_map.grad, fm.grad) < 5e-3
