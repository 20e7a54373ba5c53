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
