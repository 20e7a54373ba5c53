"""Golden iterates of the reference's ``lsqr`` / ``lsmr`` / ``cg`` FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_solver_golden.py [output directory]

The unmodified reference solvers (``src/mrinufft/extras/optim.py:249-902``) run on the reference's
exact NDFT (``RawNDFT``, ``src/mrinufft/operators/interfaces/nudft_numpy.py:82-130``) wrapped in its
own ``FourierOperatorCPU`` so that batches and densities are available (``MRInumpy`` takes neither).
Stored per case: float32 sample locations (radians), inputs, and the complex128 image after every
iteration.  ``tests/test_gpu_parity.py`` replays them on the CUDA operator.
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/src")
from mrinufft.extras.optim import cg, lsmr, lsqr  # noqa: E402
from mrinufft.operators.base import FourierOperatorCPU  # noqa: E402
from mrinufft.operators.interfaces.nudft_numpy import RawNDFT  # noqa: E402
from scipy.stats import truncnorm  # noqa: E402

OUT = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent   # optional: output directory


class NDFTFull(FourierOperatorCPU):
    backend = "ndft-full-golden"
    available = True


def crandn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def iterates(image, operator, kspace_data, damp=0.0, x0=None):
    return np.array(image, copy=True).reshape(operator.img_full_shape).astype(np.complex128)


def make(name, shape, M, C, B, sense, density, damp, seed):
    rng = np.random.default_rng(seed)
    d = len(shape)
    s = truncnorm(-3, 3, loc=0, scale=0.16).rvs(size=(M, d), random_state=seed)
    samples = (s * 2 * np.pi).astype(np.float32)
    smaps = None
    if sense:
        smaps = crandn(rng, C, *shape)
        smaps /= np.linalg.norm(smaps, axis=0, keepdims=True)
    dens = rng.uniform(0.5, 1.5, M).astype(np.float32) if density else False
    op = NDFTFull(samples.astype(np.float64), shape, density=dens, n_coils=C, n_batchs=B, smaps=smaps,
                  raw_op=RawNDFT(samples.astype(np.float64), shape), squeeze_dims=True)
    x_true = crandn(rng, *op.img_full_shape)
    y = op.op(x_true).reshape(op.ksp_full_shape).astype(np.complex64)
    y += 0.01 * crandn(rng, *y.shape)
    out = dict(samples=samples, shape=np.asarray(shape), n_coils=C, n_batchs=B, y=y, damp=np.float64(damp))
    if sense:
        out["smaps"] = smaps
    if density:
        out["density"] = dens
    for fn in (lsqr, lsmr, cg):
        if density:
            op.density = dens  # the reference leaves it off after a run with a callback (optim.py:491-495)
        np.random.seed(99)      # power method of cg (base.py:1194)
        _, its = fn(op, y.copy(), damp=damp, max_iter=8, callback=iterates, progressbar=False)
        out[f"it_{fn.__name__}"] = np.stack(its)
        print(name, fn.__name__, len(its), "iterates, |x_last - x_true| / |x_true| =",
              np.linalg.norm(its[-1] - x_true) / np.linalg.norm(x_true))
    np.savez_compressed(OUT / f"{name}.npz", **out)


if __name__ == "__main__":
    make("solvers2D_sense", (16, 24), 600, 4, 1, True, False, 0.0, 21)
    # damp = 0 here: the reference's lsqr raises for n_batchs > 1 with damp > 0 (`if r1sq < 0` on an array,
    # optim.py:448)
    make("solvers2D_batch_density", (12, 10), 300, 2, 2, False, True, 0.0, 22)
    make("solvers3D_sense_damp", (8, 10, 6), 500, 3, 1, True, False, 0.05, 23)
