"""CPU checks of the device-resident solvers (mrinufft_b200/solvers.py) against the reference's own
``cg`` / ``lsqr`` / ``lsmr`` (src/mrinufft/extras/optim.py:249-902), iterate by iterate.

Both sides run on the reference's exact-NDFT ``numpy`` backend: the reference solvers drive it
directly, ours drive it through a thin wrapper that exposes the device-level surface the solvers use
(``_op_device`` / ``_adj_device`` / ``_dc_device`` on torch tensors -- CPU tensors here).  No GPU and no
CUDA library involved: this pins the solver logic, the GPU suite pins it on the CUDA operator.
"""

import numpy as np
import pytest
import torch

from mrinufft.extras.optim import cg as ref_cg
from mrinufft.extras.optim import lsmr as ref_lsmr
from mrinufft.extras.optim import lsqr as ref_lsqr

from conftest import ndft_full

from mrinufft_b200 import solvers


class TorchFacade:
    """The reference NDFT operator behind the device-level surface of MRIB200NUFFT (test only)."""

    def __init__(self, ref):
        self.ref = ref
        self.device = torch.device("cpu")
        self._cdt = torch.complex64

    def __getattr__(self, name):
        return getattr(self.ref, name)

    @property
    def density(self):
        return self.ref.density

    @density.setter
    def density(self, value):
        self.ref.density = value

    @property
    def squeeze_dims(self):
        return self.ref.squeeze_dims

    def _op_device(self, img):
        y = self.ref.op(img.numpy().reshape(self.ref.img_full_shape))
        return torch.from_numpy(np.ascontiguousarray(y).reshape(self.ref.ksp_full_shape).astype(np.complex64))

    def _adj_device(self, ksp):
        x = self.ref.adj_op(ksp.numpy().reshape(self.ref.ksp_full_shape))
        return torch.from_numpy(np.ascontiguousarray(x).reshape(self.ref.img_full_shape).astype(np.complex64))

    def _dc_device(self, img, obs):
        g = self.ref.data_consistency(img.numpy().reshape(self.ref.img_full_shape),
                                      obs.numpy().reshape(self.ref.ksp_full_shape))
        return torch.from_numpy(np.ascontiguousarray(g).reshape(self.ref.img_full_shape).astype(np.complex64))


def _problem(n_batchs, sense, density, seed=0):
    rng = np.random.default_rng(seed)
    shape, M, C = (10, 12), 400, 3
    samples = rng.uniform(-0.5, 0.5, (M, 2)).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0, keepdims=True)
    dens = rng.uniform(0.5, 1.5, M).astype(np.float32) if density else False
    op = ndft_full(samples, shape, density=dens, n_coils=C, n_batchs=n_batchs, smaps=smaps, squeeze_dims=True)
    x_true = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape))
    y = op.op(x_true.astype(np.complex64)).reshape(op.ksp_full_shape).astype(np.complex64)
    y += 0.01 * (rng.standard_normal(y.shape) + 1j * rng.standard_normal(y.shape)).astype(np.complex64)
    x0 = (0.1 * (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape))
          ).astype(np.complex64)
    return op, y, x0


def _iterates(image, operator, kspace_data, damp=0.0, x0=None):
    return np.array(image, copy=True).reshape(operator.img_full_shape)


def _close(mine, ref, tol):
    assert len(mine) == len(ref)
    for k, (a, b) in enumerate(zip(mine, ref)):
        err = np.linalg.norm(a - b) / np.linalg.norm(b)
        assert err < tol, f"iterate {k}: rel err {err:.2e}"


@pytest.mark.parametrize("name,ref_fn", [("lsqr", ref_lsqr), ("lsmr", ref_lsmr)])
@pytest.mark.parametrize("n_batchs,sense,density,damp,use_x0", [
    (1, True, False, 0.0, False),
    (2, True, False, 0.0, False),
    (1, False, False, 0.3, True),
    (2, False, True, 0.0, False),
    (1, True, True, 0.2, False),
])
def test_bidiagonalisation_solvers_match_reference_iterates(name, ref_fn, n_batchs, sense, density, damp, use_x0):
    op, y, x0 = _problem(n_batchs, sense, density)
    kw = dict(damp=damp, max_iter=12, callback=_iterates, progressbar=False)
    if use_x0:
        kw["x0"] = x0
    dens_before = None if op.density is None else op.density.copy()
    x_ref, it_ref = ref_fn(op, y.copy(), **{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
    # the reference leaves the density switched off when a callback is given (optim.py:491-495): undo
    if dens_before is not None:
        op.density = dens_before
    y_in = y.copy()
    x, it = solvers.SOLVERS[name](TorchFacade(op), y_in, **kw)
    assert np.array_equal(y_in, y)                      # inputs are never written to
    assert isinstance(x, np.ndarray) and x.shape == x_ref.shape
    _close(it, it_ref, 2e-4)
    assert np.linalg.norm(x - x_ref) < 2e-4 * np.linalg.norm(x_ref)
    if dens_before is not None:                          # ... and the density is back in place
        assert op.density is not None and np.array_equal(op.density, dens_before)


@pytest.mark.parametrize("n_batchs,sense,density,damp", [(1, True, False, 0.0), (1, False, False, 0.1),
                                                         (1, True, True, 0.0)])
def test_cg_matches_reference_iterates(n_batchs, sense, density, damp):
    op, y, _ = _problem(n_batchs, sense, density, seed=4)
    kw = dict(damp=damp, max_iter=8, callback=_iterates, progressbar=False)
    dens_before = None if op.density is None else op.density.copy()
    np.random.seed(7)   # the power method starts from np.random.random (base.py:1194)
    x_ref, it_ref = ref_cg(op, y.copy(), **kw)
    if dens_before is not None:
        op.density = dens_before
    np.random.seed(7)
    x, it = solvers.cg(TorchFacade(op), y.copy(), **kw)
    _close(it, it_ref, 2e-4)
    assert np.linalg.norm(x - x_ref) < 2e-4 * np.linalg.norm(x_ref)


def test_stopping_rule_and_zero_rhs():
    op, y, _ = _problem(1, True, False, seed=2)
    # a consistent system converges and stops on its own, at the same iteration as the reference
    y_clean = op.op(op.adj_op(y)).reshape(op.ksp_full_shape).astype(np.complex64)
    for name, ref_fn in (("lsqr", ref_lsqr), ("lsmr", ref_lsmr)):
        _, it_ref = ref_fn(op, y_clean.copy(), max_iter=400, atol=1e-3, btol=1e-3, callback=_iterates,
                           progressbar=False)
        _, it = solvers.SOLVERS[name](TorchFacade(op), y_clean.copy(), max_iter=400, atol=1e-3, btol=1e-3,
                                      callback=_iterates)
        assert len(it) < 400 and abs(len(it) - len(it_ref)) <= 1
        # y = 0: alpha * beta == 0 -> the start is returned as is (optim.py:373-374)
        z = solvers.SOLVERS[name](TorchFacade(op), np.zeros_like(y))
        assert z.shape == tuple(op.img_full_shape) and not np.any(z)


def test_givens_follows_the_reference_branches():
    from mrinufft.extras.optim import _sym_ortho

    rng = np.random.default_rng(0)
    cases = [(np.float32([3.0]), np.float32([4.0])), (np.float32([4.0]), np.float32([-3.0])),
             (np.float32([0.0]), np.float32([2.0])), (np.float32([2.0]), np.float32([0.0])),
             (rng.standard_normal(3).astype(np.float32), rng.standard_normal(3).astype(np.float32)),
             (np.float32([1.0, 0.0]), np.float32([2.0, 1.0]))]
    for a, b in cases:
        got = solvers._givens(a, b)
        want = _sym_ortho(a, b)
        for g, w in zip(got, want):
            assert np.allclose(g, w, rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["lsqr", "lsmr"])
def test_damped_batches_equal_the_batches_solved_one_by_one(name):
    """``damp > 0`` with ``n_batchs > 1``: the reference's lsqr raises there (``if r1sq < 0`` on an array,
    optim.py:448), so the check is against the same solver run on every batch volume separately -- the
    per-batch scalars must not leak between volumes."""
    op2, y2, _ = _problem(2, False, False, seed=5)
    x2 = solvers.SOLVERS[name](TorchFacade(op2), y2.copy(), damp=0.3, max_iter=8)
    op1, _, _ = _problem(1, False, False, seed=5)     # same samples (same seed), one volume at a time
    assert np.array_equal(op1.samples, op2.samples)
    for b in range(2):
        x1 = solvers.SOLVERS[name](TorchFacade(op1), y2[b:b + 1].copy(), damp=0.3, max_iter=8)
        assert np.linalg.norm(x2[b] - x1.reshape(x2[b].shape)) < 1e-4 * np.linalg.norm(x1)
    if name == "lsqr":
        with pytest.raises(ValueError):
            ref_lsqr(op2, y2.copy(), damp=0.3, max_iter=3, progressbar=False)


@pytest.mark.timeout(300)
def test_committed_solver_goldens_are_what_the_reference_produces(tmp_path):
    """Re-run tests/golden/make_solver_golden.py against the reference checkout (build container only) and
    compare with the committed fixtures the GPU suite replays."""
    import subprocess
    import sys
    from pathlib import Path

    if not Path("/root/reference/src/mrinufft").is_dir():
        pytest.skip("the reference checkout is only present in the build container")
    golden = Path(__file__).resolve().parent / "golden"
    r = subprocess.run([sys.executable, str(golden / "make_solver_golden.py"), str(tmp_path)],
                       capture_output=True, text=True, cwd=tmp_path, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("solvers2D_sense", "solvers2D_batch_density", "solvers3D_sense_damp"):
        with np.load(golden / f"{name}.npz") as a, np.load(tmp_path / f"{name}.npz") as b:
            assert sorted(a.files) == sorted(b.files)
            for k in a.files:
                if k.startswith("it_cg"):   # cg's step size comes from a 10-step power method: looser
                    assert np.allclose(a[k], b[k], rtol=1e-4, atol=1e-6), (name, k)
                else:
                    assert np.allclose(a[k], b[k], rtol=1e-6, atol=1e-8), (name, k)


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_cg_leaves_a_density_compensated_start_and_diverges_and_so_does_ours(dim):
    """BASELINE configs[3] combines density="pipe" with pinv_solver(optim="cg").  The reference's `cg`
    (extras/optim.py:839-862) takes its step size from the density-WEIGHTED operator and then iterates on the
    un-weighted one.  Pipe's weights are normalised so that the weighted operator has a Lipschitz constant of a
    few units, while the un-weighted operator of a radial trajectory (dense centre) has one that is many times
    larger: the fixed step overshoots and the iteration diverges from the good density-compensated start it was
    given -- on the reference's own exact NDFT, with the reference's own solver.  Ours mirrors the reference
    statement by statement, so it follows the same path; handing `cg` the Lipschitz constant of the operator it
    iterates on (`lipschitz_cst=`, an argument the reference does not have) gives the monotone reconstruction
    that tools/bench_configs.py reports for that configuration."""
    from mrinufft.density import voronoi
    from mrinufft.trajectories import initialize_2D_radial, initialize_3D_phyllotaxis_radial

    if dim == 3:
        shape, traj = (12, 12, 12), initialize_3D_phyllotaxis_radial(96, 24).reshape(-1, 3).astype(np.float32)
    else:
        shape, traj = (32, 32), initialize_2D_radial(48, 64).reshape(-1, 2).astype(np.float32)
    # weights on Pipe's scale: mean |A^H D A 1| = 1 (finufft.py:236-244)
    dens = voronoi(traj, shape).astype(np.float32)
    plain = ndft_full(traj, shape, n_coils=1, squeeze_dims=False)
    dens /= np.mean(np.abs(plain.adj_op(plain.op(np.ones((1, 1, *shape), np.complex64)) * dens)))
    grid = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    r2 = sum(g ** 2 for g in grid)
    x_true = (np.exp(-3 * r2) + 0.5 * (r2 < 0.3)).astype(np.complex64).reshape(1, 1, *shape)
    op = ndft_full(traj, shape, n_coils=1, density=dens, squeeze_dims=False)
    y = op.op(x_true).astype(np.complex64)

    def nrmse(x):
        return float(np.linalg.norm(np.ravel(x) - x_true.ravel()) / np.linalg.norm(x_true))

    np.random.seed(0)
    lip_w = float(op.get_lipschitz_cst())
    np.random.seed(0)
    lip_u = float(plain.get_lipschitz_cst())
    assert lip_u > 4 * lip_w                                  # the step 1 / lip_w is far too long
    np.random.seed(0)
    _, it_ref = ref_cg(op, y.copy(), max_iter=6, callback=_iterates, progressbar=False)
    op.density = dens
    err_ref = [nrmse(x) for x in it_ref]
    assert err_ref[-1] > 10 * err_ref[0] or not np.isfinite(err_ref[-1])   # the reference diverges
    np.random.seed(0)
    _, it = solvers.cg(TorchFacade(op), y.copy(), max_iter=6, callback=_iterates)
    op.density = dens
    _close(it[:3], it_ref[:3], 1e-3)                          # ... and ours walks the same path
    assert [nrmse(x) for x in it][-1] > 10 * err_ref[0] or not np.isfinite(nrmse(it[-1]))
    # step size of the operator that is iterated on: monotone, better than the start
    np.random.seed(0)
    _, it_ok = solvers.cg(TorchFacade(op), y.copy(), max_iter=10, callback=_iterates, lipschitz_cst=lip_u)
    err_ok = [nrmse(x) for x in it_ok]
    assert all(b <= a * (1 + 1e-6) for a, b in zip(err_ok, err_ok[1:])) and err_ok[-1] < err_ok[0]
