"""GPU tests of the tile-owned complex128 spreader (csrc/double_rows.cu): `precision="double"` plans spread
without atomics, a warp owning 2 rows x 16 cells x up to 16 coils of the oversampled grid.  The reference computes
in the precision of the samples (src/mrinufft/operators/base.py:934); the bar is the exact NDFT to the accuracy
double allows, and the point-driven double spreader (option 3, bit 5) to rounding."""

import numpy as np
import pytest

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch

    import mrinufft
    import mrinufft_b200

    assert mrinufft_b200.MRIB200NUFFT.available, "libb200nufft.so missing or no GPU"
    return mrinufft, mrinufft_b200, torch


def _c(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def _pair(mrinufft, samples, shape, C, eps, **kw):
    """(operator on the row spreader, operator on the point-driven spreader)"""
    ops = []
    for dbg in (0, 32):
        op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False, eps=eps,
                                           precision="double", **kw)
        if dbg:
            op.raw_op.plan.set_option(3, dbg)
            op.raw_op._set_pts(op.samples)
        ops.append(op)
    assert ops[0].raw_op.plan.rows_class(C)["class"] == 16, "the row spreader does not serve this plan"
    assert ops[1].raw_op.plan.rows_class(C)["class"] == 0
    return ops


@pytest.mark.parametrize("shape,C,eps", [((24, 40), 1, 1e-6), ((32, 32), 5, 1e-9), ((20, 40), 16, 1e-12),
                                         ((64, 44), 19, 1e-4), ((16, 16, 20), 3, 1e-6), ((12, 20, 24), 16, 1e-9),
                                         ((16, 24, 24), 33, 1e-13), ((32, 32, 32), 8, 1e-3), ((4, 4), 2, 1e-12),
                                         ((4, 6, 4), 1, 1e-4)])
def test_row_spreader_equals_point_spreader(mods, shape, C, eps):
    """Every width (w = 4 .. 14), coil counts below, at and above the 16 coils of a launch, grids whose last
    x-tile is short, 2-D and 3-D: the adjoint of the two spreaders agrees to rounding (double atomics add in
    another order), and so does data_consistency."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(hash((shape, C)) % 2**31)
    M = 4000
    samples = rng.uniform(-0.5, 0.5, (M, len(shape)))
    samples[:50] = 0.5 - 1e-9 * rng.random((50, len(shape)))  # the periodic seam
    samples[50:100] = -0.5
    rows, pts = _pair(mrinufft, samples, shape, C, eps)
    ksp = _c(rng, 1, C, M)
    img = _c(rng, 1, C, *shape)
    a, b = rows.adj_op(ksp), pts.adj_op(ksp)
    assert a.dtype == np.complex128 and rel_l2(a, b) <= 1e-13
    assert rel_l2(rows.data_consistency(img, ksp), pts.data_consistency(img, ksp)) <= 1e-13
    # adjointness of the pair (interpolation is point-driven in both)
    y = rows.op(img)
    lhs, rhs = np.vdot(y.ravel(), ksp.ravel()), np.vdot(img.ravel(), a.ravel())
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)


@pytest.mark.parametrize("case,eps,tol", [("random2D_sense", 1e-11, 2e-10), ("random3D_sense", 1e-8, 2e-7),
                                          ("cones3D", 1e-12, 2e-11), ("nyquist_radial2D", 1e-12, 2e-11)])
def test_row_spreader_matches_reference_ndft(mods, case, eps, tol):
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = mrinufft.get_operator("b200")(g["samples"].astype(np.float64), g["shape"], n_coils=g["n_coils"],
                                       smaps=g.get("smaps"), squeeze_dims=False, eps=eps, precision="double")
    assert op.raw_op.plan.rows_class(g["n_coils"])["class"] == 16
    ref = mrinufft.get_operator("numpy")(g["samples"].astype(np.float64), g["shape"], n_coils=g["n_coils"],
                                         smaps=g.get("smaps"))
    ref.squeeze_dims = False
    ksp = g["ksp"].astype(np.complex128)
    assert rel_l2(op.adj_op(ksp), ref.adj_op(ksp)) <= tol


def test_dense_centre_split_tiles_and_update_samples(mods):
    """Half of the points in one cell (ranges longer than a chunk: tiles split over many chunks are merged with
    double atomics on pre-zeroed rows), then new trajectories of other sizes through the same plan, down to
    none at all."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(5)
    shape, C, M = (24, 24, 24), 4, 60000
    samples = rng.uniform(-0.5, 0.5, (M, 3))
    samples[: M // 2] = 0.01 + 1e-4 * rng.standard_normal((M // 2, 3))
    rows, pts = _pair(mrinufft, samples, shape, C, 1e-8)
    ksp = _c(rng, 1, C, M)
    assert rel_l2(rows.adj_op(ksp), pts.adj_op(ksp)) <= 1e-13
    for m in (M + 5000, 300, 1):
        s2 = rng.uniform(-0.5, 0.5, (m, 3))
        k2 = _c(rng, 1, C, m)
        for op in (rows, pts):
            op.samples = s2
        assert rows.raw_op.plan.rows_class(C)["class"] == 16
        assert rel_l2(rows.adj_op(k2), pts.adj_op(k2)) <= 1e-13
    ref = mrinufft.get_operator("numpy")(s2, shape, n_coils=C)
    ref.squeeze_dims = False
    assert rel_l2(rows.adj_op(k2), ref.adj_op(k2)) <= 1e-7


def test_density_and_smaps_ride_along(mods):
    mrinufft, _, _ = mods
    rng = np.random.default_rng(9)
    shape, C, M = (32, 24), 6, 3000
    samples = rng.uniform(-0.5, 0.5, (M, 2))
    smaps = _c(rng, C, *shape)
    dens = rng.uniform(0.5, 2.0, M)
    rows, pts = _pair(mrinufft, samples, shape, C, 1e-10, smaps=smaps, density=dens)
    ksp = _c(rng, 1, C, M)
    a = rows.adj_op(ksp)
    assert a.shape == (1, 1, *shape) and rel_l2(a, pts.adj_op(ksp)) <= 1e-13
    ref = mrinufft.get_operator("numpy")(samples, shape, n_coils=C, smaps=smaps)
    ref.squeeze_dims = False
    assert rel_l2(a, ref.adj_op(ksp * dens)) <= 2e-9


@pytest.mark.parametrize("shape", [(20, 36)])
def test_grids_the_row_spreader_does_not_take_stay_point_driven(mods, shape):
    """Grids whose short last x-tile (72 = 4 * 16 + 8 cells) is narrower than the kernel (w = 13): a footprint
    could touch three tiles.  A plan avoids such a grid (it takes 80 cells) unless finufft's exact grid is asked for."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(3)
    samples = rng.uniform(-0.5, 0.5, (500, 2))
    kw = dict(n_coils=2, squeeze_dims=False, eps=1e-12, precision="double")
    auto = mrinufft.get_operator("b200")(samples, shape, **kw)
    assert tuple(auto.raw_op.plan.nf) == (40, 80) and auto.raw_op.plan.rows_class(2)["class"] == 16
    op = mrinufft.get_operator("b200")(samples, shape, exact_grid=True, **kw)
    assert tuple(op.raw_op.plan.nf) == (40, 72) and op.raw_op.plan.rows_class(2)["class"] == 0
    ksp = _c(rng, 1, 2, 500)
    ref = mrinufft.get_operator("numpy")(samples, shape, n_coils=2)
    ref.squeeze_dims = False
    assert rel_l2(op.adj_op(ksp), ref.adj_op(ksp)) <= 1e-10
    assert rel_l2(auto.adj_op(ksp), ref.adj_op(ksp)) <= 1e-10
