"""GPU tests of the coil classes of the row kernels (rows_common.cuh): a call with T coils runs in the
smallest class TC >= T, where a warp's 32 lanes are TC coils x 32 / TC row groups of a larger tile, so
that the cost of a transform follows its coil count -- the reference's cost model, a coil loop
(src/mrinufft/operators/base.py:980-1010).  Every class must meet the same bar as class 32:
rel-L2 <= 5e-6 against the reference's exact NDFT (goldens), <= 2e-6 against the float64 oracle."""

import numpy as np
import pytest

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL_NDFT = 5e-6
CLASSES = {3: (1, 2, 4, 8, 16, 32), 2: (8, 16, 32)}


@pytest.fixture(scope="module")
def mods():
    import torch

    import mrinufft
    import mrinufft_b200

    assert mrinufft_b200.MRIB200NUFFT.available, "libb200nufft.so missing or no GPU"
    return mrinufft, mrinufft_b200, torch


def _c(rng, *s):
    return (rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)


def _classes_for(dim, T):
    return [c for c in CLASSES[dim] if c >= T]


GOLDEN_CLASS_CASES = [(case, c) for case, dim, T in [("random3D", 3, 1), ("cones3D", 3, 1), ("random3D_sense", 3, 2),
                                                     ("random2D_sense", 2, 3), ("spiral2D_sense", 2, 4),
                                                     ("nyquist_radial2D", 2, 2)]
                      for c in _classes_for(dim, T)]


@pytest.mark.parametrize("case,cls", GOLDEN_CLASS_CASES)
def test_every_coil_class_matches_the_reference_goldens(mods, case, cls):
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=g["n_coils"], smaps=g.get("smaps"),
                                       squeeze_dims=False)
    plan = op.raw_op.plan
    plan.set_option(0, 2)
    plan.set_option(1, 2)
    plan.set_option(4, cls)
    assert plan.rows_class(g["n_coils"])["class"] == cls
    assert rel_l2(op.op(g["img"]), g["op"]) <= TOL_NDFT
    assert rel_l2(op.adj_op(g["ksp"]), g["adj"]) <= TOL_NDFT
    assert rel_l2(op.data_consistency(g["img"], g["ksp"]), g["dc"]) <= TOL_NDFT
    info = plan.rows_class(g["n_coils"])
    assert info["visits"] > 0 and not info["unsupported"]


@pytest.mark.parametrize("C,want", [(1, 1), (2, 2), (3, 4), (4, 4), (7, 8), (8, 8), (13, 16), (16, 16), (17, 32)])
def test_the_class_follows_the_coil_count_3d(mods, C, want):
    """Natural selection (no option set): class = next power of two >= coils; against the float64 oracle."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(20 + C)
    shape, M = (24, 32, 20), 3000
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(np.float32)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False)
    assert op.raw_op.plan.rows_class(C)["class"] == want
    img, ksp = _c(rng, 1, C, *shape), _c(rng, 1, C, M)
    cpu = CpuNufft(samples, shape, precision="f64")
    y, x = op.op(img), op.adj_op(ksp)
    assert rel_l2(y[0], cpu.op(img[0])) <= 2e-6
    assert rel_l2(x[0], cpu.adj_op(ksp[0])) <= 2e-6


def test_visits_per_point_fall_with_the_class(mods):
    """The point of the classes: 28 (point, tile) visits per sample at class 32 (w = 7, 3-D), about 3 at
    class 1.  Ratios of the stream sizes against class 32, with some slack for crossing points."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(1)
    shape, M = (64, 64, 64), 200_000
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(np.float32)
    op = mrinufft.get_operator("b200")(samples, shape, squeeze_dims=False)
    plan = op.raw_op.plan
    img = _c(rng, 1, 1, *shape)
    visits = {}
    for cls in CLASSES[3]:
        plan.set_option(4, cls)
        op.op(img)
        visits[cls] = plan.rows_class(1)["visits"]
    base = visits[32]
    assert 27 * M <= base <= 40 * M
    for cls, bound in [(16, 0.60), (8, 0.38), (4, 0.24), (2, 0.17), (1, 0.12)]:
        assert visits[cls] <= bound * base, (cls, visits[cls] / base)


@pytest.mark.parametrize("shape", [(64, 64, 64), (32, 64, 128)])
@pytest.mark.parametrize("C", [1, 2, 4, 8, 16])
def test_classes_with_the_fused_fft_and_untouched_tiles(mods, shape, C):
    """Power-of-two grids: the spreader leaves tiles without visitors unwritten (the FFT substitutes zeros)
    and the type-2 FFT leaves them unwritten for the interpolator -- in units of class-32 tiles.  The larger
    tiles of the smaller classes must honour both.  Samples confined to a ball: most of the grid is empty.
    Stale workspace contents are poisoned with NaNs first.  Reference: the same operator in class 32 and
    with the point-driven kernels, SENSE, data consistency."""
    mrinufft, _, torch = mods
    rng = np.random.default_rng(30 + C)
    M = 60_000
    v = rng.standard_normal((M, 3))
    v *= (rng.uniform(0, 1, (M, 1)) ** (1 / 3)) * 0.9 / np.linalg.norm(v, axis=1, keepdims=True)
    samples = v.astype(np.float32)  # radians, |k| <= 0.9 of pi
    smaps = _c(rng, C, *shape)
    smaps /= np.linalg.norm(smaps, axis=0)
    img, ksp = _c(rng, 1, 1, *shape), _c(rng, 1, C, M)
    res = {}
    for name, opts in [("cls", {0: 2, 1: 2}), ("c32", {0: 2, 1: 2, 4: 32}), ("pd", {0: 1, 1: 1})]:
        op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
        plan = op.raw_op.plan
        for k, val in opts.items():
            plan.set_option(k, val)
        if name == "cls":
            assert plan.rows_class(C)["class"] == C
        # poison what the allocator may hand to the plan's lazily allocated buffers and run twice
        junk = torch.full((1 << 22,), float("nan"), device="cuda")
        del junk
        res[name] = [op.op(img), op.adj_op(ksp), op.data_consistency(img, ksp)]
        again = [op.op(img), op.adj_op(ksp)]  # (tiles cut by a chunk boundary are merged with atomics)
        assert rel_l2(again[0], res[name][0]) <= 1e-6 and rel_l2(again[1], res[name][1]) <= 1e-6
    for a, b, c in zip(res["cls"], res["c32"], res["pd"]):
        assert np.all(np.isfinite(a))
        assert rel_l2(a, b) <= 1e-6 and rel_l2(a, c) <= 1e-6


@pytest.mark.parametrize("shape", [(13, 21, 24), (24, 18, 24), (20, 22, 26), (40, 44), (74, 30)])
def test_classes_on_grids_with_partial_tiles_and_wrapped_footprints(mods, shape):
    """Fine grids that are not multiples of the tile extents (short last tile along y / z / x), samples near
    the periodic seam: every class against the float64 oracle."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(4)
    d, M = len(shape), 3000
    samples = rng.uniform(-np.pi, np.pi, (M, d)).astype(np.float32)
    samples[:50] = np.float32(np.pi) - rng.uniform(0, 0.2, (50, d)).astype(np.float32)
    samples[50:100] = -np.float32(np.pi) + rng.uniform(0, 0.2, (50, d)).astype(np.float32)
    op = mrinufft.get_operator("b200")(samples, shape, squeeze_dims=False)
    plan = op.raw_op.plan
    cpu = CpuNufft(samples, shape, precision="f64")
    img, ksp = _c(rng, 1, 1, *shape), _c(rng, 1, 1, M)
    y_o, x_o = cpu.op(img[0, 0]), cpu.adj_op(ksp[0])
    plan.set_option(0, 2)
    plan.set_option(1, 2)
    ran = []
    for cls in CLASSES[d]:
        plan.set_option(4, cls)
        got = plan.rows_class(1)["class"]
        if got == 0:
            pytest.skip(f"fine grid {plan.nf} is not served by the row kernels")
        ran.append(got)  # a class whose tile does not fit the wrap is replaced by a larger one
        assert rel_l2(op.op(img)[0], y_o) <= 2e-6, cls
        assert rel_l2(op.adj_op(ksp)[0], x_o) <= 2e-6, cls
    assert 32 in ran


@pytest.mark.parametrize("cls", [1, 2, 4, 8, 16, 32])
def test_classes_with_a_dense_centre_split_over_chunks(mods, cls):
    """Thousands of coincident samples at k = 0 (radial centre): their tiles are cut by chunk boundaries and
    merged with red.add on pre-zeroed rows; long ranges take the builder's whole-warp path."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(6)
    shape, M = (32, 32, 32), 30_000
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(np.float32)
    samples[:12_000] = 0.0
    samples[12_000:14_000] = rng.normal(0, 0.02, (2000, 3)).astype(np.float32)
    C = min(cls, 3)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False)
    plan = op.raw_op.plan
    plan.set_option(4, cls)
    img, ksp = _c(rng, 1, C, *shape), _c(rng, 1, C, M)
    cpu = CpuNufft(samples, shape, precision="f64")
    assert rel_l2(op.op(img)[0], cpu.op(img[0])) <= 2e-6
    assert rel_l2(op.adj_op(ksp)[0], cpu.adj_op(ksp[0])) <= 2e-6
    assert plan.rows_class(C)["class"] == cls


def test_one_plan_serves_calls_of_different_coil_counts(mods):
    """cfg-D's pattern: 16-coil transforms and the single-coil power method on the same plan, interleaved;
    every class keeps its own stream, `update_samples` invalidates all of them."""
    mrinufft, _, torch = mods
    rng = np.random.default_rng(8)
    shape, M, C = (32, 32, 32), 20_000, 16
    samples = rng.uniform(-0.5, 0.5, (M, 3)).astype(np.float32)
    smaps = _c(rng, C, *shape)
    smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    plan = op.raw_op.plan
    # device arrays: all 16 coils in one call (host arrays run as chunks of 8 coils, see `_chunks`)
    img, ksp = torch.from_numpy(_c(rng, 1, 1, *shape)).cuda(), torch.from_numpy(_c(rng, 1, C, M)).cuda()
    y0 = op.op(img).cpu().numpy()
    np.random.seed(0)
    lip0 = op.get_lipschitz_cst(max_iter=5)
    assert plan.rows_class(16)["visits"] > 0 and plan.rows_class(1)["visits"] > 0
    assert rel_l2(op.op(img).cpu().numpy(), y0) <= 1e-6  # (the interpolator adds its partial sums with atomics)
    moved = (samples + rng.uniform(-0.01, 0.01, samples.shape)).astype(np.float32)
    op.samples = moved
    assert plan.rows_class(16)["visits"] == 0 and plan.rows_class(1)["visits"] == 0
    fresh = mrinufft.get_operator("b200")(moved, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    assert rel_l2(op.op(img).cpu().numpy(), fresh.op(img).cpu().numpy()) <= 1e-6
    assert rel_l2(op.adj_op(ksp).cpu().numpy(), fresh.adj_op(ksp).cpu().numpy()) <= 1e-6
    # host arrays take the same plan through chunks of 8 coils: a third class on the same points
    assert rel_l2(op.op(img.cpu().numpy()), fresh.op(img).cpu().numpy()) <= 1e-6
    assert op.raw_op.plan.rows_class(8)["visits"] > 0
    np.random.seed(0)
    lip1 = fresh.get_lipschitz_cst(max_iter=5)
    np.random.seed(0)
    assert abs(op.get_lipschitz_cst(max_iter=5) - lip1) <= 1e-5 * lip1
    assert lip0 > 0


@pytest.mark.parametrize("eps,w", [(1e-3, 4), (1e-4, 5), (1e-5, 6)])
@pytest.mark.parametrize("shape,C,classes", [((24, 32, 20), 2, (2, 4, 8, 16, 32)), ((24, 32, 20), 1, (1,)),
                                             ((48, 40), 2, (8, 16, 32))])
def test_narrower_kernels_in_every_class(mods, eps, w, shape, C, classes):
    """Kernel widths 4, 5, 6 (eps = 1e-3 .. 1e-5) have their own generated visit loops in every coil class:
    against the float64 oracle run with the same eps (same kernel, same grid: float rounding and the device's
    polynomial fit of the kernel differ)
    and against the exact NDFT at the accuracy the width promises."""
    from oracle import es_nufft as E
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(int(-np.log10(eps)))
    d, M = len(shape), 2500
    samples = rng.uniform(-np.pi, np.pi, (M, d)).astype(np.float32)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False, eps=eps)
    plan = op.raw_op.plan
    assert plan.w == w
    plan.set_option(0, 2)
    plan.set_option(1, 2)
    img, ksp = _c(rng, 1, C, *shape), _c(rng, 1, C, M)
    cpu = CpuNufft(samples, shape, eps=eps, precision="f64")
    y_o, x_o = cpu.op(img[0]), cpu.adj_op(ksp[0])
    A = E.ndft_matrix(samples, shape) / op.norm_factor
    y_n = np.stack([A @ img[0, c].ravel() for c in range(C)])
    for cls in classes:
        plan.set_option(4, cls)
        assert plan.rows_class(C)["class"] == cls
        y, x = op.op(img)[0], op.adj_op(ksp)[0]
        # (the device evaluates the kernel with a polynomial fitted to 0.1 eps, the oracle exactly)
        tol = 3e-6 + 0.05 * eps
        assert rel_l2(y, y_o) <= tol and rel_l2(x, x_o) <= tol, cls
        assert rel_l2(y, y_n) <= 5 * eps, cls
