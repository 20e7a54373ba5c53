"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden vectors, the
CPU oracle and the reference's own behavioural contract (tests/operators/*.py of mri-nufft).

Tolerances (BASELINE.json north_star): complex64, eps = 1e-6 -> relative L2 error <= 5e-6 against the
exact NDFT, which implies <= 1e-5 against finufft (itself ~1.2e-6 from the NDFT);
sort / bin indices bit exact.
"""

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL_NDFT = 5e-6


@pytest.fixture(scope="module")
def mods():
    import torch

    import mrinufft
    import mrinufft_b200

    assert mrinufft_b200.MRIB200NUFFT.available, "libb200nufft.so missing or no GPU"
    return mrinufft, mrinufft_b200, torch


def make_op(mrinufft, g, **kw):
    return mrinufft.get_operator("b200")(
        g["samples"], g["shape"], n_coils=g["n_coils"], smaps=g.get("smaps"), squeeze_dims=False, **kw
    )


# ------------------------------------------------------------------ K1: bit-exact sort
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_sort_bit_exact(mods, case):
    from oracle import es_nufft as E

    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    plan = op.raw_op.plan
    ref = E.bin_sort(g["samples"], plan.nf, plan.w, plan.bins)
    origin, x1, key, perm = op.raw_op.sort_indices()
    assert np.array_equal(origin, ref["origin"])
    assert np.array_equal(x1.view(np.uint32), ref["x1"].view(np.uint32))
    assert np.array_equal(key, ref["key"])
    assert np.array_equal(perm, ref["perm"])


def test_sort_bit_exact_adversarial(mods):
    """x = +-pi exactly, +-pi(1 +- 2^-23), 0, denormals, |x| up to 3 pi, cell boundaries."""
    from oracle import es_nufft as E

    mrinufft, _, _ = mods
    pi32 = np.float32(np.pi)
    specials = np.array(
        [0.0, pi32, -pi32, np.nextafter(pi32, np.float32(4)), np.nextafter(pi32, np.float32(0)),
         -np.nextafter(pi32, np.float32(4)), -np.nextafter(pi32, np.float32(0)), 1e-45, -1e-45,
         1e-38, 3 * pi32, -3 * pi32, 2 * pi32, -2 * pi32, 0.5 * pi32, -0.5 * pi32,
         np.float32(2 * np.pi / 128), np.float32(2 * np.pi * 3.5 / 128), np.float32(-2 * np.pi * 3.5 / 128)],
        dtype=np.float32)
    rng = np.random.default_rng(0)
    for shape in [(64, 64), (32, 32, 32)]:
        d = len(shape)
        pts = rng.uniform(-3 * np.pi, 3 * np.pi, (20000, d)).astype(np.float32)
        pts[: len(specials)] = specials[:, None]
        pts[len(specials): 2 * len(specials), 0] = specials
        # points exactly on fine-grid cell boundaries and half cells
        nf = 2 * shape[0]
        cells = (np.arange(300) % nf - nf // 2) * (2 * np.pi / nf)
        pts[100:400, d - 1] = cells.astype(np.float32)
        pts[400:700, 0] = (cells + np.pi / nf).astype(np.float32)
        op = mrinufft.get_operator("b200")(pts, shape)
        assert np.array_equal(op.samples, pts)  # |x| > 0.5 -> taken as radians, not rescaled
        plan = op.raw_op.plan
        ref = E.bin_sort(pts, plan.nf, plan.w, plan.bins)
        origin, x1, key, perm = op.raw_op.sort_indices()
        assert np.array_equal(origin, ref["origin"])
        assert np.array_equal(x1.view(np.uint32), ref["x1"].view(np.uint32))
        assert np.array_equal(key, ref["key"])
        assert np.array_equal(perm, ref["perm"])
        assert np.array_equal(np.sort(perm), np.arange(len(pts)))


# ------------------------------------------------------------------ op / adj_op vs the reference's NDFT goldens
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_op_matches_reference_ndft(mods, case):
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    y = op.op(g["img"])
    assert y.shape == g["op"].shape and y.dtype == np.complex64
    assert rel_l2(y, g["op"]) <= TOL_NDFT


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_adj_op_matches_reference_ndft(mods, case):
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    x = op.adj_op(g["ksp"])
    assert x.shape == g["adj"].shape and x.dtype == np.complex64
    assert rel_l2(x, g["adj"]) <= TOL_NDFT


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_density_and_data_consistency_match_reference(mods, case):
    mrinufft, _, _ = mods
    g = load_golden(case)
    opd = make_op(mrinufft, g, density=g["density"])
    assert rel_l2(opd.adj_op(g["ksp"]), g["adj_density"]) <= TOL_NDFT
    op = make_op(mrinufft, g)
    dc = op.data_consistency(g["img"], g["ksp"])
    assert dc.shape == g["dc"].shape
    assert rel_l2(dc, g["dc"]) <= TOL_NDFT
    # data_consistency == adj_op(op(x) - y)   (tests/operators/test_batch.py:166-199)
    naive = op.adj_op(op.op(g["img"]) - g["ksp"])
    assert rel_l2(dc, naive) <= 2e-6


def test_known_answer_cartesian_grid(mods):
    """NDFT on a Cartesian grid == fftn(fftshift(img)) (reference tests/test_ndft.py:58-79)."""
    import scipy.fft as sfft

    mrinufft, _, _ = mods
    g = load_golden("grid2D")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"])
    img = g["img"][0, 0]
    y = op.op(img) * op.norm_factor
    ref = sfft.fftshift(sfft.fftn(sfft.fftshift(img.astype(np.complex128))))
    assert rel_l2(y.reshape(g["shape"]), ref) <= TOL_NDFT


@pytest.mark.parametrize("case", ["random2D_sense", "random3D", "cones3D"])
def test_matches_cpu_oracle_tightly(mods, case):
    """Same kernel family / parameters as the finufft restatement -> far closer to it than either
    is to the NDFT (float32 rounding only)."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    cpu = CpuNufft(g["samples"], g["shape"], eps=1e-6, precision="f64")
    smaps = g.get("smaps")
    if smaps is not None:
        y_ref = cpu.op(g["img"][0, 0], smaps)
    else:
        y_ref = np.stack([cpu.op(g["img"][0, c])[0] for c in range(g["n_coils"])])
    y = op.op(g["img"])
    assert rel_l2(y, y_ref.reshape(y.shape)) <= 1.5e-6


# ------------------------------------------------------------------ adjointness, Lipschitz (test_interfaces.py:138-176)
@pytest.mark.parametrize("case", ["random2D", "random3D_sense", "spiral2D_sense", "cones3D"])
def test_adjointness(mods, case):
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    rng = np.random.default_rng(3)
    errs = []
    for _ in range(5):
        x = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(np.complex64)
        y = (rng.standard_normal(op.ksp_full_shape) + 1j * rng.standard_normal(op.ksp_full_shape)).astype(np.complex64)
        lhs = np.vdot(op.op(x).astype(np.complex128), y.astype(np.complex128))
        rhs = np.vdot(x.astype(np.complex128), op.adj_op(y).astype(np.complex128))
        errs.append(abs(lhs - rhs) / abs(rhs))
    assert np.mean(errs) < 5e-5
    assert np.mean(errs) < 5e-6  # spread and interp share the sorted points and weights


def test_lipschitz(mods):
    mrinufft, _, _ = mods
    g = load_golden("random2D")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"])
    np.random.seed(0)
    L = op.get_lipschitz_cst(max_iter=30)
    assert isinstance(L, np.floating)
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = (rng.standard_normal(g["shape"]) + 1j * rng.standard_normal(g["shape"])).astype(np.complex64)
        assert np.linalg.norm(op.adj_op(op.op(x))) <= 1.05 * L * np.linalg.norm(x)


# ------------------------------------------------------------------ shapes / squeeze / errors (base.py:225-273)
def test_shapes_squeeze_and_errors(mods):
    mrinufft, _, _ = mods
    g = load_golden("random2D")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"])
    img = g["img"][0, 0]
    y = op.op(img)
    assert y.shape == (1000,)
    assert op.adj_op(y).shape == g["shape"]
    with pytest.raises(ValueError):
        op.op(img[:-1])
    with pytest.raises(ValueError):
        op.adj_op(y[:-1])
    with pytest.raises(ValueError):
        mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=3, n_trans=2)
    with pytest.raises(ValueError):
        op.density = np.ones(3, np.float32)
    assert mrinufft.get_operator("b200").__name__ == "MRIB200NUFFT"
    assert abs(op.norm_factor - np.sqrt(64 * 128 * 4)) < 1e-9


@pytest.mark.parametrize("n_batchs,n_coils,n_trans,sense", [(1, 1, 1, False), (3, 1, 1, False),
                                                             (1, 4, 2, True), (2, 4, 1, True),
                                                             (2, 4, 4, False), (3, 2, 2, False),
                                                             (2, 6, 3, True)])
def test_batched_equals_flat(mods, n_batchs, n_coils, n_trans, sense):
    """Batched op/adj == per (batch, coil) single transforms (tests/operators/test_batch.py:106-163)."""
    mrinufft, _, _ = mods
    g = load_golden("nyquist_radial2D")
    rng = np.random.default_rng(5)
    shape = g["shape"]
    smaps = None
    if sense:
        smaps = (rng.standard_normal((n_coils, *shape)) + 1j * rng.standard_normal((n_coils, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(g["samples"], shape, n_coils=n_coils, n_batchs=n_batchs,
                                       n_trans=n_trans, smaps=smaps, squeeze_dims=False)
    flat = mrinufft.get_operator("b200")(g["samples"], shape, squeeze_dims=True)
    img = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(np.complex64)
    ksp = (rng.standard_normal(op.ksp_full_shape) + 1j * rng.standard_normal(op.ksp_full_shape)).astype(np.complex64)
    y = op.op(img)
    x = op.adj_op(ksp)
    y_ref = np.zeros_like(y)
    x_ref = np.zeros_like(x)
    for b in range(n_batchs):
        for c in range(n_coils):
            if sense:
                y_ref[b, c] = flat.op(img[b, 0] * smaps[c])
                x_ref[b, 0] += np.conj(smaps[c]) * flat.adj_op(ksp[b, c])
            else:
                y_ref[b, c] = flat.op(img[b, c])
                x_ref[b, c] = flat.adj_op(ksp[b, c])
    assert rel_l2(y, y_ref) < 1e-6
    assert rel_l2(x, x_ref) < 1e-6


def test_inputs_are_read_only_safe(mods):
    """Inputs are never mutated and read-only arrays are accepted (test_batch.py:202-210)."""
    mrinufft, _, _ = mods
    g = load_golden("random2D_sense")
    op = make_op(mrinufft, g)
    img, ksp = g["img"].copy(), g["ksp"].copy()
    img.setflags(write=False)
    ksp.setflags(write=False)
    op.op(img)
    op.adj_op(ksp)
    op.data_consistency(img, ksp)
    assert np.array_equal(img, g["img"]) and np.array_equal(ksp, g["ksp"])


def test_array_types_in_same_type_out(mods):
    mrinufft, _, torch = mods
    g = load_golden("random2D_sense")
    op = make_op(mrinufft, g)
    y_np = op.op(g["img"])
    y_tc = op.op(torch.from_numpy(g["img"]))
    y_tg = op.op(torch.from_numpy(g["img"]).cuda())
    assert isinstance(y_np, np.ndarray)
    assert isinstance(y_tc, torch.Tensor) and y_tc.device.type == "cpu"
    assert isinstance(y_tg, torch.Tensor) and y_tg.device.type == "cuda"
    # same arithmetic whatever the container type (the interpolation epilogue accumulates with
    # float atomics, so repeated runs may differ in the last bits, never more)
    assert rel_l2(y_tc.numpy(), y_np) < 1e-6 and rel_l2(y_tg.cpu().numpy(), y_np) < 1e-6
    x_tg = op.adj_op(torch.from_numpy(g["ksp"]).cuda())
    assert x_tg.is_cuda and rel_l2(x_tg.cpu().numpy(), g["adj"]) <= TOL_NDFT
    # float64 / complex128 inputs are cast
    y64 = op.op(g["img"].astype(np.complex128))
    assert y64.dtype == np.complex64 and rel_l2(y64, y_np) < 1e-6
    # smaps given as a CUDA tensor
    op2 = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=g["n_coils"],
                                        smaps=torch.from_numpy(g["smaps"]).cuda(), squeeze_dims=False)
    assert rel_l2(op2.op(g["img"]), y_np) < 1e-6


# ------------------------------------------------------------------ updates (tests/operators/test_update.py:131-248)
def test_update_samples_density_smaps(mods):
    mrinufft, _, _ = mods
    g = load_golden("random2D_sense")
    rng = np.random.default_rng(9)
    op = make_op(mrinufft, g)
    new_samples = g["samples"] + rng.uniform(-0.05, 0.05, g["samples"].shape).astype(np.float32)
    op.samples = new_samples
    fresh = mrinufft.get_operator("b200")(new_samples, g["shape"], n_coils=g["n_coils"],
                                          smaps=g["smaps"], squeeze_dims=False)
    assert rel_l2(op.op(g["img"]), fresh.op(g["img"])) < 1e-6
    assert rel_l2(op.adj_op(g["ksp"]), fresh.adj_op(g["ksp"])) < 1e-6
    # in-place jitter on the returned array then re-assignment (test_update.py:139-145)
    s = op.samples
    s += np.float32(0.01)
    op.samples = s
    fresh = mrinufft.get_operator("b200")(s, g["shape"], n_coils=g["n_coils"], smaps=g["smaps"],
                                          squeeze_dims=False)
    assert rel_l2(op.op(g["img"]), fresh.op(g["img"])) < 1e-6
    # density / smaps setters
    op.density = g["density"]
    fresh = mrinufft.get_operator("b200")(s, g["shape"], n_coils=g["n_coils"], smaps=g["smaps"],
                                          squeeze_dims=False, density=g["density"])
    assert rel_l2(op.adj_op(g["ksp"]), fresh.adj_op(g["ksp"])) < 1e-6
    new_smaps = np.ascontiguousarray(g["smaps"][::-1])
    op.smaps = new_smaps
    fresh.smaps = new_smaps
    assert rel_l2(op.op(g["img"]), fresh.op(g["img"])) < 1e-6
    op.density = None
    assert not op.uses_density


# ------------------------------------------------------------------ pipe density (test_density_for_op.py:20-57)
@pytest.mark.parametrize("osf", [1.5, 2])
@pytest.mark.parametrize("dim", [2, 3])
def test_pipe_density(mods, osf, dim):
    mrinufft, _, _ = mods
    from mrinufft.trajectories import initialize_2D_radial, initialize_3D_phyllotaxis_radial

    if dim == 2:
        traj, shape = initialize_2D_radial(128, 16).astype(np.float32), (32, 32)
    else:
        traj, shape = initialize_3D_phyllotaxis_radial(256, 16).astype(np.float32), (32, 32, 32)
    d = mrinufft.get_operator("b200").pipe(traj.reshape(-1, dim), shape, max_iter=10, osf=osf)
    assert d.shape == (traj.reshape(-1, dim).shape[0],) and d.dtype == np.float32
    assert np.all(np.isfinite(d)) and np.all(d > 0)
    r = np.linalg.norm(traj.reshape(-1, dim), axis=-1)
    # density compensation of a radial trajectory grows like r^(d-1) away from the centre and the
    # k-space edge (reference: tests/operators/test_density_for_op.py:20-57, Pearson r within 0.3)
    mask = (r > 0.05) & (r < 0.4)
    corr = np.corrcoef(d[mask], r[mask] ** (dim - 1))[0, 1]
    assert corr > 0.7
    # density="pipe" through the operator constructor (base.py:608-614 -> density/nufft_based.py)
    op = mrinufft.get_operator("b200")(traj, shape, density=True)
    assert op.uses_density and op.density.shape == d.shape
    ref = mrinufft.get_operator("b200").pipe(traj.reshape(-1, dim), shape)
    assert np.allclose(op.density, ref, rtol=1e-5)


def test_pipe_matches_cpu_oracle(mods):
    """d <- d / |G G^H d| with the spread/interp-only ES kernel, against the C oracle."""
    from oracle.c_oracle import CpuNufft
    from oracle import es_nufft as E

    mrinufft, _, _ = mods
    from mrinufft.trajectories import initialize_2D_radial

    traj = initialize_2D_radial(64, 32).astype(np.float32).reshape(-1, 2)
    shape = (48, 48)
    d = mrinufft.get_operator("b200").pipe(traj, shape, max_iter=10, osf=2, normalize=False)
    # oracle: grid = image shape, kernel from eps=1e-6 / sigma=2
    samples = (traj * 2 * np.pi).astype(np.float32)
    cpu = CpuNufft(samples, shape, precision="f64")
    cpu.nfs = shape
    cpu.nf_arr = np.asarray(shape, np.int32)
    for a in range(2):
        from oracle.c_oracle import fold

        cpu.origin[a], cpu.x1[a] = fold(samples[:, a], shape[a], cpu.w)
    cpu.perm = np.argsort(E.make_key(cpu.origin, shape, E.default_bins(2), cpu.w), kind="stable").astype(np.int32)
    dd = np.ones(len(samples))
    norm2 = np.prod(shape) * 4.0
    for _ in range(10):
        dd = dd / np.abs(cpu.interp(cpu.spread(dd.astype(np.complex128)))[0]) * norm2
    assert rel_l2(d, dd) < 2e-5


@pytest.mark.parametrize("sense", [True, False])
@pytest.mark.parametrize("precision", ["single", "double"])
def test_host_arrays_in_several_chunks_are_pipelined_and_equal_the_device_path(mods, sense, precision):
    """numpy in, several library calls per batch (coil_chunk < n_coils, n_batchs > 1): the copy-stream
    pipeline (role of cufinufft's async_transfer, tests/operators/test_cufinufft_async.py:50-95) returns what
    the plain device path returns; inputs may be read-only and are not modified."""
    mrinufft, _, torch = mods
    rng = np.random.default_rng(8)
    shape, M, C, B = (20, 24, 16), 3000, 6, 2
    rdt, cdt = (np.float32, np.complex64) if precision == "single" else (np.float64, np.complex128)
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(rdt)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(cdt)
        smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, n_batchs=B, smaps=smaps, squeeze_dims=False,
                                       coil_chunk=2, precision=precision,
                                       density=rng.uniform(0.5, 1.5, M).astype(rdt))
    assert op._host_pipeline_applies(np.zeros(1)) and len(op._chunks()) == 3
    x = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(cdt)
    y = (rng.standard_normal((B, C, M)) + 1j * rng.standard_normal((B, C, M))).astype(cdt)
    x0, y0 = x.copy(), y.copy()
    x.setflags(write=False)
    y.setflags(write=False)
    ax = op.op(x)                                             # pipelined (numpy in)
    ax_dev = op.op(torch.from_numpy(x0).cuda()).cpu().numpy() # plain device path (torch in)
    assert isinstance(ax, np.ndarray) and ax.shape == (B, C, M) and ax.dtype == cdt
    tol = 2e-6 if precision == "single" else 1e-11   # same kernels; atomics may commit in another order
    assert rel_l2(ax, ax_dev) < tol
    ahy = op.adj_op(y)
    ahy_dev = op.adj_op(torch.from_numpy(y0).cuda()).cpu().numpy()
    assert isinstance(ahy, np.ndarray) and ahy.shape == tuple(op.img_full_shape)
    assert rel_l2(ahy, ahy_dev) < tol
    assert np.array_equal(x, x0) and np.array_equal(y, y0)
    # twice in a row: the rotating buffers and events of the first call do not leak into the second
    assert rel_l2(op.op(x), ax) < tol and rel_l2(op.adj_op(y), ahy) < tol
    # squeeze_dims follows the plain path
    op1 = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, coil_chunk=2, precision=precision)
    xs = x0[0, 0] if sense else x0[0]
    assert op1.op(xs).shape == (C, M) and op1.adj_op(y0[0]).shape == (shape if sense else (C, *shape))


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,dbl", [(1, 1, False), (3, 1001, False), (40, 777, False), (2, 300001, True)])
def test_solver_vector_kernels_against_numpy(mods, B, n, dbl):
    """`b200_vec_*` (csrc/vecops.cu) through the solver context's wrappers against the array expressions of
    optim.py they replace, in float64 numpy: results and the accumulated norms / dot products; more batches
    than one launch carries scalars for (40 > 32), a length that is not a multiple of anything."""
    _, mb, torch = mods
    from mrinufft_b200.solvers import _Ctx

    rng = np.random.default_rng(B * 7 + n)
    cdt, npc = (torch.complex128, np.complex128) if dbl else (torch.complex64, np.complex64)
    ctx = _Ctx.__new__(_Ctx)
    ctx.cdt, ctx.rnp = cdt, (np.float64 if dbl else np.float32)
    ctx.reduce_ksp = ctx.reduce_img = lambda v: v
    vec = lambda: (rng.standard_normal((B, n)) + 1j * rng.standard_normal((B, n))).astype(npc)
    dev = lambda a: torch.from_numpy(a.copy()).cuda()
    sc = lambda: rng.standard_normal(B)
    tol = 1e-12 if dbl else 2e-6
    col = lambda a: np.asarray(a)[:, None]
    nrm = lambda a: np.linalg.norm(a.astype(np.complex128), axis=1)

    x, y, a, b = vec(), vec(), sc(), sc()
    out = dev(x)
    got = ctx.axpby(out, a, out, b, dev(y), norm="ksp")
    want = col(a) * x.astype(np.complex128) + col(b) * y
    assert rel_l2(out.cpu().numpy(), want) <= tol and np.allclose(got, nrm(want), rtol=tol * 10)
    out = dev(x)
    assert ctx.axpby(out, 1 / a, out) is None and rel_l2(out.cpu().numpy(), x / col(a)) <= tol
    assert np.allclose(ctx.inorm(dev(x)), nrm(x), rtol=tol * 10)

    gsq, num, den = ctx.cg_dots(dev(x), dev(y))
    xd, yd = x.astype(np.complex128).ravel(), y.astype(np.complex128).ravel()
    assert np.isclose(gsq, np.vdot(xd, xd).real, rtol=1e-10)
    assert np.isclose(num, np.dot(xd, xd - yd), rtol=1e-9, atol=1e-9 * n * B)
    assert np.isclose(den, np.dot(yd, yd), rtol=1e-9, atol=1e-9 * n * B)

    xi, v, g, beta, L = vec(), vec(), vec(), complex(0.3, -0.2), 1.7
    xt, vt = dev(xi), dev(v)
    ctx.cg_step(xt, vt, dev(g), beta, L)
    vn = g.astype(np.complex128) + beta * v
    assert rel_l2(vt.cpu().numpy(), vn) <= tol and rel_l2(xt.cpu().numpy(), xi - vn / L) <= tol

    xi, w, v, t1, t2 = vec(), vec(), vec(), sc(), sc()
    xt, wt = dev(xi), dev(w)
    got = ctx.lsqr_step(xt, wt, dev(v), t1, t2)
    assert np.allclose(got, nrm(w), rtol=tol * 10)
    assert rel_l2(xt.cpu().numpy(), xi + col(t1) * w) <= tol and rel_l2(wt.cpu().numpy(), v + col(t2) * w) <= tol

    xi, hb, h, v, a, b, c = vec(), vec(), vec(), vec(), sc(), sc(), sc()
    xt, hbt, ht = dev(xi), dev(hb), dev(h)
    got = ctx.lsmr_step(xt, hbt, ht, dev(v), a, b, c)
    hbn = h.astype(np.complex128) + col(a) * hb
    xn = xi + col(b) * hbn
    assert rel_l2(hbt.cpu().numpy(), hbn) <= tol and rel_l2(xt.cpu().numpy(), xn) <= tol
    assert rel_l2(ht.cpu().numpy(), v + col(c) * h) <= tol and np.allclose(got, nrm(xn), rtol=tol * 10)


# ------------------------------------------------------------------ solvers
def test_cg_matches_reference_golden(mods):
    mrinufft, _, _ = mods
    g = load_golden("cg2D_sense")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=4, smaps=g["smaps"])
    np.random.seed(1234)
    L = op.get_lipschitz_cst()
    assert abs(L - g["lipschitz"]) / g["lipschitz"] < 1e-4
    np.random.seed(1234)
    x = op.pinv_solver(g["y"], optim="cg", max_iter=10)
    assert x.shape == g["shape"]
    assert rel_l2(x, g["x_cg"]) < 1e-4


@pytest.mark.parametrize("optim", ["cg", "lsqr", "lsmr"])
def test_pinv_solver_residual_decreases(mods, optim):
    """tests/operators/test_optim.py:60-68 and test_batch.py:255-272."""
    mrinufft, _, _ = mods
    from mrinufft.extras.optim import loss_l2_reg

    g = load_golden("cg2D_sense")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=4, smaps=g["smaps"])
    x, res = op.pinv_solver(g["y"], optim=optim, max_iter=5, callback=loss_l2_reg, progressbar=False)
    assert x.shape == g["shape"]
    res = np.asarray(res).ravel()
    assert res[-1] <= res[0]


def _collect_iterates(image, operator, kspace_data, damp=0.0, x0=None):
    return np.array(image, copy=True).reshape(operator.img_full_shape)


@pytest.mark.parametrize("case", ["solvers2D_sense", "solvers2D_batch_density", "solvers3D_sense_damp"])
@pytest.mark.parametrize("optim", ["lsqr", "lsmr", "cg"])
def test_device_solvers_replay_reference_iterates(mods, case, optim):
    """The device-resident solvers on the CUDA operator against the iterates of the reference's own
    solvers on its exact NDFT (tests/golden/make_solver_golden.py), iteration by iteration."""
    mrinufft, _, _ = mods
    g = load_golden(case)
    dens = g["density"] if "density" in g else False
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=g["n_coils"], n_batchs=int(g["n_batchs"]),
                                       smaps=g.get("smaps"), density=dens)
    y = g["y"].copy()
    np.random.seed(99)  # start of the power method behind cg's step size (base.py:1194)
    x, its = op.pinv_solver(y, optim=optim, damp=float(g["damp"]), max_iter=8, callback=_collect_iterates,
                            progressbar=False)
    assert np.array_equal(y, g["y"])
    ref = g[f"it_{optim}"]
    assert len(its) == len(ref)
    for k in range(len(ref)):
        assert rel_l2(its[k], ref[k]) < 1e-3, (optim, k, rel_l2(its[k], ref[k]))
    assert rel_l2(x.reshape(ref[-1].shape), ref[-1]) < 1e-3
    if dens is not False:  # the density is switched off during the iteration and restored afterwards
        assert op.uses_density and np.allclose(op.density, g["density"])


@pytest.mark.parametrize("optim", ["lsqr", "lsmr"])
def test_device_solvers_equal_reference_solvers_on_the_same_operator(mods, optim):
    """``pinv_solver`` (device resident) == the reference's solver driving the same b200 operator through
    host arrays (extras/optim.py via ``with_numpy_cupy``); torch CUDA in -> torch CUDA out."""
    mrinufft, _, torch = mods
    from mrinufft.extras import get_optimizer

    g = load_golden("solvers2D_sense")
    op = mrinufft.get_operator("b200")(g["samples"], g["shape"], n_coils=g["n_coils"], smaps=g["smaps"])
    want = get_optimizer(optim)(operator=op, kspace_data=g["y"].copy(), max_iter=6, progressbar=False)
    got = op.pinv_solver(g["y"], optim=optim, max_iter=6)
    assert isinstance(got, np.ndarray) and got.shape == want.shape
    assert rel_l2(got, want) < 2e-4
    yt = torch.from_numpy(g["y"]).cuda()
    got_t = op.pinv_solver(yt, optim=optim, max_iter=6)
    assert torch.is_tensor(got_t) and got_t.is_cuda
    assert rel_l2(got_t.cpu().numpy(), got) < 1e-5


@pytest.mark.parametrize("window_fun", ["ellipse", "rect"])
def test_low_frequency_smaps_on_device(mods, window_fun):
    """``smaps={"name": "low_frequency", ...}`` (extras/smaps.py:220-306, default mask / blur): same steps
    with the reference's centre extraction + its lsqr on the exact NDFT."""
    mrinufft, _, torch = mods
    from conftest import ndft_full
    from mrinufft.extras.optim import lsqr
    from mrinufft.extras.smaps import _extract_kspace_center
    from mrinufft.trajectories import initialize_2D_radial

    rng = np.random.default_rng(5)
    shape, C = (24, 20), 3
    samples = (initialize_2D_radial(32, 64).reshape(-1, 2) * 2 * np.pi).astype(np.float32)
    ksp = (rng.standard_normal((C, len(samples))) + 1j * rng.standard_normal((C, len(samples)))).astype(np.complex64)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps={
        "name": "low_frequency", "kspace_data": ksp, "threshold": 1.0, "max_iter": 6, "window_fun": window_fun})
    assert op.uses_sense and op.n_coils == C
    assert isinstance(op.smaps, np.ndarray) and op.smaps.shape == (C, *shape)
    k_c, s_c, _ = _extract_kspace_center(kspace_data=ksp, kspace_loc=op.samples, threshold=1.0, density=None,
                                         window_fun=window_fun)
    assert (k_c.shape[-1] < len(samples)) == (window_fun == "rect")  # windows weight, "rect" selects
    ref_op = ndft_full(s_c.astype(np.float64), shape, n_coils=C, squeeze_dims=True)
    maps = lsqr(ref_op, k_c.copy(), max_iter=6, progressbar=False)
    maps = maps / np.linalg.norm(maps, axis=0)
    assert rel_l2(op.smaps, maps) < 1e-3
    assert np.allclose(np.linalg.norm(op.smaps, axis=0), 1.0, atol=1e-5)
    # the maps are live: SENSE op / adj_op run with them
    x = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    y = op.op(x)
    assert y.shape == (C, len(samples))
    want = ndft_full(samples.astype(np.float64), shape, n_coils=C, smaps=op.smaps).op(x)
    assert rel_l2(y, want) < TOL_NDFT


# ------------------------------------------------------------------ autodiff (tests/operators/test_autodiff.py)
def test_autodiff_data_and_trajectory(mods):
    mrinufft, _, torch = mods
    from oracle import es_nufft as E

    rng = np.random.default_rng(2)
    shape = (12, 16)
    M = 200
    samples = rng.uniform(-np.pi, np.pi, (M, 2)).astype(np.float32)
    op = mrinufft.get_operator("b200", wrt_data=True, wrt_traj=True)(samples, shape, squeeze_dims=False)
    norm = np.sqrt(np.prod(shape) * 4.0)
    x = torch.from_numpy((rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)).cuda()
    x.requires_grad_(True)
    ktraj = op.samples
    y = op.op(x)
    # dense NDFT model in torch (float64) for the expected gradients
    st = torch.from_numpy(samples.astype(np.float64)).requires_grad_(True)
    r = torch.stack(torch.meshgrid(*[torch.arange(s, dtype=torch.float64) - s // 2 for s in shape], indexing="ij"), 0).reshape(2, -1)
    A = torch.exp(-1j * (st @ r)) / norm
    xd = x.detach().cpu().to(torch.complex128).reshape(-1).requires_grad_(True)
    yd = A @ xd
    assert rel_l2(y.detach().cpu().numpy().ravel(), yd.detach().numpy()) <= TOL_NDFT
    w = torch.from_numpy((rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64))
    loss = torch.sum(torch.abs(y.reshape(-1) - w.cuda()) ** 2)
    loss.backward()
    lossd = torch.sum(torch.abs(yd - w.to(torch.complex128)) ** 2)
    lossd.backward()
    assert rel_l2(x.grad.cpu().numpy().ravel(), xd.grad.numpy()) < 1e-4
    assert rel_l2(ktraj.grad.numpy(), st.grad.numpy()) < 1e-3
    # adjoint direction
    ktraj.grad = None
    k = torch.from_numpy((rng.standard_normal((1, 1, M)) + 1j * rng.standard_normal((1, 1, M))).astype(np.complex64)).cuda()
    k.requires_grad_(True)
    img = op.adj_op(k)
    st2 = torch.from_numpy(samples.astype(np.float64)).requires_grad_(True)
    A2 = torch.exp(-1j * (st2 @ r)) / norm
    kd = k.detach().cpu().to(torch.complex128).reshape(-1).requires_grad_(True)
    imgd = A2.conj().T @ kd
    assert rel_l2(img.detach().cpu().numpy().ravel(), imgd.detach().numpy()) <= TOL_NDFT
    t = torch.from_numpy((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64))
    torch.sum(torch.abs(img.reshape(shape) - t.cuda()) ** 2).backward()
    torch.sum(torch.abs(imgd.reshape(shape) - t.to(torch.complex128)) ** 2).backward()
    assert rel_l2(k.grad.cpu().numpy().ravel(), kd.grad.numpy()) < 1e-4
    assert rel_l2(ktraj.grad.numpy(), st2.grad.numpy()) < 1e-3
    _ = E


# ------------------------------------------------------------------ full-size properties (sampled exact NDFT)
@pytest.mark.parametrize("shape,M,C", [((256, 256), 1 << 18, 4), ((96, 128, 80), 1 << 19, 3)])
def test_large_sampled_ndft_and_linearity(mods, shape, M, C):
    from oracle import es_nufft as E
    from scipy.stats import truncnorm

    mrinufft, _, torch = mods
    d = len(shape)
    rng = np.random.default_rng(0)
    samples = (truncnorm(-3, 3, 0, 0.16).rvs((M, d), random_state=0) * 2 * np.pi).astype(np.float32)
    smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
    smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    img = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    y = op.op(img[None, None])
    idx = rng.choice(M, 64, replace=False)
    for c in (0, C - 1):
        ref = E.ndft_type2_sampled(samples, img.astype(np.complex128) * smaps[c], idx) / op.norm_factor
        assert rel_l2(y[0, c, idx], ref) <= TOL_NDFT
    ksp = (rng.standard_normal((1, C, M)) + 1j * rng.standard_normal((1, C, M))).astype(np.complex64)
    x = op.adj_op(ksp)
    vox = np.stack([rng.integers(0, s, 24) for s in shape], -1)
    ref = np.zeros(len(vox), np.complex128)
    for c in range(C):
        ref += np.conj(smaps[c][tuple(vox.T)]) * E.ndft_type1_sampled(samples, ksp[0, c], shape, vox)
    ref /= op.norm_factor
    assert rel_l2(x[0, 0][tuple(vox.T)], ref) <= TOL_NDFT
    # linearity + adjointness at full size
    img2 = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    y2 = op.op(img2[None, None])
    y12 = op.op((img + 2j * img2)[None, None])
    assert rel_l2(y12, y + 2j * y2) < 2e-6
    lhs = np.vdot(y.astype(np.complex128), ksp.astype(np.complex128))
    rhs = np.vdot(img.astype(np.complex128), x[0, 0].astype(np.complex128))
    assert abs(lhs - rhs) / abs(rhs) < 5e-6


def _ndft_type2_at(torch, samples_d, img64, idx):
    """Exact type-2 NDFT (float64, on the device, nothing of ours involved) of the images `img64` (C, *shape) at
    the sample locations `idx`: y[c, j] = sum_n img[c, n] exp(-i k_j . (n - N/2)).  One phase vector per
    location, shared by all coils."""
    shape = img64.shape[1:]
    axes = [torch.arange(n, device=img64.device, dtype=torch.float64) - n // 2 for n in shape]
    out = torch.empty((img64.shape[0], len(idx)), dtype=torch.complex128, device=img64.device)
    flat = img64.reshape(img64.shape[0], -1)
    for j, i in enumerate(idx):
        k = samples_d[int(i)].to(torch.float64)
        ph = torch.polar(torch.ones_like(axes[0]), -k[0] * axes[0])
        for a in range(1, len(shape)):
            ph = (ph[..., None] * torch.polar(torch.ones_like(axes[a]), -k[a] * axes[a])).reshape(-1)
        out[:, j] = flat @ ph.reshape(-1)
    return out


def _ndft_type1_at(torch, samples_d, ksp64, shape, vox):
    """Exact type-1 NDFT (float64, on the device) of the k-space rows `ksp64` (C, M) at the voxels `vox`:
    f[c, v] = sum_j ksp[c, j] exp(+i k_j . (n_v - N/2))."""
    s64 = samples_d.to(torch.float64)
    out = torch.empty((ksp64.shape[0], len(vox)), dtype=torch.complex128, device=ksp64.device)
    for v, n in enumerate(vox):
        r = torch.tensor([int(n[a]) - shape[a] // 2 for a in range(len(shape))], dtype=torch.float64, device=s64.device)
        ph = torch.polar(torch.ones(s64.shape[0], dtype=torch.float64, device=s64.device), s64 @ r)
        out[:, v] = ksp64 @ ph
    return out


def test_device_ndft_helpers_agree_with_the_cpu_oracle(mods):
    """The float64 torch evaluations of the exact NDFT used at full size, pinned to the numpy oracle."""
    from oracle import es_nufft as E

    _, _, torch = mods
    rng = np.random.default_rng(2)
    shape, M = (12, 10, 14), 300
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(np.float32)
    img = rng.standard_normal((2, *shape)) + 1j * rng.standard_normal((2, *shape))
    ksp = rng.standard_normal((2, M)) + 1j * rng.standard_normal((2, M))
    idx = np.array([0, 7, 299])
    vox = np.array([[0, 0, 0], [6, 5, 7], [11, 9, 13]])
    sd = torch.from_numpy(samples).cuda()
    y = _ndft_type2_at(torch, sd, torch.from_numpy(img).cuda(), idx).cpu().numpy()
    x = _ndft_type1_at(torch, sd, torch.from_numpy(ksp).cuda(), shape, vox).cpu().numpy()
    for c in range(2):
        assert rel_l2(y[c], E.ndft_type2_sampled(samples, img[c], idx)) < 1e-12
        assert rel_l2(x[c], E.ndft_type1_sampled(samples, ksp[c], shape, vox)) < 1e-12


def test_baseline_config_at_its_full_size(mods):
    """BASELINE.json configs[2] exactly -- 3-D 256^3, 32 coils with smaps, M = 2^23 phyllotaxis-radial samples,
    complex64 -- against the exact NDFT (float64) at sampled locations, at the bar of this file (5e-6):
    `op` at 64 k-space locations for ALL 32 coils, among them samples of the k = 0 crowd (16 384 coincident
    points, whose tiles the spreader / interpolator split over many chunks and merge with atomics); `adj_op`
    with density weights at 24 voxels (every voxel sums all 2^23 samples of all coils);
    `data_consistency` == adj_op(op(x) - y); linearity, adjointness, and the invariants of the sort."""
    mrinufft, _, torch = mods
    if torch.cuda.mem_get_info()[0] < 80e9:
        pytest.skip("needs about 70 GB of device memory")
    from mrinufft.trajectories import initialize_3D_phyllotaxis_radial

    shape, C, M = (256, 256, 256), 32, 1 << 23
    traj = initialize_3D_phyllotaxis_radial(16384, 512).astype(np.float32).reshape(-1, 3)
    gen = torch.Generator(device="cuda").manual_seed(0)

    def crandn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, device="cuda", generator=gen))

    smaps = crandn(C, *shape)
    smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
    dens = torch.rand(M, device="cuda", generator=gen) + 0.5
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False, density=dens)
    assert op.n_samples == M and len(op._chunks()) == 1          # one library call for all 32 coils
    assert op.raw_op.plan.rows_class(C)["class"] == 32
    img, img2, ksp = crandn(1, 1, *shape), crandn(1, 1, *shape), crandn(1, C, M)
    samples_d = torch.from_numpy(op.samples).cuda()
    y = op.op(img)
    # ---- type 2 against the exact NDFT: 64 locations, all coils
    rng = np.random.default_rng(0)
    centre = np.flatnonzero(np.linalg.norm(op.samples, axis=1) < 1e-6)
    assert len(centre) >= 1000                                     # the k = 0 crowd of the radial trajectory
    idx = np.sort(np.concatenate([rng.choice(centre, 8, replace=False), rng.choice(M, 56, replace=False)]))
    coil_imgs = (img[0].to(torch.complex128) * smaps.to(torch.complex128))
    ref = _ndft_type2_at(torch, samples_d, coil_imgs, idx) / op.norm_factor
    del coil_imgs
    got = y[0][:, torch.from_numpy(idx).cuda()].to(torch.complex128)
    for c in range(C):
        assert float(torch.linalg.norm(got[c] - ref[c]) / torch.linalg.norm(ref[c])) <= TOL_NDFT, c
    # ---- type 1 (density weighted, SENSE combined) against the exact NDFT: 24 voxels
    x = op.adj_op(ksp)
    vox = np.stack([rng.integers(0, s, 24) for s in shape], -1)
    vox[0], vox[1], vox[2] = (128, 128, 128), (0, 0, 0), (255, 255, 255)
    f = _ndft_type1_at(torch, samples_d, ksp[0].to(torch.complex128) * dens.to(torch.float64), shape, vox)
    vi = tuple(torch.from_numpy(vox[:, a]).cuda() for a in range(3))
    ref1 = torch.sum(torch.conj(smaps.to(torch.complex128)[(slice(None), *vi)]) * f, dim=0) / op.norm_factor
    got1 = x[0, 0][vi].to(torch.complex128)
    assert float(torch.linalg.norm(got1 - ref1) / torch.linalg.norm(ref1)) <= TOL_NDFT
    del f
    # ---- data consistency: the fused call equals its definition
    dc = op.data_consistency(img, ksp)
    dc -= op.adj_op(y - ksp)
    assert float(torch.linalg.norm(dc) / torch.linalg.norm(x)) < 2e-6
    del dc
    # ---- linearity
    y12 = op.op(img + 2j * img2)
    y12 -= y + 2j * op.op(img2)
    assert float(torch.linalg.norm(y12) / torch.linalg.norm(y)) < 3e-6
    del y12
    # ---- adjointness (density off), inner products in float64 on the device
    op.density = None
    x = op.adj_op(ksp)
    lhs = torch.sum(torch.conj(y.to(torch.complex128)) * ksp.to(torch.complex128))
    rhs = torch.sum(torch.conj(img.to(torch.complex128)) * x.to(torch.complex128))
    assert float(torch.abs(lhs - rhs) / torch.abs(rhs)) < 5e-6
    # the sort (perm = stable argsort of the bin keys): keys ascending along perm, ties in caller order,
    # permutation complete
    _, _, key, perm = op.raw_op.sort_indices()
    steps = np.diff(key[perm].astype(np.int64))
    assert np.all(steps >= 0)
    assert np.all(np.diff(perm.astype(np.int64))[steps == 0] > 0)
    assert np.array_equal(np.sort(perm), np.arange(M, dtype=perm.dtype))


@pytest.mark.parametrize("C", [16, 4])
def test_strong_scaling_shards_of_the_baseline_config_at_full_size(mods, C):
    """What one rank of a 2- / 8-GPU run of BASELINE configs[2] executes: the 256^3 / M = 2^23 transforms with
    16 and 4 coils, i.e. coil classes 16 and 4 of the row kernels, against the exact NDFT at sampled locations
    (all coils) and through adjointness."""
    mrinufft, _, torch = mods
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs about 40 GB of device memory")
    from mrinufft.trajectories import initialize_3D_phyllotaxis_radial

    shape, M = (256, 256, 256), 1 << 23
    traj = initialize_3D_phyllotaxis_radial(16384, 512).astype(np.float32).reshape(-1, 3)
    gen = torch.Generator(device="cuda").manual_seed(C)

    def crandn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, device="cuda", generator=gen))

    smaps = crandn(C, *shape)
    smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    assert op.raw_op.plan.rows_class(C)["class"] == C
    img, ksp = crandn(1, 1, *shape), crandn(1, C, M)
    samples_d = torch.from_numpy(op.samples).cuda()
    rng = np.random.default_rng(C)
    centre = np.flatnonzero(np.linalg.norm(op.samples, axis=1) < 1e-6)
    idx = np.sort(np.concatenate([rng.choice(centre, 4, replace=False), rng.choice(M, 28, replace=False)]))
    y = op.op(img)
    ref = _ndft_type2_at(torch, samples_d, img[0].to(torch.complex128) * smaps.to(torch.complex128), idx) / op.norm_factor
    got = y[0][:, torch.from_numpy(idx).cuda()].to(torch.complex128)
    for c in range(C):
        assert float(torch.linalg.norm(got[c] - ref[c]) / torch.linalg.norm(ref[c])) <= TOL_NDFT, c
    x = op.adj_op(ksp)
    vox = np.stack([rng.integers(0, s, 12) for s in shape], -1)
    f = _ndft_type1_at(torch, samples_d, ksp[0].to(torch.complex128), shape, vox)
    vi = tuple(torch.from_numpy(vox[:, a]).cuda() for a in range(3))
    ref1 = torch.sum(torch.conj(smaps.to(torch.complex128)[(slice(None), *vi)]) * f, dim=0) / op.norm_factor
    got1 = x[0, 0][vi].to(torch.complex128)
    assert float(torch.linalg.norm(got1 - ref1) / torch.linalg.norm(ref1)) <= TOL_NDFT
    lhs = torch.sum(torch.conj(y.to(torch.complex128)) * ksp.to(torch.complex128))
    rhs = torch.sum(torch.conj(img.to(torch.complex128)) * x.to(torch.complex128))
    assert float(torch.abs(lhs - rhs) / torch.abs(rhs)) < 5e-6


def test_secondary_baseline_configs_at_their_sizes(mods):
    """BASELINE configs[0] (README demo: 2-D radial 100 x 500, 512^2, one coil, density="voronoi" through the
    constructor) and configs[1] (2-D spiral 64 x 2048, 320^2, 32 coils with smaps) at their stated sizes against
    the exact NDFT at sampled locations."""
    mrinufft, _, torch = mods
    from mrinufft.trajectories import initialize_2D_radial, initialize_2D_spiral

    rng = np.random.default_rng(5)
    for name, traj, shape, C, kw in [
        ("A", initialize_2D_radial(100, 500), (512, 512), 1, {"density": "voronoi"}),
        ("B", initialize_2D_spiral(64, 2048, nb_revolutions=8), (320, 320), 32, {}),
    ]:
        traj = traj.astype(np.float32).reshape(-1, 2)
        smaps = None
        if C > 1:
            smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
            smaps /= np.linalg.norm(smaps, axis=0)
        op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False, **kw)
        M = op.n_samples
        if name == "A":
            assert op.uses_density and op.density.shape == (M,)
        img = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
        ksp = (rng.standard_normal((1, C, M)) + 1j * rng.standard_normal((1, C, M))).astype(np.complex64)
        sd = torch.from_numpy(op.samples).cuda()
        idx = np.sort(rng.choice(M, 48, replace=False))
        sm = torch.ones((1, *shape), dtype=torch.complex128, device="cuda") if smaps is None else \
            torch.from_numpy(smaps).cuda().to(torch.complex128)
        ref = (_ndft_type2_at(torch, sd, torch.from_numpy(img[0]).cuda().to(torch.complex128) * sm, idx)
               / op.norm_factor).cpu().numpy()
        y = op.op(img)
        for c in range(C):
            assert rel_l2(y[0, c, idx], ref[c]) <= TOL_NDFT, (name, c)
        x = op.adj_op(ksp)
        vox = np.stack([rng.integers(0, s, 16) for s in shape], -1)
        k64 = torch.from_numpy(ksp[0]).cuda().to(torch.complex128)
        if op.uses_density:
            k64 = k64 * torch.from_numpy(np.asarray(op.density)).cuda().to(torch.float64)
        f = _ndft_type1_at(torch, sd, k64, shape, vox)
        vi = tuple(torch.from_numpy(vox[:, a]).cuda() for a in range(2))
        ref1 = (torch.sum(torch.conj(sm[(slice(None), *vi)]) * f, dim=0) / op.norm_factor).cpu().numpy()
        assert rel_l2(x[0, 0][tuple(vox.T)], ref1) <= TOL_NDFT, name


def test_pipe_density_3d_at_128_matches_the_oracle_on_sampled_points(mods):
    """3-D Pipe density at 128^3 (M = 2^19 radial samples): ten spread / interp-only iterations on the device
    (single-coil calls: coil class 1 of the row kernels) against the C oracle's spread / interp, compared on
    the whole weight vector (the iteration is global: every weight depends on all the others)."""
    from oracle import es_nufft as E
    from oracle.c_oracle import CpuNufft, fold

    mrinufft, _, _ = mods
    from mrinufft.trajectories import initialize_3D_phyllotaxis_radial

    shape = (128, 128, 128)
    traj = initialize_3D_phyllotaxis_radial(2048, 256).astype(np.float32).reshape(-1, 3)
    d = mrinufft.get_operator("b200").pipe(traj, shape, max_iter=10, osf=2, normalize=False)
    samples = (traj * 2 * np.pi).astype(np.float32)
    cpu = CpuNufft(samples, (8, 8, 8), precision="f64")   # (tables only: the grid is replaced below)
    cpu.nfs = shape
    cpu.nf_arr = np.asarray(shape, np.int32)
    cpu.origin = np.empty((3, len(samples)), np.int32)
    cpu.x1 = np.empty((3, len(samples)), np.float32)
    for a in range(3):
        cpu.origin[a], cpu.x1[a] = fold(samples[:, a], shape[a], cpu.w)
    cpu.perm = np.argsort(E.make_key(cpu.origin, shape, E.default_bins(3), cpu.w), kind="stable").astype(np.int32)
    dd = np.ones(len(samples))
    norm2 = np.prod(shape) * 8.0
    for _ in range(10):
        dd = dd / np.abs(cpu.interp(cpu.spread(dd.astype(np.complex128)))[0]) * norm2
    assert np.all(np.isfinite(d)) and rel_l2(d, dd) < 5e-5
    idx = np.random.default_rng(0).choice(len(d), 64, replace=False)
    assert np.allclose(d[idx], dd[idx], rtol=2e-4)


def test_empty_and_tiny_inputs(mods):
    mrinufft, _, _ = mods
    rng = np.random.default_rng(0)
    shape = (16, 16)
    one = np.array([[0.3, -1.2]], np.float32)
    op = mrinufft.get_operator("b200")(one, shape)
    img = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    from oracle import es_nufft as E

    A = E.ndft_matrix(one, shape)
    assert rel_l2(op.op(img) * op.norm_factor, A @ img.ravel()) <= TOL_NDFT
    x = op.adj_op(np.array([1 + 2j], np.complex64))
    assert rel_l2(x.ravel() * op.norm_factor, A.conj().T @ np.array([1 + 2j])) <= TOL_NDFT


# ------------------------------------------------------------------ kernel variants
@pytest.mark.parametrize("case", ["random2D_sense", "random3D_sense", "nyquist_radial2D", "cones3D"])
@pytest.mark.parametrize("spread_m,interp_m", [(1, 1), (2, 1), (1, 2), (2, 2)])
def test_kernel_variants_match_goldens(mods, case, spread_m, interp_m):
    """Point-driven (method 1) and row-owned (method 2) spread / interp kernels all meet the bar."""
    mrinufft, _, _ = mods
    g = load_golden(case)
    op = make_op(mrinufft, g)
    op.raw_op.plan.set_option(0, spread_m)
    op.raw_op.plan.set_option(1, interp_m)
    assert rel_l2(op.op(g["img"]), g["op"]) <= TOL_NDFT
    assert rel_l2(op.adj_op(g["ksp"]), g["adj"]) <= TOL_NDFT
    assert rel_l2(op.data_consistency(g["img"], g["ksp"]), g["dc"]) <= TOL_NDFT


def test_row_owned_spread_is_bit_reproducible(mods):
    """No atomics in the row-owned spreader: repeated runs give identical bits."""
    mrinufft, _, _ = mods
    g = load_golden("random3D_sense")
    op = make_op(mrinufft, g)
    op.raw_op.plan.set_option(0, 2)
    a = op.adj_op(g["ksp"])
    b = op.adj_op(g["ksp"])
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(40, 44), (74, 30), (20, 22, 26)])
def test_odd_grid_sizes_wrap(mods, shape):
    """Fine grids that are not multiples of the 32-cell tile (partial last tile, wrapped taps)."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(4)
    d = len(shape)
    M = 3000
    samples = rng.uniform(-np.pi, np.pi, (M, d)).astype(np.float32)
    samples[:50] = np.float32(np.pi) - rng.uniform(0, 0.2, (50, d)).astype(np.float32)  # near the seam
    samples[50:100] = -np.float32(np.pi) + rng.uniform(0, 0.2, (50, d)).astype(np.float32)
    op = mrinufft.get_operator("b200")(samples, shape)
    cpu = CpuNufft(samples, shape, precision="f64")
    img = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    ksp = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    for sm, im in [(1, 1), (2, 2)]:
        op.raw_op.plan.set_option(0, sm)
        op.raw_op.plan.set_option(1, im)
        assert rel_l2(op.op(img), cpu.op(img)[0]) <= 2e-6
        assert rel_l2(op.adj_op(ksp), cpu.adj_op(ksp)[0]) <= 2e-6


# ------------------------------------------------------------------ fused zero-padding-aware FFT passes
@pytest.mark.parametrize("shape,C,sense", [((32, 64), 3, True), ((16, 128), 1, False), ((16, 32, 64), 3, True),
                                           ((64, 16, 32), 2, False), ((128, 128, 128), 2, True),
                                           ((16, 256), 3, True), ((16, 16, 256), 2, False),
                                           ((256, 16), 2, True), ((256, 16, 16), 3, True), ((16, 256, 32), 2, False)])
def test_pruned_fft_matches_cufft_path_and_oracle(mods, shape, C, sense):
    """Power-of-two grids take the fused pad/crop + pruned FFT passes (option key 2 = 2); they must
    agree with the cuFFT + k_pad/k_crop path (key 2 = 1) and with the CPU oracle, both signs, SENSE
    coil sum, accumulate over coil chunks, data consistency."""
    from oracle.c_oracle import CpuNufft

    mrinufft, _, _ = mods
    rng = np.random.default_rng(7)
    d = len(shape)
    M = 20000 if np.prod(shape) > 1e5 else 4000
    samples = rng.uniform(-np.pi, np.pi, (M, d)).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    # coil_chunk=2 with C=3 exercises accumulate=1 on the second chunk
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False,
                                       coil_chunk=2 if C == 3 else None)
    plan = op.raw_op.plan
    assert all(n in (32, 64, 128, 256, 512, 1024) for n in plan.nf)
    img = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(np.complex64)
    ksp = (rng.standard_normal(op.ksp_full_shape) + 1j * rng.standard_normal(op.ksp_full_shape)).astype(np.complex64)
    res = {}
    for method in (1, 2):
        plan.set_option(2, method)
        res[method] = (op.op(img), op.adj_op(ksp), op.data_consistency(img, ksp))
        with op.grad_traj_plan():  # opposite sign (+ conjugated smaps)
            res[method] += (op.op(img), op.adj_op(ksp))
    for a, b in zip(res[1], res[2]):
        assert rel_l2(b, a) < 1e-6
    cpu = CpuNufft(samples, shape, precision="f64")
    if np.prod(shape) <= 1e5:
        y_o = cpu.op(img[0, 0], smaps) if sense else cpu.op(img[0])
        x_o = cpu.adj_op(ksp[0], smaps)
        assert rel_l2(res[2][0][0], y_o) <= 2e-6
        assert rel_l2(res[2][1][0, 0] if sense else res[2][1][0], x_o) <= 2e-6


@pytest.mark.parametrize("shape,C", [((64, 64, 64), 3), ((256, 64, 128), 2), ((128, 128), 4), ((64, 256, 64), 1)])
def test_tma_variant_of_the_fft_passes_equals_the_plain_loads(mods, shape, C):
    """Option 2 = 3: the strided FFT passes bring their tiles in with cp.async.bulk.tensor + mbarrier instead of
    per-thread global loads (boxes of zero padding / untouched tiles are not issued).  Same arithmetic in the
    same order: results must agree with option 2 = 2 to rounding of the spreader's / interpolator's atomics
    (tiles cut by chunk boundaries), for op, adj_op (incl. the spreader's untouched tiles: samples confined to
    a ball) and the opposite sign."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(13)
    d, M = len(shape), 30_000
    v = rng.standard_normal((M, d))
    v *= (rng.uniform(0, 1, (M, 1)) ** (1 / d)) * 2.0 / np.linalg.norm(v, axis=1, keepdims=True)
    samples = v.astype(np.float32)
    smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
    smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    plan = op.raw_op.plan
    plan.set_option(4, 32)  # whole tiles are written without atomics: bit-reproducible spreading
    img = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(np.complex64)
    ksp = (rng.standard_normal(op.ksp_full_shape) + 1j * rng.standard_normal(op.ksp_full_shape)).astype(np.complex64)
    res = {}
    for method in (2, 3):
        plan.set_option(2, method)
        res[method] = [op.adj_op(ksp)]
        with op.grad_traj_plan():
            res[method].append(op.adj_op(ksp))
        res[method].append(op.op(img))
    for a, b in zip(res[2], res[3]):
        assert np.all(np.isfinite(b)) and rel_l2(b, a) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("shape,C,sense,dens", [((64, 32), 3, True, True), ((48, 40), 2, False, False),
                                                ((32, 16, 32), 3, True, False), ((24, 20, 16), 1, False, True),
                                                ((22, 26), 2, True, False)])
def test_toeplitz_gram_matches_adj_op_of_op(mods, shape, C, sense, dens):
    """Reference: tests/operators/test_batch.py:275-293 (Toeplitz gram_op vs adj_op(op), rtol 2e-4).
    Pruned-FFT path (2^k grids), cuFFT path (other 2N grids) and the reference construction
    (oversampled grid != 2N)."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(7)
    M = 4000
    samples = rng.uniform(-np.pi, np.pi, (M, len(shape))).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    density = rng.uniform(0.5, 1.5, M).astype(np.float32) if dens else False
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, density=density,
                                       squeeze_dims=False)
    img_shape = (1, 1, *shape) if sense else (1, C, *shape)
    x = (rng.standard_normal(img_shape) + 1j * rng.standard_normal(img_shape)).astype(np.complex64)
    direct = op.adj_op(op.op(x))
    gram = op.gram_op(x, toeplitz=True)
    assert gram.shape == direct.shape
    err = np.linalg.norm(gram - direct) / np.linalg.norm(direct)
    assert err <= 2e-5, err  # tolerance: relative L2; the reference uses rtol 2e-4 element-wise
    # the kernel is cached and re-used; a torch input gives a torch output
    import torch

    xt = torch.from_numpy(x).cuda()
    gt = op.gram_op(xt)
    assert gt.is_cuda and np.allclose(gt.cpu().numpy(), gram, rtol=1e-5, atol=1e-6)
    # changing the density invalidates the cached kernel
    op.density = None
    g2 = op.gram_op(x)
    d2 = op.adj_op(op.op(x))
    assert np.linalg.norm(g2 - d2) / np.linalg.norm(d2) <= 2e-5


@pytest.mark.gpu
def test_stacked_b200_matches_full_3d(mods):
    """`get_operator("stacked-b200")` (generic fallback of base.py:151-158: 2-D b200 NUFFT per plane +
    FFT along z, src/mrinufft/operators/stacked.py:36-357) against the full 3-D b200 operator on a
    stack-of-spirals trajectory (reference: tests/operators/test_stacked.py)."""
    mrinufft, _, _ = mods
    from mrinufft.operators.stacked import stacked2traj3d

    rng = np.random.default_rng(5)
    shape, M2 = (32, 32, 16), 600
    traj2d = rng.uniform(-0.5, 0.5, (M2, 2)).astype(np.float32)
    z_index = np.arange(shape[-1])
    op_st = mrinufft.get_operator("stacked-b200")(traj2d, shape, smaps=None, z_index=z_index, n_coils=2, squeeze_dims=False)
    traj3d = stacked2traj3d(traj2d, z_index, shape[-1]).astype(np.float32)
    op_3d = mrinufft.get_operator("b200")(traj3d, shape, n_coils=2, squeeze_dims=False)
    img = (rng.standard_normal((1, 2, *shape)) + 1j * rng.standard_normal((1, 2, *shape))).astype(np.complex64)
    y_st = np.asarray(op_st.op(img)).reshape(1, 2, -1)
    y_3d = np.asarray(op_3d.op(img)).reshape(1, 2, -1)
    # the two operators use different normalisations of the z transform: compare up to one scalar
    s = np.vdot(y_st, y_3d) / np.vdot(y_st, y_st)
    assert np.linalg.norm(s * y_st - y_3d) <= 1e-4 * np.linalg.norm(y_3d)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,C,sense", [((32, 40), 3, True), ((24, 20, 16), 1, False), ((32, 32), 2, False)])
def test_off_resonance_batched_matches_reference_wrapper(mods, shape, C, sense):
    """`MRIB200NUFFT.with_off_resonance_correction` (interpolators riding the coil batch) against the
    reference's own `MRIFourierCorrected` loop (src/mrinufft/operators/off_resonance.py:232-332) run on
    its exact-NDFT `numpy` backend with the same interpolators, plus adjointness.  The calibrationless
    multi-coil case exercises the fallback to the reference loop on top of the b200 operator."""
    mrinufft, _, _ = mods
    from mrinufft.operators.off_resonance import MRIFourierCorrected

    from mrinufft_b200.off_resonance import MRIB200FourierCorrected

    rng = np.random.default_rng(11)
    NS, NK = 8, 64
    samples = rng.uniform(-0.5, 0.5, (NS * NK, len(shape))).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    b0 = (40 * rng.standard_normal(shape)).astype(np.float32)
    t = np.linspace(0, 5e-3, NK).astype(np.float32)
    ref = MRIFourierCorrected(mrinufft.get_operator("numpy")(samples, shape, n_coils=C, smaps=smaps),
                              b0, t, interpolator={"name": "mti", "L": 5})
    ref.squeeze_dims = False
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    # (a tuple (B, C) cannot be passed: the reference's own decorator turns it into a list and
    # `compute_interpolator` then rejects it, off_resonance.py:153-192 -- same deterministic method instead)
    orc = op.with_off_resonance_correction(t, b0, interpolator={"name": "mti", "L": 5})
    assert np.allclose(np.asarray(orc.B), np.asarray(ref.B)) and np.allclose(np.asarray(orc.C), np.asarray(ref.C))
    assert isinstance(orc, MRIB200FourierCorrected)
    img_shape = (1, 1 if (sense or C == 1) else C, *shape)
    x = (rng.standard_normal(img_shape) + 1j * rng.standard_normal(img_shape)).astype(np.complex64)
    y = (rng.standard_normal((1, C, NS * NK)) + 1j * rng.standard_normal((1, C, NS * NK))).astype(np.complex64)
    ax, ahy = orc.op(x), orc.adj_op(y)
    assert (orc._fused is not None) == (sense or C == 1)
    ax_ref, ahy_ref = ref.op(x), ref.adj_op(y)
    assert ax.shape == ax_ref.shape and ahy.shape == ahy_ref.shape
    assert rel_l2(ax, ax_ref) <= 5e-6 and rel_l2(ahy, ahy_ref) <= 5e-6  # tolerance: exact NDFT, eps=1e-6
    lhs, rhs = np.vdot(ax.ravel(), y.ravel()), np.vdot(x.ravel(), ahy.ravel())
    assert abs(lhs - rhs) <= 5e-5 * abs(lhs)
    # torch in -> torch out, same device
    import torch

    xt = torch.from_numpy(x).cuda()
    yt = orc.op(xt)
    assert yt.is_cuda and np.allclose(yt.cpu().numpy(), ax, rtol=1e-5, atol=1e-6)


def test_paired_batch_autograd_uses_each_items_own_maps(mods):
    """``make_autograd(paired_batch=D)`` (autodiff.py:291-360) with torch CUDA maps per item: forward and data
    gradients of item d are those of maps[d] (same check as tests/test_autodiff_cpu.py on the reference NDFT)."""
    mrinufft, _, torch = mods
    rng = np.random.default_rng(12)
    shape, M, C, D = (12, 16), 300, 3, 3
    samples = rng.uniform(-np.pi, np.pi, (M, 2)).astype(np.float32)
    maps = (rng.standard_normal((D, C, *shape)) + 1j * rng.standard_normal((D, C, *shape))).astype(np.complex64)
    maps /= np.linalg.norm(maps, axis=1, keepdims=True)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=maps[0], squeeze_dims=False)
    ag = op.make_autograd(wrt_data=True, paired_batch=D)
    maps_t = torch.from_numpy(maps).cuda()
    x = torch.from_numpy((rng.standard_normal((D, 1, 1, *shape)) + 1j * rng.standard_normal((D, 1, 1, *shape))
                          ).astype(np.complex64)).cuda().requires_grad_(True)
    w = torch.from_numpy((rng.standard_normal((D, 1, C, M)) + 1j * rng.standard_normal((D, 1, C, M))
                          ).astype(np.complex64)).cuda()
    y = ag.op(x, smaps=maps_t)
    assert tuple(y.shape) == (D, 1, C, M)
    torch.sum(torch.abs(y - w) ** 2).backward()
    for d in range(D):
        op.smaps = maps[d]
        assert rel_l2(y[d].detach().cpu().numpy(), op.op(x[d].detach().cpu().numpy())) < 1e-6
        want = 2 * op.adj_op((y[d] - w[d]).detach().cpu().numpy())
        assert rel_l2(x.grad[d].cpu().numpy(), want) < 1e-5


@pytest.mark.parametrize("C,sense", [(1, False), (3, True)])
def test_field_map_autograd_through_the_batched_orc_operator(mods, C, sense):
    """``with_off_resonance_correction(...).make_autograd(wrt_field_map=True)`` (off_resonance.py:399-446,
    autodiff.py:44-55, 89-100) on the batched b200 operator, CUDA tensors in and out, against torch's
    autograd through the dense model (same check as tests/test_autodiff_cpu.py on the reference NDFT)."""
    mrinufft, _, torch = mods
    from conftest import check_orc_autograd

    from mrinufft_b200.off_resonance import MRIB200FourierCorrected

    rng = np.random.default_rng(3)
    shape, NK = (8, 10), 48
    unit = rng.uniform(-0.5, 0.5, (NK, 2)).astype(np.float32)
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    b0 = (30 * rng.standard_normal(shape)).astype(np.float32)
    t = np.linspace(0, 4e-3, NK).astype(np.float32)
    op = mrinufft.get_operator("b200")(unit * 2 * np.pi, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    orc = op.with_off_resonance_correction(t, b0, interpolator={"name": "mti", "L": 24})
    assert isinstance(orc, MRIB200FourierCorrected)
    ag = orc.make_autograd(wrt_data=True, wrt_field_map=True)
    # (the field-map tensor stays on the host, like the samples of the trajectory gradient; re-assigning it
    # would recompute the interpolators with default settings -- `update_field_map` drops the kwargs,
    # off_resonance.py:199-206)
    check_orc_autograd(ag, orc, unit, shape, t, smaps, C, "cuda", rng)
    assert orc._fused is not None  # the interpolators rode the coil batch


@pytest.mark.gpu
@pytest.mark.parametrize("shape,precision", [((30,), "single"), ((40, 36), "single"), ((64, 32), "single"),
                                             ((20, 24, 28), "single"), ((33, 14, 22), "single"),
                                             ((40, 36), "double"), ((20, 24, 28), "double")])
def test_own_any_length_fft_matches_the_library_fft(mods, shape, precision):
    """Option key 2 = 4: the library's own any-length FFT passes (csrc/fft_any.cu: Stockham in shared memory,
    radix 4 / 2 / 3 / 5 / 7 / direct-DFT primes, fastest axis and strided axes) in place of the cuFFT execution
    of grids that are not powers of two and of complex128 plans: same results, both signs, several coils."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(11)
    d, C, M = len(shape), 3, 3000
    dbl = precision == "double"
    samples = rng.uniform(-np.pi, np.pi, (M, d)).astype(np.float64 if dbl else np.float32)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False, precision=precision)
    cdt = np.complex128 if dbl else np.complex64
    img = (rng.standard_normal(op.img_full_shape) + 1j * rng.standard_normal(op.img_full_shape)).astype(cdt)
    ksp = (rng.standard_normal(op.ksp_full_shape) + 1j * rng.standard_normal(op.ksp_full_shape)).astype(cdt)
    res = {}
    for method in (0, 4):
        op.raw_op.plan.set_option(2, method)
        res[method] = (op.op(img), op.adj_op(ksp))
        with op.grad_traj_plan():
            res[method] += (op.op(img), op.adj_op(ksp))
    for a, b in zip(res[0], res[4]):
        assert rel_l2(b, a) < (1e-13 if dbl else 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dbl", [((3000,), False), ((40, 36), False), ((64, 64), True), ((12, 10, 22), False),
                                        ((16, 32, 8), True), ((7, 13, 1), False)])
def test_plan_less_fft_entry_point(mods, shape, dbl):
    """`b200_fft_c2c` (csrc/fft_any.cu): both signs, several arrays per call, against numpy's fftn in double."""
    _, _, torch = mods
    from mrinufft_b200 import _lib

    rng = np.random.default_rng(5)
    T = 3
    x = (rng.standard_normal((T, *shape)) + 1j * rng.standard_normal((T, *shape))).astype(np.complex128 if dbl else np.complex64)
    axes = tuple(range(1, 1 + len(shape)))
    for sign, ref in ((-1, np.fft.fftn(x.astype(np.complex128), axes=axes)),
                      (+1, np.fft.ifftn(x.astype(np.complex128), axes=axes) * np.prod(shape))):
        d = torch.from_numpy(x.copy()).cuda()
        _lib.fft_c2c(d.data_ptr(), T, shape, sign, dbl, torch.cuda.current_stream().cuda_stream)
        assert rel_l2(d.cpu().numpy(), ref) <= (1e-13 if dbl else 1e-6)


@pytest.mark.gpu
def test_python_mirror_of_the_grid_choice(mods):
    """`_lib.grid_size` (sizes the workspace before a plan exists) == the grid `b200_plan_create` chooses."""
    _, _, torch = mods
    from mrinufft_b200 import _lib

    for shape in [(30,), (320, 320), (225, 225), (105, 111), (192, 192, 192), (160, 160, 160), (224, 224, 224),
                  (96, 100, 36), (24, 28, 12), (20, 36), (17, 19, 23), (64, 64, 64), (225, 225, 40)]:
        for eps, dbl, exact in [(1e-6, False, False), (1e-4, False, False), (1e-12, True, False), (1e-6, False, True)]:
            plan = _lib.Plan(shape, 1, eps=eps, double=dbl, exact_grid=exact)
            assert tuple(plan.nf) == _lib.grid_size(shape, eps, 2.0, dbl, exact), (shape, eps, dbl, exact)
            plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(24, 24, 12), (24, 28, 12), (48, 24, 26)])
def test_power_of_two_grid_for_3d_sizes_with_other_factors(mods, shape):
    """A 3-D plan whose 2N has factors 3 / 5 / 7 takes the next power of two as its grid when that is at most a
    third larger on every axis (csrc/api.cu; B200_EXACT_GRID keeps next235even): same accuracy bar against the
    reference's exact NDFT, SENSE maps, both signs; the Toeplitz Gram operator keeps working through its own
    exact-grid plan where 2N is an admissible grid size."""
    mrinufft, _, torch = mods
    rng = np.random.default_rng(3)
    C, M = 3, 4000
    samples = rng.uniform(-0.5, 0.5, (M, 3)).astype(np.float32)
    smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
    smaps /= np.linalg.norm(smaps, axis=0)
    op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    nf = tuple(op.raw_op.plan.nf)
    assert all(n & (n - 1) == 0 for n in nf) and nf != tuple(2 * s for s in shape)
    ref = mrinufft.get_operator("numpy")(samples.astype(np.float64), shape, n_coils=C, smaps=smaps.astype(np.complex128))
    ref.squeeze_dims = False
    img = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
    ksp = (rng.standard_normal((1, C, M)) + 1j * rng.standard_normal((1, C, M))).astype(np.complex64)
    assert rel_l2(op.op(img), ref.op(img.astype(np.complex128))) <= 5e-6
    assert rel_l2(op.adj_op(ksp), ref.adj_op(ksp.astype(np.complex128))) <= 5e-6
    assert rel_l2(op.data_consistency(img, ksp), ref.adj_op(ref.op(img.astype(np.complex128)) - ksp)) <= 5e-6
    with op.grad_traj_plan():  # opposite sign: conj(A conj(x)) with conjugated maps
        y_flip = op.op(img)
    assert y_flip.shape == ksp.shape and np.isfinite(y_flip).all()
    g = op.gram_op(img, toeplitz=True)
    assert rel_l2(g, op.adj_op(op.op(img))) <= 2e-5
    exact = tuple(2 * s for s in shape)
    if shape == (24, 24, 12):
        assert op._toeplitz_plan() is not None and tuple(op._toeplitz_plan().nf) == exact


@pytest.mark.gpu
@pytest.mark.parametrize("Z,Y,C", [(1, 5, 2), (2, 16, 1), (12, 24, 3), (15, 17, 2), (22, 40, 2), (49, 16, 1),
                                    (64, 33, 4), (97, 8, 1), (176, 20, 2), (256, 32, 2), (512, 16, 1)])
def test_stack_fftz_kernel_against_numpy(mods, Z, Y, C):
    """`b200_stack_fftz_forward` / `_adjoint` (csrc/stack_fftz.cu) for lengths with every kind of factor
    (powers of 4 and 2, 3, 5, 7, the direct-DFT primes 11 and 97, Z = 1), ragged y tiles, with and without
    sensitivity maps, against `_fftz` / `_ifftz` of the reference (stacked.py:178-195) on float64 numpy."""
    _, mb, torch = mods
    from mrinufft_b200 import _lib

    rng = np.random.default_rng(Z * 131 + Y)
    X = 3
    zsel = np.sort(rng.choice(Z, size=max(1, (2 * Z) // 3), replace=False)).astype(np.int32)
    NZ = len(zsel)
    cplx = lambda *sh: (rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(np.complex64)
    fz = lambda a: np.fft.fftshift(np.fft.fft(np.fft.ifftshift(a, axes=-1), axis=-1, norm="ortho"), axes=-1) / np.sqrt(2)
    ifz = lambda a: np.fft.fftshift(np.fft.ifft(np.fft.ifftshift(a, axes=-1), axis=-1, norm="ortho"), axes=-1) / np.sqrt(2)
    smaps = cplx(C, X, Y, Z)
    zs_d = torch.from_numpy(zsel).cuda()
    scale = 1.0 / np.sqrt(2.0 * Z)
    for sense in (False, True):
        img = cplx(1 if sense else C, X, Y, Z)
        sm_d = torch.from_numpy(smaps).cuda() if sense else None
        planes_d = torch.full((C * NZ, X, Y), float("nan"), dtype=torch.complex64, device="cuda")
        _lib.stack_fftz(False, torch.from_numpy(img).cuda().data_ptr(), sm_d.data_ptr() if sense else None,
                        planes_d.data_ptr(), zs_d.data_ptr(), C, X, Y, Z, NZ, scale,
                        torch.cuda.current_stream().cuda_stream)
        coil = img.astype(np.complex128) * smaps if sense else img.astype(np.complex128)
        want = np.moveaxis(fz(coil)[..., zsel], -1, 1).reshape(C * NZ, X, Y)
        assert rel_l2(planes_d.cpu().numpy(), want) <= 1e-6
        planes = cplx(C * NZ, X, Y)
        out_d = torch.full((1 if sense else C, X, Y, Z), float("nan"), dtype=torch.complex64, device="cuda")
        _lib.stack_fftz(True, torch.from_numpy(planes).cuda().data_ptr(), sm_d.data_ptr() if sense else None,
                        out_d.data_ptr(), zs_d.data_ptr(), C, X, Y, Z, NZ, scale,
                        torch.cuda.current_stream().cuda_stream)
        kz = np.zeros((C, X, Y, Z), np.complex128)
        kz[..., zsel] = np.moveaxis(planes.reshape(C, NZ, X, Y), 1, -1)
        want = ifz(kz)
        if sense:
            want = np.sum(want * smaps.conj(), axis=0, keepdims=True)
        assert rel_l2(out_d.cpu().numpy(), want) <= 1e-6


@pytest.mark.gpu
def test_stacked_b200_rejects_repeated_planes(mods):
    mrinufft, _, _ = mods
    traj2d = np.random.default_rng(0).uniform(-0.5, 0.5, (50, 2)).astype(np.float32)
    op = mrinufft.get_operator("stacked-b200")(traj2d, (16, 16, 8), z_index=np.array([1, 1, 3]), n_coils=1)
    with pytest.raises(ValueError, match="distinct"):
        op.op(np.zeros((1, 1, 16, 16, 8), np.complex64))


@pytest.mark.gpu
@pytest.mark.parametrize("sense", [False, True])
def test_native_stacked_matches_generic_stacked(mods, sense):
    """`MRIB200StackedNUFFT` (device resident) against the reference's generic `MRIStackedNUFFT` driving the
    same 2-D b200 operator through host arrays (stacked.py:36-357): op, adj_op, adjointness, partial
    z_index, numpy and torch inputs."""
    mrinufft, mb, torch = mods
    from mrinufft.operators.stacked import MRIStackedNUFFT

    rng = np.random.default_rng(9)
    shape, M2, C = (32, 24, 12), 500, 3
    traj2d = rng.uniform(-0.5, 0.5, (M2, 2)).astype(np.float32)
    z_index = np.array([0, 2, 3, 6, 7, 8, 11])
    smaps = None
    if sense:
        smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
        smaps /= np.linalg.norm(smaps, axis=0)
    nat = mrinufft.get_operator("stacked-b200")(traj2d, shape, smaps=smaps, z_index=z_index, n_coils=C)
    assert isinstance(nat, mb.MRIB200StackedNUFFT)
    gen = MRIStackedNUFFT(traj2d, shape, "b200", smaps, z_index=z_index, n_coils=C)
    img_shape = (1, 1 if sense else C, *shape)
    x = (rng.standard_normal(img_shape) + 1j * rng.standard_normal(img_shape)).astype(np.complex64)
    y = (rng.standard_normal((1, C, len(z_index) * M2)) + 1j * rng.standard_normal((1, C, len(z_index) * M2)))
    y = y.astype(np.complex64)
    ax, ahy = nat.op(x), nat.adj_op(y)
    assert isinstance(ax, np.ndarray) and ax.shape == (1, C, len(z_index) * M2) and ahy.shape == img_shape
    assert rel_l2(ax, gen.op(x)) <= 2e-6 and rel_l2(ahy, gen.adj_op(y)) <= 2e-6
    lhs, rhs = np.vdot(ax.ravel(), y.ravel()), np.vdot(x.ravel(), ahy.ravel())
    assert abs(lhs - rhs) <= 5e-5 * abs(lhs)
    xt = torch.from_numpy(x).cuda()
    assert nat.op(xt).is_cuda and rel_l2(nat.op(xt).cpu().numpy(), ax) <= 1e-6
    dc = nat.data_consistency(x, y)
    assert rel_l2(dc, nat.adj_op(nat.op(x) - y)) <= 1e-5
    if sense:  # device-resident maps (the base setter only takes numpy, base.py:759-760)
        nat_d = mrinufft.get_operator("stacked-b200")(traj2d, shape, smaps=torch.from_numpy(smaps).cuda(),
                                                      z_index=z_index, n_coils=C)
        assert rel_l2(nat_d.op(x), ax) <= 1e-6 and rel_l2(nat_d.adj_op(y), ahy) <= 1e-6


@pytest.mark.gpu
def test_rows_fallback_when_stream_does_not_fit(mods):
    """When the visit stream would overflow its 32-bit indices the tiled kernels hand over to the
    point-driven ones (spread_rows.cu `unsupported`); forced here through option key 3, bit 3."""
    mrinufft, _, _ = mods
    g = load_golden("random3D_sense")
    op = make_op(mrinufft, g)
    ref_y, ref_x = op.op(g["img"]), op.adj_op(g["ksp"])
    op2 = make_op(mrinufft, g)
    op2.raw_op.plan.set_option(3, 8)
    op2.raw_op._set_pts(op2.samples)
    assert rel_l2(op2.op(g["img"]), ref_y) <= 2e-6 and rel_l2(op2.adj_op(g["ksp"]), ref_x) <= 2e-6
    assert rel_l2(op2.op(g["img"]), g["op"]) <= 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("case,eps,tol", [("random2D_sense", 1e-6, 5e-6), ("random3D", 1e-6, 5e-6),
                                          ("random2D", 1e-11, 2e-10), ("cones3D", 1e-10, 2e-9)])
def test_double_precision_path_matches_reference_ndft(mods, case, eps, tol):
    """`precision="double"` (complex128 kernels, B200_DOUBLE plans) against the reference's NDFT goldens:
    the same bar as single precision at eps = 1e-6, and the accuracy only double can reach at small eps
    (rel-L2 within ~20 eps of the exact transform).  Outputs are complex128, inputs are not mutated."""
    mrinufft, _, torch = mods
    g = load_golden(case)
    op = mrinufft.get_operator("b200")(
        g["samples"].astype(np.float64), g["shape"], n_coils=g["n_coils"], smaps=g.get("smaps"),
        squeeze_dims=False, eps=eps, precision="double")
    assert op.cpx_dtype == np.complex128 and op.samples.dtype == np.float64
    # exact transforms of the float64 samples (the goldens were made with float32 sample values, which
    # convert exactly): evaluate the NDFT with the reference's numpy backend in double
    ref = mrinufft.get_operator("numpy")(g["samples"].astype(np.float64), g["shape"], n_coils=g["n_coils"],
                                         smaps=g.get("smaps"))
    ref.squeeze_dims = False
    img = g["img"].astype(np.complex128)
    ksp = g["ksp"].astype(np.complex128)
    y, x = op.op(img), op.adj_op(ksp)
    assert y.dtype == np.complex128 and x.dtype == np.complex128
    assert rel_l2(y, ref.op(img)) <= tol and rel_l2(x, ref.adj_op(ksp)) <= tol
    dc = op.data_consistency(img, ksp)
    assert rel_l2(dc, ref.adj_op(ref.op(img) - ksp)) <= 2 * tol
    lhs, rhs = np.vdot(y.ravel(), ksp.ravel()), np.vdot(img.ravel(), x.ravel())
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    yt = op.op(torch.from_numpy(img).cuda())
    assert yt.is_cuda and yt.dtype == torch.complex128


@pytest.mark.gpu
def test_empty_tile_skipping_follows_update_samples(mods):
    """On 3-D power-of-two grids tiles without visitors are neither written by the spreader nor read by the
    FFT (and not written by the type-2 FFT): the bit strings must follow `update_samples`.  A compact
    trajectory (most tiles empty) is replaced by a full one and vice versa; results must equal those of
    freshly built operators, for both transform types, with coil chunks (accumulate) and SENSE."""
    mrinufft, _, _ = mods
    rng = np.random.default_rng(21)
    shape, C, M = (32, 32, 32), 5, 6000
    compact = (0.12 * rng.standard_normal((M, 3))).clip(-0.49, 0.49).astype(np.float32)
    full = rng.uniform(-0.5, 0.5, (M, 3)).astype(np.float32)
    smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
    smaps /= np.linalg.norm(smaps, axis=0)
    img = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
    ksp = (rng.standard_normal((1, C, M)) + 1j * rng.standard_normal((1, C, M))).astype(np.complex64)
    kw = dict(n_coils=C, smaps=smaps, squeeze_dims=False, coil_chunk=2)
    op = mrinufft.get_operator("b200")(compact, shape, **kw)
    for a, b in ((compact, full), (full, compact)):
        op.update_samples(a)
        _ = op.op(img), op.adj_op(ksp)
        op.update_samples(b)
        fresh = mrinufft.get_operator("b200")(b, shape, **kw)
        assert rel_l2(op.op(img), fresh.op(img)) <= 1e-6
        assert rel_l2(op.adj_op(ksp), fresh.adj_op(ksp)) <= 1e-6
        assert rel_l2(op.data_consistency(img, ksp), fresh.data_consistency(img, ksp)) <= 1e-6
    # ... and the compact case against the exact NDFT (sampled): adjointness is the cheap strong check
    op.update_samples(compact)
    y, x = op.op(img), op.adj_op(ksp)
    lhs, rhs = np.vdot(y.ravel(), ksp.ravel()), np.vdot(img.ravel(), x.ravel())
    assert abs(lhs - rhs) <= 5e-5 * abs(lhs)
