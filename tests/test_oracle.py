"""CPU tests: the oracle (finufft-algorithm restatement) is pinned to the reference.

* against the golden vectors produced by the reference's exact NDFT backend
  (tests/golden/make_golden.py; reference test tests/operators/test_operator_ref.py:54-74 uses
  atol = rtol = 1e-4 between finufft and that NDFT -- the oracle is held to 3e-6 relative L2);
* against the reference's known-answer test (tests/test_ndft.py:58-79);
* numpy restatement == C restatement; fold/sort spec self-consistency.
"""

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, rel_l2
from oracle import es_nufft as E
from oracle.c_oracle import CpuNufft, fold


def _oracle_ops(g, precision="f64", cls=CpuNufft):
    cpu = cls(g["samples"], g["shape"], eps=1e-6, precision=precision)
    smaps = g.get("smaps")
    if smaps is not None:
        y = cpu.op(g["img"][0, 0], smaps)[None]
        x = cpu.adj_op(g["ksp"][0], smaps)[None, None]
        xd = cpu.adj_op(g["ksp"][0], smaps, density=g["density"])[None, None]
    else:
        y = np.stack([cpu.op(g["img"][0, c])[0] for c in range(g["n_coils"])])[None]
        x = np.stack([cpu.adj_op(g["ksp"][0, c])[0] for c in range(g["n_coils"])])[None]
        xd = np.stack([cpu.adj_op(g["ksp"][0, c], density=g["density"])[0] for c in range(g["n_coils"])])[None]
    return y, x, xd


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_c_oracle_matches_reference_ndft(case):
    g = load_golden(case)
    y, x, xd = _oracle_ops(g)
    assert rel_l2(y, g["op"]) < 3e-6
    assert rel_l2(x, g["adj"]) < 3e-6
    assert rel_l2(xd, g["adj_density"]) < 3e-6


@pytest.mark.parametrize("case", ["random2D", "random3D"])
def test_float32_oracle_close(case):
    g = load_golden(case)
    y, x, _ = _oracle_ops(g, "f32")
    assert rel_l2(y, g["op"]) < 5e-6
    assert rel_l2(x, g["adj"]) < 5e-6


@pytest.mark.parametrize("case", ["random2D", "grid2D", "random3D"])
def test_numpy_restatement_equals_c(case):
    g = load_golden(case)
    o = E.ESNufftOracle(g["samples"], g["shape"], eps=1e-6)
    cpu = CpuNufft(g["samples"], g["shape"], eps=1e-6, precision="f64")
    img, ksp = g["img"][0, 0], g["ksp"][0, 0]
    assert rel_l2(cpu.type2(img), o.type2(img)) < 5e-7   # C uses the float32 first-tap offset
    assert rel_l2(cpu.type1(ksp), o.type1(ksp)) < 5e-7


def test_known_answer_cartesian_grid():
    """reference tests/test_ndft.py:58-79: NDFT on the fftfreq grid == fftshift(fftn(fftshift))."""
    import scipy.fft as sfft

    g = load_golden("grid2D")
    img = g["img"][0, 0].astype(np.complex128)
    ref = sfft.fftshift(sfft.fftn(sfft.fftshift(img)))
    norm = np.sqrt(np.prod(g["shape"]) * 4.0)
    assert rel_l2(g["op"].reshape(g["shape"]) * norm, ref) < 1e-6   # the golden itself
    cpu = CpuNufft(g["samples"], g["shape"], precision="f64")
    assert rel_l2(cpu.type2(img).reshape(g["shape"]), ref) < 3e-6
    A = E.ndft_matrix(g["samples"], g["shape"])
    assert rel_l2((A @ img.ravel()).reshape(g["shape"]), ref) < 1e-6   # restated matrix


def test_restated_ndft_matrix_matches_golden():
    g = load_golden("random2D")
    A = E.ndft_matrix(g["samples"], g["shape"]) / np.sqrt(np.prod(g["shape"]) * 4.0)
    assert rel_l2(A @ g["img"].reshape(-1).astype(np.complex128), g["op"].ravel()) < 1e-9
    idx = np.arange(0, 1000, 13)
    s = E.ndft_type2_sampled(g["samples"], g["img"][0, 0].astype(np.complex128), idx)
    assert rel_l2(s / np.sqrt(np.prod(g["shape"]) * 4.0), g["op"].ravel()[idx]) < 1e-9


@pytest.mark.parametrize("eps,w", [(1e-2, 3), (1e-4, 5), (1e-6, 7), (1e-8, 9)])
def test_kernel_parameters_and_accuracy(eps, w):
    assert E.kernel_params(eps, 2.0)[0] == w
    g = load_golden("random2D")
    o = CpuNufft(g["samples"], g["shape"], eps=eps, precision="f64")
    assert o.w == w and o.nfs == (128, 256)
    err = rel_l2(o.op(g["img"][0, 0])[None], g["op"][0])
    assert err < max(3 * eps, 3e-7)  # the C oracle keeps the first-tap offset in float32


def test_fine_grid_sizes():
    assert E.next235even(148) == 150 and E.next235even(512) == 512 and E.next235even(640) == 640
    assert E.fine_grid_size(74, 7, 2.0) == 150 and E.fine_grid_size(5, 7, 2.0) == 16
    assert E.kernel_params(1e-6, 1.25)[0] == 10


def test_fold_spec_numpy_equals_c_and_edges():
    pi32 = np.float32(np.pi)
    rng = np.random.default_rng(0)
    x = rng.uniform(-3 * np.pi, 3 * np.pi, 100000).astype(np.float32)
    x[:8] = [0.0, pi32, -pi32, np.nextafter(pi32, np.float32(4)), -np.nextafter(pi32, np.float32(4)),
             1e-45, 3 * pi32, -3 * pi32]
    for nf, w in [(128, 7), (150, 7), (512, 5), (16, 7), (640, 10)]:
        o, x1, g = E.fold_points(x, nf, w)
        o2, x12 = fold(x, nf, w)
        assert np.array_equal(o, o2) and np.array_equal(x1.view(np.uint32), x12.view(np.uint32))
        assert o.min() >= 0 and o.max() < nf
        assert np.all(x1 >= -w / 2) and np.all(x1 < -w / 2 + 1 + 1e-6)
        assert np.all((g >= 0) & (g < nf))
    # x = +-pi fold onto the same fine-grid position (periodicity), x = 0 sits at nf/2
    o, x1, g = E.fold_points(np.array([pi32, -pi32, 0.0], np.float32), 128, 7)
    assert abs(g[0] - g[1]) < 1e-4 or abs(abs(g[0] - g[1]) - 128) < 1e-4
    assert g[2] == 64.0


def test_bin_sort_is_stable_and_ordered():
    rng = np.random.default_rng(1)
    s = rng.uniform(-np.pi, np.pi, (5000, 3)).astype(np.float32)
    s[100:200] = s[100]  # duplicates keep their input order
    r = E.bin_sort(s, (64, 48, 80), 7)
    ks = r["key"][r["perm"]]
    assert np.all(np.diff(ks) >= 0)
    same = np.diff(ks) == 0
    assert np.all(np.diff(r["perm"])[same] > 0)
    assert r["key"].max() < 64 * 48 * (80 // E.BIN_X) * 2 and r["key"].min() >= 0


def test_empty_and_single_point():
    shape = (16, 16)
    one = np.array([[0.3, -1.2]], np.float32)
    cpu = CpuNufft(one, shape, precision="f64")
    rng = np.random.default_rng(0)
    img = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    A = E.ndft_matrix(one, shape)
    assert rel_l2(cpu.type2(img)[0], A @ img.ravel()) < 3e-6
    assert rel_l2(cpu.type1(np.array([1 + 2j])).ravel(), A.conj().T @ np.array([1 + 2j])) < 3e-6
    r = E.bin_sort(np.zeros((0, 2), np.float32), (32, 32), 7)
    assert r["perm"].shape == (0,)


@pytest.mark.timeout(300)
def test_committed_goldens_are_what_the_reference_produces(tmp_path):
    """Re-run tests/golden/make_golden.py against the reference checkout (build container only) and compare
    with the committed fixtures: they are outputs of the unmodified reference, not of anything in this repo."""
    import subprocess
    import sys
    from pathlib import Path

    if not Path("/root/reference/src/mrinufft").is_dir():
        pytest.skip("the reference checkout is only present in the build container")
    golden = Path(__file__).resolve().parent / "golden"
    r = subprocess.run([sys.executable, str(golden / "make_golden.py"), str(tmp_path)],
                       capture_output=True, text=True, cwd=tmp_path, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    made = sorted(p.name for p in tmp_path.glob("*.npz"))
    assert made == sorted(f"{c}.npz" for c in GOLDEN_CASES + ["cg2D_sense"])
    for name in made:
        with np.load(golden / name) as a, np.load(tmp_path / name) as b:
            assert sorted(a.files) == sorted(b.files)
            for k in a.files:
                assert np.allclose(a[k], b[k], rtol=1e-6, atol=1e-8), (name, k)
