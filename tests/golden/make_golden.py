"""Generate the golden vectors of tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_golden.py

It imports the unmodified reference (``/root/reference/src``) and evaluates its exact-NDFT backend
``MRInumpy`` (``src/mrinufft/operators/interfaces/nudft_numpy.py:133-150``, the reference of
``tests/operators/test_operator_ref.py:54-74``) through the reference's own
``get_operator("numpy")`` on seeded inputs shaped like the reference's trajectory cases
(``tests/case_trajectories.py``).  Stored: the float32 sample locations, the complex64 inputs, and
the complex128 outputs of ``op`` / ``adj_op`` (including the reference's ``1/norm_factor``), SENSE
and calibrationless, plus a ``cg`` run (``extras/optim.py:801-902``) on the NDFT backend.
"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/src")
import mrinufft  # noqa: E402
from mrinufft import get_operator  # noqa: E402
from mrinufft.trajectories import initialize_2D_radial, initialize_2D_spiral  # noqa: E402
from mrinufft.trajectories import initialize_3D_cones  # noqa: E402
from scipy.stats import truncnorm  # noqa: E402

OUT = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent   # optional: output directory


def crandn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def make_case(name, samples_unit, shape, n_coils, sense, seed):
    """samples_unit: (M, d) in [-0.5, 0.5] float64 -> stored as float32 radians."""
    rng = np.random.default_rng(seed)
    samples = (samples_unit.reshape(-1, len(shape)) * 2 * np.pi).astype(np.float32)
    smaps = None
    if sense:
        smaps = crandn(rng, n_coils, *shape)
        smaps /= np.linalg.norm(smaps, axis=0, keepdims=True)
    # float32 radians are handed to the reference exactly as the b200 backend will see them
    op = get_operator("numpy")(samples.astype(np.float64), shape, n_coils=n_coils, smaps=smaps)
    op.squeeze_dims = False
    img = crandn(rng, *op.img_full_shape)
    ksp = crandn(rng, *op.ksp_full_shape)
    y = op.op(img)
    x = op.adj_op(ksp)
    density = rng.uniform(0.5, 1.5, samples.shape[0]).astype(np.float32)
    opd = get_operator("numpy")(samples.astype(np.float64), shape, n_coils=n_coils, smaps=smaps)
    opd.squeeze_dims = False
    opd.density = density
    xd = opd.adj_op(ksp)
    dc = op.data_consistency(img, ksp)
    out = dict(samples=samples, shape=np.asarray(shape), n_coils=n_coils, img=img, ksp=ksp,
               op=y.astype(np.complex128), adj=x.astype(np.complex128),
               density=density, adj_density=xd.astype(np.complex128), dc=dc.astype(np.complex128))
    if smaps is not None:
        out["smaps"] = smaps
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, samples.shape, shape, "norm", op.norm_factor)


def main():
    np.random.seed(0)
    # case_random2D (tests/case_trajectories.py:15-21): M=1000 truncnorm sigma 0.16 on (64,128)
    s = truncnorm(-3, 3, loc=0, scale=0.16).rvs(size=(1000, 2), random_state=0)
    make_case("random2D", s, (64, 128), 1, False, 1)
    make_case("random2D_sense", s, (64, 128), 3, True, 2)
    # case_random3D shrunk (M=200000 on (64,128,74) needs a 970 GB matrix): same law, small grid
    s = truncnorm(-3, 3, loc=0, scale=0.16).rvs(size=(600, 3), random_state=1)
    make_case("random3D", s, (16, 24, 20), 1, False, 3)
    make_case("random3D_sense", s, (16, 24, 20), 2, True, 4)
    # case_nyquist_radial2D (L41-44): 128 x 16 on 32^2, reaches |k| = 0.5 exactly
    s = initialize_2D_radial(128, 16)
    make_case("nyquist_radial2D", s, (32, 32), 2, False, 5)
    # spiral, multi coil
    s = initialize_2D_spiral(8, 256, nb_revolutions=4)
    make_case("spiral2D_sense", s, (48, 40), 4, True, 6)
    # case_grid2D (L64-68): Cartesian grid N=16 on fftfreq -> known answer fftn(fftshift)
    N = 16
    f = np.fft.fftshift(np.fft.fftfreq(N))
    g = np.stack(np.meshgrid(f, f, indexing="ij"), -1)
    make_case("grid2D", g, (N, N), 1, False, 7)
    # 3D cones, non power-of-two (even) sizes.  Odd sizes are left out on purpose: the reference
    # NDFT uses half-integer positions linspace(-s/2, s/2-1, s) for odd s (nudft_numpy.py:38)
    # whereas finufft -- the parity target -- uses the integer modes -(s-1)/2..(s-1)/2.
    s = initialize_3D_cones(16, 32)
    make_case("cones3D", s, (14, 18, 22), 1, False, 8)

    # cg golden on the NDFT backend: 16x24 image, M=600, 4 coils with smaps
    rng = np.random.default_rng(11)
    s = truncnorm(-3, 3, loc=0, scale=0.16).rvs(size=(600, 2), random_state=3)
    samples = (s * 2 * np.pi).astype(np.float32)
    shape = (16, 24)
    smaps = crandn(rng, 4, *shape)
    smaps /= np.linalg.norm(smaps, axis=0, keepdims=True)
    op = get_operator("numpy")(samples.astype(np.float64), shape, n_coils=4, smaps=smaps)
    x_true = crandn(rng, *shape)
    y = op.op(x_true).astype(np.complex64)
    np.random.seed(1234)  # the power method starts from np.random.random (base.py:1194)
    lip = op.get_lipschitz_cst()
    np.random.seed(1234)
    x_cg = op.pinv_solver(y, optim="cg", max_iter=10, progressbar=False)
    np.savez_compressed(OUT / "cg2D_sense.npz", samples=samples, shape=np.asarray(shape), smaps=smaps,
                        y=y, x_true=x_true, lipschitz=np.float64(lip), x_cg=x_cg.astype(np.complex128))
    print("cg golden: lipschitz", lip, "rel err", np.linalg.norm(x_cg - x_true) / np.linalg.norm(x_true))
    print("reference version", getattr(mrinufft, "__version__", "?"))


if __name__ == "__main__":
    main()
