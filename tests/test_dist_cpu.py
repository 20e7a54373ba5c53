"""CPU tests of the multi-GPU host logic: world_size 2, gloo backend, numpy stand-in for the
rank-local operator (the exact-NDFT oracle).  N-rank result == 1-rank result."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


class _NumpyLocalOp:
    """Rank-local stand-in with the operator surface used by CoilShardedOperator (test only)."""

    def __init__(self, samples, shape, n_coils=1, smaps=None, squeeze_dims=False):
        from oracle import es_nufft as E

        self.shape = tuple(shape)
        self.n_coils = n_coils
        self.smaps = smaps
        self.A = E.ndft_matrix(samples, shape) / np.sqrt(np.prod(shape) * 2.0 ** len(shape))

    def op(self, image):
        if self.smaps is not None:
            return np.stack([self.A @ (image.reshape(self.shape) * s).ravel() for s in self.smaps])[None]
        return np.stack([self.A @ im.ravel() for im in image.reshape(self.n_coils, -1)])[None]

    def adj_op(self, ksp):
        ksp = ksp.reshape(self.n_coils, -1)
        if self.smaps is not None:
            return sum(np.conj(s) * (self.A.conj().T @ k).reshape(self.shape)
                       for s, k in zip(self.smaps, ksp))[None, None]
        return np.stack([(self.A.conj().T @ k).reshape(self.shape) for k in ksp])[None]

    def data_consistency(self, image, obs):
        return self.adj_op(self.op(image) - obs.reshape(1, self.n_coils, -1))

    # device-level surface used by mrinufft_b200.solvers (CPU tensors here)
    device = torch.device("cpu")
    _cdt = torch.complex128
    squeeze_dims = False
    uses_density = False
    density = None
    n_batchs = 1

    @property
    def img_full_shape(self):
        return (1, 1 if self.smaps is not None else self.n_coils, *self.shape)

    @property
    def ksp_full_shape(self):
        return (1, self.n_coils, self.A.shape[0])

    def _op_device(self, img):
        return torch.from_numpy(self.op(img.numpy()))

    def _adj_device(self, ksp):
        return torch.from_numpy(np.ascontiguousarray(self.adj_op(ksp.numpy())))

    def _dc_device(self, img, obs):
        return torch.from_numpy(np.ascontiguousarray(self.data_consistency(img.numpy(), obs.numpy())))

    def get_lipschitz_cst(self, max_iter=10):
        # the largest eigenvalue of A^H A plus a rank-dependent error, like a power method started
        # from unseeded random numbers would give: the sharded operator must agree on rank 0's value
        return float(np.linalg.norm(self.A, 2) ** 2) * (1.0 + 0.01 * self._rank)

    _rank = 0


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "baseline" / "_ref"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mrinufft_b200.dist import CoilShardedOperator, coil_slice

    rng = np.random.default_rng(0)  # same data on every rank
    shape, M, C = (8, 10), 120, 5
    samples = rng.uniform(-np.pi, np.pi, (M, 2)).astype(np.float32)
    smaps = rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))
    img = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    ksp = rng.standard_normal((C, M)) + 1j * rng.standard_normal((C, M))
    full = _NumpyLocalOp(samples, shape, C, smaps)
    lo, hi = coil_slice(C, rank, world)

    sh = CoilShardedOperator(samples, shape, C, smaps=smaps, local_factory=_NumpyLocalOp)
    assert (sh.lo, sh.hi) == (lo, hi)
    y = sh.op(img)
    assert np.allclose(y, full.op(img)[:, lo:hi])
    x = sh.adj_op(ksp[lo:hi])
    assert np.allclose(x, full.adj_op(ksp))                      # all-reduced == single rank
    g = sh.data_consistency(img, ksp[lo:hi])
    assert np.allclose(g, full.data_consistency(img, ksp))

    # calibrationless: results stay sharded, scalars are all-reduced
    imgs = rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))
    fullc = _NumpyLocalOp(samples, shape, C, None)
    shc = CoilShardedOperator(samples, shape, C, smaps=None, local_factory=_NumpyLocalOp)
    assert np.allclose(shc.op(imgs[lo:hi]), fullc.op(imgs)[:, lo:hi])
    xl = shc.adj_op(ksp[lo:hi])
    assert np.allclose(xl, fullc.adj_op(ksp)[:, lo:hi])
    local_dot = torch.tensor(np.vdot(xl, xl).real)
    tot = shc.reduce_scalar(local_dot)
    ref = np.vdot(fullc.adj_op(ksp), fullc.adj_op(ksp)).real
    assert abs(float(tot) - ref) / ref < 1e-12
    assert float(sh.reduce_scalar(local_dot)) == float(local_dot)  # SENSE: replicated iterate
    # pinv_solver on coil-sharded data == the same solver on one rank (SENSE: replicated image,
    # calibrationless: this rank's coils), with a rank-dependent Lipschitz estimate on purpose
    from mrinufft_b200 import solvers

    sh.local._rank = rank
    shc.local._rank = rank
    for name in ("lsqr", "lsmr", "cg"):
        kw = dict(max_iter=6)
        if name == "cg":
            kw["lipschitz_cst"] = full.get_lipschitz_cst()
        want = solvers.SOLVERS[name](full, ksp[None], **kw)
        got = sh.pinv_solver(ksp[None, lo:hi], optim=name, max_iter=6)
        assert got.shape == want.shape and np.allclose(got, want, rtol=1e-9, atol=1e-11), name
        if name == "cg":
            kw["lipschitz_cst"] = fullc.get_lipschitz_cst()
        wantc = solvers.SOLVERS[name](fullc, ksp[None], damp=0.2, **kw)
        gotc = shc.pinv_solver(ksp[None, lo:hi], optim=name, max_iter=6, damp=0.2)
        assert np.allclose(gotc, wantc[:, lo:hi], rtol=1e-9, atol=1e-11), name
    Path(tmp, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_coil_sharded_world2_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
