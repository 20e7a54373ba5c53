#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (configs[0], [1], [3]) on one GPU: wall-clock of the public
API calls with device-resident torch inputs (CUDA events), one JSON line per configuration.

    python tools/bench_configs.py [--quick]

cfg-A  README demo: 2-D radial 100 x 500, 512^2, single coil, density="voronoi"
cfg-B  2-D spiral 64 x 2048, 320^2, 32 coils with smaps: op / adj_op / data_consistency
cfg-D  3-D 256^3, 16 coils with smaps, density="pipe" + pinv_solver(optim="cg", max_iter=10) on a smooth
       phantom: NRMSE per iteration, as written and with the step size of the operator that is iterated on
cfg-E  2-D 256^2, 8 coils with smaps, M = 32768, off-resonance correction with L = 10 interpolators riding
       the coil batch, one unrolled gradient step with torch autograd (data + field-map gradients)
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for _p in (ROOT, ROOT / "baseline" / "_ref"):
    if _p.exists() and str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import mrinufft  # noqa: E402
import mrinufft_b200  # noqa: E402,F401
from mrinufft.trajectories import (initialize_2D_radial, initialize_2D_spiral,  # noqa: E402
                                   initialize_3D_phyllotaxis_radial)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def crandn(*s):
    return torch.view_as_complex(torch.randn(*s, 2, device="cuda", dtype=torch.float32))


def cfg_a():
    """BASELINE configs[0], the README demo as written: density="voronoi" through the constructor."""
    traj = initialize_2D_radial(100, 500).astype(np.float32)
    t0 = time.perf_counter()
    op = mrinufft.get_operator("b200")(traj, (512, 512), density="voronoi")
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    img, ksp = crandn(512, 512), crandn(op.n_samples)
    return {"config": "A: README demo, 2D radial 100x500, 512^2, 1 coil, density=voronoi", "setup_s_incl_voronoi": setup,
            "op_ms": timed(lambda: op.op(img), 20), "adj_op_ms": timed(lambda: op.adj_op(ksp), 20),
            "rows_class": op.raw_op.plan.rows_class(1)["class"]}


def cfg_b():
    traj = initialize_2D_spiral(64, 2048, nb_revolutions=8).astype(np.float32)
    C, shape = 32, (320, 320)
    smaps = crandn(C, *shape)
    smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    img, ksp = crandn(1, 1, *shape), crandn(1, C, op.n_samples)
    M = op.n_samples
    r = {"config": f"B: 2D spiral 64x2048 (M={M}), 320^2, 32 coils with smaps, eps=1e-6",
         "op_ms": timed(lambda: op.op(img), 20), "adj_op_ms": timed(lambda: op.adj_op(ksp), 20),
         "data_consistency_ms": timed(lambda: op.data_consistency(img, ksp), 20),
         "gram_toeplitz_ms": timed(lambda: op.gram_op(img), 20)}
    r["pair_k_samples_per_s"] = M * C / ((r["op_ms"] + r["adj_op_ms"]) * 1e-3) / 1e3
    return r


def smooth_phantom(n, dev):
    """Sum of soft-edged ellipsoids and Gaussian blobs, complex with a slowly varying phase: a volume whose
    spectrum lives inside the ball a radial trajectory samples (a white-noise volume does not)."""
    ax = torch.linspace(-1, 1, n, device=dev)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = torch.zeros((n, n, n), device=dev)
    for (cz, cy, cx, rz, ry, rx, amp) in [(0, 0, 0, .8, .7, .6, 1.0), (.1, -.2, .15, .35, .3, .25, -.4),
                                          (-.25, .2, -.2, .2, .25, .15, .6), (.3, .3, .1, .12, .1, .15, .8)]:
        r = torch.sqrt(((z - cz) / rz) ** 2 + ((y - cy) / ry) ** 2 + ((x - cx) / rx) ** 2)
        vol += amp * torch.sigmoid((1 - r) * 5)
    phase = 0.6 * torch.sin(2.0 * x + 1.0) * torch.cos(1.5 * y) + 0.3 * z
    return (vol * torch.exp(1j * phase)).to(torch.complex64)


def smooth_smaps(C, n, dev):
    """Birdcage-like sensitivity maps: coils on a sphere around the volume, magnitude falling with distance,
    a linear phase per coil, normalised to unit root-sum-of-squares."""
    ax = torch.linspace(-1, 1, n, device=dev)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
    maps = torch.empty((C, n, n, n), dtype=torch.complex64, device=dev)
    for c in range(C):
        th, ph = np.pi * (c + 0.5) / C, 2 * np.pi * ((c * 0.618) % 1.0)
        cz, cy, cx = 1.4 * np.cos(th), 1.4 * np.sin(th) * np.sin(ph), 1.4 * np.sin(th) * np.cos(ph)
        d2 = (z - cz) ** 2 + (y - cy) ** 2 + (x - cx) ** 2
        maps[c] = torch.exp(-d2 / 1.5) * torch.exp(1j * (1.2 * (cx * x + cy * y + cz * z) + 0.4 * c))
    maps /= torch.sqrt(torch.sum(maps.real ** 2 + maps.imag ** 2, dim=0, keepdim=True))
    return maps


def cfg_d(quick):
    """BASELINE configs[3] as written: 3-D 256^3, 16 coils with sensitivity maps, density="pipe" through the
    constructor, pinv_solver(optim="cg", max_iter=10) -- a SENSE reconstruction of a smooth phantom from its
    noisy radial k-space, NRMSE against the truth after every iteration.

    The reference's `cg` takes its step size from the density-WEIGHTED operator and then iterates on the
    un-weighted one (extras/optim.py:839-842): with Pipe's normalisation the weighted Lipschitz constant is a
    fraction of the un-weighted one and the iteration leaves the (good) density-compensated start it was given
    and diverges -- on the reference's own exact NDFT too (tests/test_solvers_cpu.py pins that).  `as_written`
    mirrors it; `step_from_iterated_operator` hands `cg` the Lipschitz constant of the operator it iterates
    on (`lipschitz_cst=`, the argument the coil-sharded solver uses) and is the reconstruction."""
    n = 128 if quick else 256
    traj = initialize_3D_phyllotaxis_radial(4096 if quick else 16384, 512).astype(np.float32).reshape(-1, 3)
    C, shape, dev = 16, (n, n, n), torch.device("cuda")
    smaps = smooth_smaps(C, n, dev)
    x_true = smooth_phantom(n, dev).reshape(1, 1, *shape)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, density="pipe", squeeze_dims=False)
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    ksp = op._op_device(x_true)
    ksp += 0.01 * torch.linalg.norm(ksp) / np.sqrt(ksp.numel()) * crandn(*ksp.shape)
    nt = torch.linalg.norm(x_true)
    hist = []

    def cb(x, operator, y, **kw):
        hist.append(float(torch.linalg.norm(x.reshape(x_true.shape) - x_true) / nt))

    out = {"config": f"D: 3D {n}^3, 16 coils with smaps, M={op.n_samples} radial, density=pipe, "
                     "pinv_solver(optim=cg, max_iter=10), smooth phantom + 1 % noise",
           "setup_s_incl_pipe_density": setup, "rows_class_16_coils": op.raw_op.plan.rows_class(C)["class"]}
    x_dc = op._adj_device(ksp)
    x_dc *= torch.vdot(x_dc.ravel(), x_true.ravel()) / torch.vdot(x_dc.ravel(), x_dc.ravel())
    out["nrmse_density_compensated_adjoint_best_scale"] = float(torch.linalg.norm(x_dc - x_true) / nt)
    del x_dc
    np.random.seed(0)
    lip_w = float(op.get_lipschitz_cst())
    dens = op.density
    op.density = None
    np.random.seed(0)
    lip_u = float(op.get_lipschitz_cst())
    op.density = dens
    out["lipschitz_density_weighted"], out["lipschitz_unweighted"] = lip_w, lip_u
    for name, kw in (("as_written", {}), ("step_from_iterated_operator", {"lipschitz_cst": lip_u})):
        hist.clear()
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, _ = op.pinv_solver(ksp, optim="cg", max_iter=10, callback=cb, **kw)
        torch.cuda.synchronize()
        out[name] = {"cg_10_iterations_s_incl_callback": time.perf_counter() - t0,
                     "nrmse_per_iteration": [round(h, 5) if np.isfinite(h) else str(h) for h in hist],
                     "monotone": bool(all(b <= a for a, b in zip(hist, hist[1:]))),
                     "result_finite": bool(torch.isfinite(x).all())}
        del x
    # the reference's default optimiser, device resident (mrinufft_b200/solvers.py), from a zero start
    hist.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, _ = op.pinv_solver(ksp, optim="lsqr", max_iter=10, callback=cb)
    torch.cuda.synchronize()
    out["lsqr"] = {"lsqr_10_iterations_s_incl_callback": time.perf_counter() - t0,
                   "nrmse_per_iteration": [round(h, 5) for h in hist]}
    del x
    return out


def cfg_e():
    C, shape, L = 8, (256, 256), 10
    traj = initialize_2D_spiral(16, 2048, nb_revolutions=8).astype(np.float32).reshape(-1, 2)
    smaps = crandn(C, *shape)
    smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
    yy, xx = np.meshgrid(np.linspace(-1, 1, shape[0]), np.linspace(-1, 1, shape[1]), indexing="ij")
    b0 = (60.0 * np.exp(-(xx ** 2 + yy ** 2) * 2)).astype(np.float32)
    t = np.linspace(0, 20e-3, 2048).astype(np.float32)
    orc = op.with_off_resonance_correction(t, b0, interpolator={"name": "mti", "L": L})
    ag = orc.make_autograd(wrt_data=True, wrt_field_map=True)
    y = crandn(1, C, op.n_samples)
    x = crandn(1, 1, *shape).requires_grad_(True)

    def step():
        x.grad = None
        ag.field_map.grad = None
        loss = torch.mean(torch.abs(ag.op(x) - y) ** 2)
        loss.backward()

    r = {"config": f"E: 2D spiral 16x2048 (M={op.n_samples}), 256^2, {C} coils with smaps, ORC with {orc.n_interpolators} "
                   "interpolators riding the coil batch, autograd (data + field map)",
         "orc_op_ms": timed(lambda: orc.op(x.detach()), 20),
         "orc_adj_op_ms": timed(lambda: orc.adj_op(y), 20),
         "unrolled_step_forward_backward_ms": timed(step, 10),
         "fused": orc._fused is not None}
    return r


def run_all(quick=False):
    """Every configuration, one dict each (bench.py puts them under ``extras.configs``)."""
    out = []
    for f in (cfg_a, cfg_b, lambda: cfg_d(quick), cfg_e):
        try:
            out.append(f())
        except Exception as exc:  # noqa: BLE001
            out.append({"error": repr(exc)[:300]})
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    for r in run_all("--quick" in sys.argv):
        print(json.dumps(r), flush=True)
