#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (configs[0], [1], [3]) on one GPU: wall-clock of the public
API calls with device-resident torch inputs (CUDA events), one JSON line per configuration.

    python tools/bench_configs.py [--quick]

cfg-A  README demo: 2-D radial 100 x 500, 512^2, single coil (density: pipe instead of voronoi)
cfg-B  2-D spiral 64 x 2048, 320^2, 32 coils with smaps: op / adj_op / data_consistency
cfg-D  3-D 256^3, 16 coils, density="pipe" + pinv_solver(optim="cg" | "lsqr", max_iter=10), and the
       reference's host-driven lsqr on the same operator for two iterations (what device residency saves)
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
    traj = initialize_2D_radial(100, 500).astype(np.float32)
    t0 = time.perf_counter()
    op = mrinufft.get_operator("b200")(traj, (512, 512), density="pipe")
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    img, ksp = crandn(512, 512), crandn(op.n_samples)
    return {"config": "A: 2D radial 100x500, 512^2, 1 coil, density=pipe", "setup_s_incl_pipe": setup,
            "op_ms": timed(lambda: op.op(img)), "adj_op_ms": timed(lambda: op.adj_op(ksp))}


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


def cfg_d(quick):
    n = 128 if quick else 256
    traj = initialize_3D_phyllotaxis_radial(4096 if quick else 16384, 512).astype(np.float32).reshape(-1, 3)
    C, shape = 16, (n, n, n)
    t0 = time.perf_counter()
    op_d = mrinufft.get_operator("b200")(traj, shape, n_coils=C, density="pipe", squeeze_dims=False)
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    # consistent data: k-space of a random 16-coil volume plus 1 % noise
    x_true = crandn(1, C, *shape)
    ksp = op_d._op_device(x_true)
    ksp += 0.01 * torch.linalg.norm(ksp) / np.sqrt(ksp.numel()) * crandn(*ksp.shape)
    # density-compensated adjoint as the starting point, then CG on the un-weighted normal equations.
    # (pinv_solver(optim="cg") on the density-weighted operator itself follows the reference statement by
    # statement -- extras/optim.py:832-842 takes the step size from the density-WEIGHTED operator and then
    # iterates on the un-weighted one -- and therefore diverges on a centre-heavy radial trajectory.)
    x0 = op_d._adj_device(ksp)
    del op_d
    torch.cuda.empty_cache()
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, squeeze_dims=False)
    res = []

    def cb(x, operator, y, **kw):
        res.append(float(torch.linalg.norm(operator.op(x) - y) / torch.linalg.norm(y)))

    def err(x):
        return float(torch.linalg.norm(x.reshape(x_true.shape) - x_true) / torch.linalg.norm(x_true))

    np.random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, _ = op.pinv_solver(ksp, optim="cg", max_iter=10, x_init=x0, callback=cb)
    torch.cuda.synchronize()
    cg_s = time.perf_counter() - t0
    out = {"config": f"D: 3D {n}^3, 16 coils (calibrationless), M={op.n_samples}, k-space of a random volume + 1 % noise",
           "setup_s_incl_pipe": setup, "cg_10_iterations_s_incl_callback": cg_s,
           "cg_rel_residual_first_last": [res[0], res[-1]], "cg_rel_image_error": err(x),
           "result_finite": bool(torch.isfinite(x).all())}
    del x
    # the reference's default optimiser, device resident (mrinufft_b200/solvers.py), from a zero start
    res.clear()
    x, _ = op.pinv_solver(ksp, optim="lsqr", max_iter=10, callback=cb)
    out["lsqr_rel_residual_first_last"] = [res[0], res[-1]]
    out["lsqr_rel_image_error"] = err(x)
    del x
    for name in ("lsqr", "lsmr"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.pinv_solver(ksp, optim=name, max_iter=10)
        torch.cuda.synchronize()
        out[f"{name}_10_iterations_s"] = time.perf_counter() - t0
    # the reference's lsqr (extras/optim.py:249-495) driving the same operator through host arrays
    from mrinufft.extras import get_optimizer

    ksp_h = ksp.cpu().numpy()
    t0 = time.perf_counter()
    get_optimizer("lsqr")(operator=op, kspace_data=ksp_h, max_iter=2, progressbar=False)
    out["reference_lsqr_host_driven_s_per_iteration"] = (time.perf_counter() - t0) / 2
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


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    for f in (cfg_a, cfg_b, lambda: cfg_d(quick), cfg_e):
        try:
            print(json.dumps(f()), flush=True)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"error": repr(exc)[:300]}), flush=True)
