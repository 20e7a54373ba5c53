#!/usr/bin/env python
"""precision="double" against the single-precision path: 3-D 128^3, M = 2^21 radial, C coils with smaps
(op / adj_op ms, device resident), with the point-driven atomic spreader (option 3, bit 5) and the tile-owned row
spreader (csrc/double_rows.cu).  Round 2, one B200, 8 coils: single 3.0 / 2.8 ms; double 10.8 / 51.8 ms with
atomics (2 w^3 double atomics per point and coil, 2e11 / s at the L2), 10.8 / 12.6 ms with the row spreader,
9.4 / 12.6 ms with the interpolation gather in bin-sorted point order as well (option 3, bit 6 = caller's order).
Measured on the way and not kept: points sorted by 4 x 4 x 16-cell bricks under the atomic spreader (61 ms:
neighbouring lanes then collide on cells), lanes = coils (61 ms), brick tiles in shared memory with shared-memory
double atomics, i.e. the classic sub-problem spreader (108 .. 220 ms: those atomics are compare-and-swap loops on
sm_100a)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "baseline" / "_ref"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
import mrinufft  # noqa: E402
import mrinufft_b200  # noqa: E402,F401
from mrinufft.trajectories import initialize_3D_phyllotaxis_radial  # noqa: E402


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = 128
traj = initialize_3D_phyllotaxis_radial(4096, 512).reshape(-1, 3)
for prec, dbg, C in (("single", 0, 8), ("double", 32, 8), ("double", 64, 8), ("double", 0, 8), ("double", 0, 16), ("double", 0, 4)):
    cdt = torch.complex64 if prec == "single" else torch.complex128
    smaps = torch.randn(C, n, n, n, dtype=cdt, device="cuda")
    op = mrinufft.get_operator("b200")(traj.astype(np.float64 if prec == "double" else np.float32), (n,) * 3,
                                       n_coils=C, smaps=smaps, squeeze_dims=False, precision=prec)
    if dbg:  # option 3, bit 5: point-driven double spreader (atomics); bit 6: interpolation in the caller's point order
        op.raw_op.plan.set_option(3, dbg)
    img = torch.randn(1, 1, n, n, n, dtype=cdt, device="cuda")
    ksp = torch.randn(1, C, op.n_samples, dtype=cdt, device="cuda")
    t0 = time.perf_counter()
    op.raw_op._set_pts(op.samples)
    torch.cuda.synchronize()
    setpts_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"precision": prec, "spreader": "points" if dbg & 32 else "rows", "interp_order": "caller" if dbg else "sorted",
                      "op_ms": timed(lambda: op._op_device(img)), "adj_op_ms": timed(lambda: op._adj_device(ksp)),
                      "setpts_ms": setpts_ms, "M": op.n_samples, "coils": C, "n": n}), flush=True)
    del op, smaps, img, ksp
    torch.cuda.empty_cache()
