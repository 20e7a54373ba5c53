#!/usr/bin/env python
"""op / adj_op times of configurations whose grids go through the any-length FFT passes (csrc/fft_any.cu):
grids that are not powers of two and complex128 plans.  B200_FFT_METHOD=4 selects them (option key 2), the default is cuFFT.  One JSON line per configuration; `fft_ms` is the
library's own event timing of the FFT stage where the plan records it (single precision)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "baseline" / "_ref"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))
import mrinufft  # noqa: E402
import mrinufft_b200  # noqa: E402,F401


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(tag, shape, M, C, precision="single"):
    rng = np.random.default_rng(0)
    d = len(shape)
    traj = rng.uniform(-0.5, 0.5, (M, d)).astype(np.float64 if precision == "double" else np.float32)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, squeeze_dims=False, precision=precision)
    op.raw_op.plan.set_option(2, int(os.environ.get("B200_FFT_METHOD", "0")))  # 4: the own any-length passes
    cdt = torch.complex128 if precision == "double" else torch.complex64
    img = torch.randn((1, C, *shape), dtype=cdt, device="cuda")
    ksp = torch.randn((1, C, M), dtype=cdt, device="cuda")
    out = {"config": tag, "shape": list(shape), "M": M, "coils": C, "precision": precision,
           "op_ms": round(timed(lambda: op.op(img)), 4), "adj_op_ms": round(timed(lambda: op.adj_op(ksp)), 4)}
    if precision == "single":
        op.raw_op.plan.enable_timing(True)
        op.op(img)
        out["fft_ms_op"] = round(op.raw_op.plan.last_timings()["fft_ms"], 4)
    print(json.dumps(out), flush=True)
    del op
    torch.cuda.empty_cache()


run("cfg-B grid 640^2", (320, 320), 131072, 32)
run("2-D 384^2 -> 768^2", (384, 384), 200000, 8)
run("3-D 96^3 -> 192^3", (96, 96, 96), 1 << 20, 8)
run("3-D 192^3 -> 384^3", (192, 192, 192), 1 << 21, 8)
run("3-D 160x192x224", (160, 192, 224), 1 << 21, 4)
run("1-D 3000", (3000,), 50000, 4)
run("double 2-D 256^2", (256, 256), 131072, 8, "double")
run("double 3-D 128^3", (128, 128, 128), 1 << 21, 8, "double")
run("double 3-D 96^3", (96, 96, 96), 1 << 20, 8, "double")
