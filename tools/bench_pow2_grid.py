#!/usr/bin/env python
"""op / adj_op at matrix sizes whose 2N is not a power of two: a 3-D plan takes the next power of two as its grid
when that is at most a third larger per axis (csrc/api.cu), 2-D plans and larger ratios keep next235even + cuFFT.
One JSON line per size with the grid the plan chose."""
import json, sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/baseline/_ref")
import numpy as np, torch, mrinufft, mrinufft_b200
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 3)
def run(shape, M, C):
    rng = np.random.default_rng(0)
    d = len(shape)
    # radial-like density: gaussian cloud clipped
    traj = np.clip(rng.normal(0, 0.18, (M, d)), -0.499, 0.499).astype(np.float32)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, squeeze_dims=False)
    img = torch.randn((1, C, *shape), dtype=torch.complex64, device="cuda")
    ksp = torch.randn((1, C, M), dtype=torch.complex64, device="cuda")
    ref = mrinufft.get_operator("b200")(traj[:2000], shape, n_coils=1, squeeze_dims=False)
    print(json.dumps({"shape": list(shape), "nf": list(op.raw_op.plan.nf), "coils": C, "op_ms": timed(lambda: op.op(img)), "adj_op_ms": timed(lambda: op.adj_op(ksp))}), flush=True)
run((192,192,192), 1<<21, 8); run((225,225), 100000, 8); run((232,232,100), 1<<20, 4); run((160,160,160), 1<<21, 8); run((224,224,224), 1<<21, 8); run((320,320), 131072, 32); run((384,384), 200000, 8); run((192,192), 100000, 8)
