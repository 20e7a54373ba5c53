#!/usr/bin/env python
"""Where the host-array (e2e) step of BASELINE configs[2] goes: op and adj_op through numpy arrays, timed
separately, for several host chunk sizes.  Round 2, one B200: chunks of 8 coils 66.7 / 60.3 ms, of 16
72.1 / 63.7, one chunk of 32 (no overlap) 81.1 / 77.6; cutting the last (op) / first (adj_op) chunk in two to
shorten the copy that nothing hides: 66.6 / 60.4, no gain -- the D2H / H2D chain is the critical path."""
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

n, C = 256, 32
traj = initialize_3D_phyllotaxis_radial(16384, 512).reshape(-1, 3).astype(np.float32)
M = traj.shape[0]
smaps = torch.view_as_complex(torch.randn(C, n, n, n, 2, device="cuda"))
img = torch.empty((1, 1, n, n, n), dtype=torch.complex64, pin_memory=True).normal_().numpy()
ksp = torch.empty((1, C, M), dtype=torch.complex64, pin_memory=True).normal_().numpy()


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for hc in (sys.argv[1:] or ["8", "16", "32"]):
    kw = {"host_chunk": int(hc)}
    op = mrinufft.get_operator("b200")(traj, (n,) * 3, n_coils=C, smaps=smaps, squeeze_dims=False, coil_chunk=C, **kw)
    print(json.dumps({"host_chunk": hc, "chunks": [b - a for a, b in op._chunks(host=True)],
                      "op_ms": t(lambda: op.op(img)), "adj_op_ms": t(lambda: op.adj_op(ksp))}), flush=True)
    del op
    torch.cuda.empty_cache()
