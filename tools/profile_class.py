#!/usr/bin/env python
"""One op + one adj_op with T coils at the geometry of BASELINE configs[2] (for ncu captures of one coil class)."""
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

T = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
spokes, ns = (16384, 512) if n == 256 else (4096, 256)
from mrinufft.trajectories import initialize_3D_phyllotaxis_radial  # noqa: E402

traj = initialize_3D_phyllotaxis_radial(spokes, ns).reshape(-1, 3).astype(np.float32)
dev = torch.device("cuda", 0)
op = mrinufft.get_operator("b200")(traj, (n,) * 3, n_coils=T, squeeze_dims=False, coil_chunk=T)
img = torch.view_as_complex(torch.randn(1, T, n, n, n, 2, device=dev))
ksp = torch.view_as_complex(torch.randn(1, T, traj.shape[0], 2, device=dev))
for _ in range(reps):
    op._op_device(img)
    op._adj_device(ksp)
torch.cuda.synchronize()
