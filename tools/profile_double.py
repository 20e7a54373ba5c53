#!/usr/bin/env python
"""One op + one adj_op of a precision="double" operator (3-D 128^3, M = 2^21 radial, C coils with smaps), for ncu
launch lists / captures of the complex128 kernels.  usage: profile_double.py [C] [reps]"""
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
from mrinufft.trajectories import initialize_3D_phyllotaxis_radial  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = 128
traj = initialize_3D_phyllotaxis_radial(4096, 512).reshape(-1, 3).astype(np.float64)
smaps = torch.randn(C, n, n, n, dtype=torch.complex128, device="cuda")
op = mrinufft.get_operator("b200")(traj, (n,) * 3, n_coils=C, smaps=smaps, squeeze_dims=False, precision="double")
img = torch.randn(1, 1, n, n, n, dtype=torch.complex128, device="cuda")
ksp = torch.randn(1, C, op.n_samples, dtype=torch.complex128, device="cuda")
for _ in range(reps):
    op._op_device(img)
    op._adj_device(ksp)
torch.cuda.synchronize()
