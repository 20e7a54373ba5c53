#!/usr/bin/env python
"""One warm-up and one profiled op + adj_op of BASELINE configs[2] (256^3, 32 coils with smaps, M = 2^23) for
ncu; plan options as key=value arguments (e.g. 2=3 for the TMA variant of the FFT passes)."""
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

opts = dict(kv.split("=") for kv in sys.argv[1:])
T = int(opts.pop("T", 32))
traj = initialize_3D_phyllotaxis_radial(16384, 512).reshape(-1, 3).astype(np.float32)
dev = torch.device("cuda", 0)
smaps = torch.view_as_complex(torch.randn(T, 256, 256, 256, 2, device=dev))
op = mrinufft.get_operator("b200")(traj, (256,) * 3, n_coils=T, smaps=smaps, squeeze_dims=False, coil_chunk=T)
for k, v in opts.items():
    op.raw_op.plan.set_option(int(k), int(v))
img = torch.view_as_complex(torch.randn(1, 1, 256, 256, 256, 2, device=dev))
ksp = torch.view_as_complex(torch.randn(1, T, traj.shape[0], 2, device=dev))
for _ in range(2):
    op._op_device(img)
    op._adj_device(ksp)
torch.cuda.synchronize()
