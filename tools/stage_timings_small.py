"""Secondary configurations A / B / E and the per-stage library timings of cfg-B (one GPU)."""
import sys, json
ROOT = __import__("pathlib").Path(__file__).resolve().parent.parent
for _p in (ROOT, ROOT / "baseline" / "_ref", ROOT / "tools"):
    sys.path.insert(0, str(_p))
import numpy as np, torch, mrinufft, mrinufft_b200
from bench_configs import cfg_a, cfg_b, cfg_e
from mrinufft.trajectories import initialize_2D_spiral
print(json.dumps(cfg_a())); print(json.dumps(cfg_b())); print(json.dumps(cfg_e()))
traj = initialize_2D_spiral(64, 2048, nb_revolutions=8).astype(np.float32)
C, shape = 32, (320, 320)
smaps = torch.view_as_complex(torch.randn(C, *shape, 2, device="cuda")); smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
img = torch.view_as_complex(torch.randn(1, 1, *shape, 2, device="cuda")); ksp = torch.view_as_complex(torch.randn(1, C, op.n_samples, 2, device="cuda"))
plan = op.raw_op.plan
for _ in range(3): op._op_device(img); op._adj_device(ksp)
plan.enable_timing(True)
op._op_device(img); t2 = plan.last_timings(); op._adj_device(ksp); t1 = plan.last_timings()
print("cfg-B stages type2", {k: round(v, 3) for k, v in t2.items()}); print("cfg-B stages type1", {k: round(v, 3) for k, v in t1.items()})
