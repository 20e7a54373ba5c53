"""Host arrays spanning two coil chunks: pipelined copy streams against the sequential path and the device-resident one."""
# host-array op / adj_op with several chunks: pipelined vs sequential (coil_chunk forces 2 calls of 16 coils... and a 64-coil case)
import sys, time, json
ROOT = __import__("pathlib").Path(__file__).resolve().parent.parent
for _p in (ROOT, ROOT / "baseline" / "_ref"):
    sys.path.insert(0, str(_p))
import numpy as np, torch, mrinufft, mrinufft_b200
from mrinufft.trajectories import initialize_3D_phyllotaxis_radial
traj = initialize_3D_phyllotaxis_radial(4096, 512).astype(np.float32).reshape(-1, 3)
shape, C = (128, 128, 128), 64
smaps = torch.view_as_complex(torch.randn(C, *shape, 2, device="cuda"))
smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False, coil_chunk=32)
M = op.n_samples
x = torch.empty((1, 1, *shape), dtype=torch.complex64, pin_memory=True).normal_().numpy()
y = torch.empty((1, C, M), dtype=torch.complex64, pin_memory=True).normal_().numpy()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
res = {"config": f"3D 128^3, {C} coils in 2 chunks of 32, M={M}, host arrays"}
res["op_pipelined_ms"] = t(lambda: op.op(x)); res["adj_pipelined_ms"] = t(lambda: op.adj_op(y))
op._host_pipeline_applies = lambda arr: False
res["op_sequential_ms"] = t(lambda: op.op(x)); res["adj_sequential_ms"] = t(lambda: op.adj_op(y))
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
res["op_device_ms"] = t(lambda: op.op(xd)); res["adj_device_ms"] = t(lambda: op.adj_op(yd))
print(json.dumps(res))
