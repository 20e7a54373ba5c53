"""cfg-B: is a small configuration bound by the host (enqueue time, cProfile) or by the device (events)?"""
import sys, time, cProfile, pstats, io
ROOT = __import__("pathlib").Path(__file__).resolve().parent.parent
for _p in (ROOT, ROOT / "baseline" / "_ref"):
    sys.path.insert(0, str(_p))
import numpy as np, torch, mrinufft, mrinufft_b200
from mrinufft.trajectories import initialize_2D_spiral
traj = initialize_2D_spiral(64, 2048, nb_revolutions=8).astype(np.float32)
C, shape = 32, (320, 320)
smaps = torch.view_as_complex(torch.randn(C, *shape, 2, device="cuda")); smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False)
img = torch.view_as_complex(torch.randn(1, 1, *shape, 2, device="cuda")); ksp = torch.view_as_complex(torch.randn(1, C, op.n_samples, 2, device="cuda"))
for _ in range(5): op.op(img); op.adj_op(ksp)
torch.cuda.synchronize()
# wall per call when only enqueueing (no sync) = host overhead; with sync each call = latency
def wall(fn, n=200, sync=False):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("op  enqueue-loop ms", wall(lambda: op.op(img)), " synced ms", wall(lambda: op.op(img), sync=True))
print("adj enqueue-loop ms", wall(lambda: op.adj_op(ksp)), " synced ms", wall(lambda: op.adj_op(ksp), sync=True))
# device time only: events around the raw library call
raw = op.raw_op; out = torch.empty((C, op.n_samples), dtype=torch.complex64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): raw.type2(img[0, 0], op._smaps_d, out, 1.0, False)
e1.record(); torch.cuda.synchronize(); print("raw type2 device ms", e0.elapsed_time(e1) / 50)
imgo = torch.empty(shape, dtype=torch.complex64, device="cuda")
torch.cuda.synchronize(); e0.record()
for _ in range(50): raw.type1(ksp[0], None, op._smaps_d, imgo, False, 1.0, False)
e1.record(); torch.cuda.synchronize(); print("raw type1 device ms", e0.elapsed_time(e1) / 50)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): op.op(img)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:2500])
