#!/usr/bin/env python
"""Row-kernel time against the coil count at the geometry of BASELINE configs[2] (256^3, M = 2^23 radial):
a call with T coils runs in coil class TC = next power of two >= T (rows_common.cuh).  Prints one JSON
line per T: spread / interp row kernel ms (CUDA events recorded by the library around the kernel), the
whole op / adj_op ms, visits of the class's stream, and the ratio to T = 32."""
import json
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    spokes, ns = (16384, 512) if n == 256 else (4096, 256)
    from mrinufft.trajectories import initialize_3D_phyllotaxis_radial

    traj = initialize_3D_phyllotaxis_radial(spokes, ns).reshape(-1, 3).astype(np.float32)
    M = traj.shape[0]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)

    def crandn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, device=dev, generator=g))

    base = None
    for T in (32, 16, 8, 4, 2, 1):
        op = mrinufft.get_operator("b200")(traj, (n,) * 3, n_coils=T, squeeze_dims=False, coil_chunk=T)
        plan = op.raw_op.plan
        img, ksp = crandn(1, T, n, n, n), crandn(1, T, M)
        for _ in range(2):
            op._op_device(img)
            op._adj_device(ksp)
        plan.enable_timing(True)
        r = {"interp_rows_ms": [], "spread_rows_ms": [], "interp_ms": [], "spread_ms": [], "fft_ms": []}
        for _ in range(3):
            op._op_device(img)
            t2 = plan.last_timings()
            op._adj_device(ksp)
            t1 = plan.last_timings()
            r["interp_rows_ms"].append(t2["rows_ms"])
            r["interp_ms"].append(t2["interp_ms"])
            r["spread_rows_ms"].append(t1["rows_ms"])
            r["spread_ms"].append(t1["spread_ms"])
            r["fft_ms"].append(t1["fft_ms"] + t2["fft_ms"])
        plan.enable_timing(False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            op._op_device(img)
            op._adj_device(ksp)
        e1.record()
        torch.cuda.synchronize()
        line = {k: round(float(np.mean(v)), 3) for k, v in r.items()}
        line.update(T=T, pair_ms=round(e0.elapsed_time(e1) / 3, 3), n=n, M=M, **plan.rows_class(T))
        if base is None:
            base = dict(line)
        line["spread_rows_vs_T32"] = round(line["spread_rows_ms"] / base["spread_rows_ms"], 3)
        line["interp_rows_vs_T32"] = round(line["interp_rows_ms"] / base["interp_rows_ms"], 3)
        line["pair_vs_T32"] = round(line["pair_ms"] / base["pair_ms"], 3)
        print(json.dumps(line), flush=True)
        del op, img, ksp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
