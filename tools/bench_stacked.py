#!/usr/bin/env python
"""The stacked operator's z transform: the fused kernel (csrc/stack_fftz.cu) against the torch formulation it
replaced (smaps multiply, ifftshift, fft, fftshift, plane selection, permute + copy; adjoint: scatter, the same
chain backwards, conj-smaps coil sum), and the whole `stacked-b200` op / adj_op.  One JSON line per shape."""
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
from mrinufft_b200 import _lib  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 4)


def run(X, Y, Z, C, accel, M2):
    rng = np.random.default_rng(0)
    zsel = np.arange(0, Z, accel)
    NZ = len(zsel)
    zs = torch.as_tensor(zsel.astype(np.int32), device="cuda")
    zl = zs.long()
    sm = torch.randn((C, X, Y, Z), dtype=torch.complex64, device="cuda")
    img = torch.randn((1, X, Y, Z), dtype=torch.complex64, device="cuda")
    planes = torch.empty((C * NZ, X, Y), dtype=torch.complex64, device="cuda")
    out = torch.empty((1, X, Y, Z), dtype=torch.complex64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    sc = 1.0 / np.sqrt(2.0 * Z)

    def fwd_torch():
        coil = img * sm
        kz = torch.fft.fftshift(torch.fft.fft(torch.fft.ifftshift(coil, dim=-1), dim=-1, norm="ortho"), dim=-1) * 0.70710678
        return kz.index_select(-1, zl).permute(0, 3, 1, 2).reshape(C * NZ, X, Y).contiguous()

    def adj_torch():
        imgz = torch.zeros((C, X, Y, Z), dtype=torch.complex64, device="cuda")
        imgz.index_copy_(-1, zl, planes.reshape(C, NZ, X, Y).permute(0, 2, 3, 1))
        imgc = torch.fft.fftshift(torch.fft.ifft(torch.fft.ifftshift(imgz, dim=-1), dim=-1, norm="ortho"), dim=-1) * 0.70710678
        return torch.sum(imgc * torch.conj(sm), dim=0, keepdim=True)

    r = {"volume": [X, Y, Z], "coils": C, "planes": NZ,
         "fused_forward_ms": timed(lambda: _lib.stack_fftz(False, img.data_ptr(), sm.data_ptr(), planes.data_ptr(), zs.data_ptr(), C, X, Y, Z, NZ, sc, st)),
         "torch_forward_ms": timed(fwd_torch),
         "fused_adjoint_ms": timed(lambda: _lib.stack_fftz(True, planes.data_ptr(), sm.data_ptr(), out.data_ptr(), zs.data_ptr(), C, X, Y, Z, NZ, sc, st)),
         "torch_adjoint_ms": timed(adj_torch)}
    traj2d = rng.uniform(-0.5, 0.5, (M2, 2)).astype(np.float32)
    op = mrinufft.get_operator("stacked-b200")(traj2d, (X, Y, Z), smaps=sm, z_index=zsel, n_coils=C, squeeze_dims=False)
    ksp = torch.randn((1, C, NZ * M2), dtype=torch.complex64, device="cuda")
    x5 = img.reshape(1, 1, X, Y, Z)
    r["stacked_op_ms"] = timed(lambda: op.op(x5), 3)
    r["stacked_adj_op_ms"] = timed(lambda: op.adj_op(ksp), 3)
    print(json.dumps(r), flush=True)


run(256, 256, 128, 8, 2, 32768)
run(192, 192, 176, 8, 1, 32768)
run(320, 320, 64, 16, 2, 65536)
