#!/usr/bin/env python
"""N-GPU == 1-GPU check of the coil-sharded operator on real hardware (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py

Every rank builds the full operator (all coils) and the coil-sharded one, and compares op / adj_op /
data_consistency and pinv_solver (cg, lsqr, lsmr; SENSE and calibrationless).  Rank 0 prints one JSON
line.  The CPU twin of this check (gloo, numpy stand-in operator) is tests/test_dist_cpu.py.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
for _p in (ROOT, ROOT / "baseline" / "_ref"):
    if _p.exists() and str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import mrinufft  # noqa: E402
import mrinufft_b200  # noqa: E402,F401
from mrinufft_b200.dist import CoilShardedOperator, coil_slice  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(a).cpu().numpy().ravel(), torch.as_tensor(b).cpu().numpy().ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    rng = np.random.default_rng(0)  # same data on every rank
    shape, M, C = (48, 40, 32), 20000, 8
    samples = rng.uniform(-np.pi, np.pi, (M, 3)).astype(np.float32)
    smaps = (rng.standard_normal((C, *shape)) + 1j * rng.standard_normal((C, *shape))).astype(np.complex64)
    smaps /= np.linalg.norm(smaps, axis=0)
    img = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
    imgs = (rng.standard_normal((1, C, *shape)) + 1j * rng.standard_normal((1, C, *shape))).astype(np.complex64)
    lo, hi = coil_slice(C, rank, world)
    out = {"world": world, "coils": C, "shape": shape, "M": M}
    worst = 0.0
    for sense in (True, False):
        full = mrinufft.get_operator("b200")(samples, shape, n_coils=C, smaps=smaps if sense else None,
                                             squeeze_dims=False)
        sh = CoilShardedOperator(samples, shape, C, smaps=smaps if sense else None)
        x = torch.from_numpy(img if sense else imgs).cuda()
        y_full = full.op(x)
        y = sh.op(x if sense else x[:, lo:hi].contiguous())
        errs = {"op": rel(y, y_full[:, lo:hi])}
        a_full = full.adj_op(y_full)
        a = sh.adj_op(y)
        errs["adj_op"] = rel(a, a_full if sense else a_full[:, lo:hi])
        # host arrays in, host arrays out: the sum over ranks still happens on the device
        a_h = sh.adj_op(y.cpu().numpy())
        assert isinstance(a_h, np.ndarray)
        errs["adj_op_host_arrays"] = rel(a_h, a_full if sense else a_full[:, lo:hi])
        g_full = full.data_consistency(x, 0.5 * y_full)
        g = sh.data_consistency(x if sense else x[:, lo:hi].contiguous(), 0.5 * y)
        errs["data_consistency"] = rel(g, g_full if sense else g_full[:, lo:hi])
        for name in ("cg", "lsqr", "lsmr"):
            kw = {"max_iter": 5}
            np.random.seed(3)
            want = full.pinv_solver(y_full, optim=name, **kw)
            np.random.seed(3 + rank)  # the sharded operator has to agree on rank 0's Lipschitz estimate
            got = sh.pinv_solver(y, optim=name, **kw)
            errs[f"pinv_{name}"] = rel(got, want if sense else want[:, lo:hi])
        out["sense" if sense else "calibrationless"] = errs
        worst = max(worst, max(errs.values()))
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["worst_rel_err_over_ranks"] = float(t.item())
    out["ok"] = bool(t.item() < 1e-4)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
