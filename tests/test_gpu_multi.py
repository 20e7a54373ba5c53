"""Two-GPU test of the coil-sharded operator on real hardware (NCCL over NVLink): N-GPU == 1-GPU for op,
adj_op (device and host arrays), data_consistency and the three solvers, SENSE and calibrationless.  Runs
`tools/dist_check.py` under torchrun; skipped on boxes with one GPU (run it with `gpurun --gpus 2`).  The
gloo / world-size-2 twin of this test on CPU is tests/test_dist_cpu.py."""

import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.timeout(600)
def test_coil_sharded_operator_on_two_gpus_equals_one_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / "dist_check.py")],
        capture_output=True, text=True, cwd=ROOT, timeout=560)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["world"] == 2 and out["worst_rel_err_over_ranks"] < 1e-4, out
    assert "adj_op_host_arrays" in out["sense"]


def test_one_process_two_devices():
    """Two operators on two devices of ONE process (`gpu_device_id`): per-device kernel attributes (dynamic
    shared-memory limits) and workspaces; both must give the single-device result."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import mrinufft
    import mrinufft_b200  # noqa: F401

    rng = np.random.default_rng(0)
    shape, M, C = (32, 32, 32), 20000, 4
    samples = rng.uniform(-0.5, 0.5, (M, 3)).astype(np.float32)
    img = (rng.standard_normal((1, C, *shape)) + 1j * rng.standard_normal((1, C, *shape))).astype(np.complex64)
    outs = []
    for dev in (0, 1):
        op = mrinufft.get_operator("b200")(samples, shape, n_coils=C, squeeze_dims=False, gpu_device_id=dev)
        y = op.op(img)
        outs.append((y, op.adj_op(y)))
    for a, b in zip(outs[0], outs[1]):
        assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(a)
