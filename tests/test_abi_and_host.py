"""CPU tests: the C-ABI library loads and exports every symbol of include/b200nufft.h (no compute
calls without a GPU); backend registration; host-side logic."""

import ctypes
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_header_symbol():
    from mrinufft_b200 import _lib

    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 18 and "b200_type1" in syms and "b200_plan_setpts" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200nufft.h but not exported"
    assert set(syms) == set(_lib._SIGNATURES), "ctypes prototypes out of sync with the header"
    assert lib.b200_abi_version() == 1
    assert isinstance(lib.b200_last_error(), bytes)


def test_plan_create_rejects_bad_arguments_without_gpu():
    from mrinufft_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p(None)
    n = (ctypes.c_int64 * 3)(16, 16, 1)
    assert lib.b200_plan_create(ctypes.byref(h), 4, n, 1, 1e-6, 2.0, 0, 0) == -1   # dim = 4
    assert b"bad argument" in lib.b200_last_error()
    assert lib.b200_plan_create(ctypes.byref(h), 2, n, 0, 1e-6, 2.0, 0, 0) == -1   # n_trans = 0
    assert lib.b200_plan_destroy(None) == 0
    assert lib.b200_plan_info(None, None) == -1
    if not torch.cuda.is_available():
        assert lib.b200_plan_create(ctypes.byref(h), 2, n, 1, 1e-6, 2.0, 0, 0) == -2  # no device
        assert h.value is None


def test_backend_registers_in_mrinufft_registry():
    import mrinufft
    import mrinufft_b200
    from mrinufft.operators.base import FourierOperatorBase, check_backend, list_backends

    assert "b200" in list_backends(False)
    available, cls = FourierOperatorBase.interfaces["b200"]
    assert cls is mrinufft_b200.MRIB200NUFFT and cls.__name__ == "MRIB200NUFFT"
    assert cls.autograd_available and hasattr(cls, "pipe")
    assert check_backend("b200") == available
    if not torch.cuda.is_available():
        assert not available
        with pytest.raises(ValueError):
            mrinufft.get_operator("b200")
        # the product path fails loudly, it never falls back to a CPU implementation
        with pytest.raises(RuntimeError):
            cls(np.zeros((4, 2), np.float32), (8, 8))


def test_product_code_never_imports_the_oracle():
    import re

    for f in (ROOT / "mrinufft_b200").rglob("*.py"):
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f"{f} imports the oracle"
    for f in (ROOT / "mrinufft_b200" / "csrc").glob("*.c*"):
        assert "oracle/" not in f.read_text().replace("oracle/es_nufft.py::", "")


def test_cg_beta_follows_numpy_complex_ordering():
    from mrinufft_b200.solvers import _lex_max0

    for b in [0.3 + 0.9j, -0.3 + 0.9j, 0.5j, -0.5j, 1.0 + 0j, -1.0 + 0j, 0j]:
        assert _lex_max0(b) == max(0, np.complex128(b))


def test_coil_slices_partition():
    from mrinufft_b200.dist import coil_slice

    for C, W in [(32, 8), (32, 3), (7, 2), (4, 4)]:
        parts = [coil_slice(C, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == C
        assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        coil_slice(2, 0, 4)


def test_array_shim_round_trip_cpu():
    from mrinufft_b200._arrays import describe, from_device, to_device

    a = np.arange(6, dtype=np.complex128).reshape(2, 3)
    a.setflags(write=False)
    t = to_device(a, torch.device("cpu"), torch.complex64)
    assert t.dtype == torch.complex64 and t.shape == (2, 3)
    assert describe(a) == ("numpy", None)
    assert describe(torch.zeros(2))[0] == "torch"
    back = from_device(t, "numpy", None)
    assert isinstance(back, np.ndarray) and np.array_equal(back, a.astype(np.complex64))
    tc = torch.zeros(3, dtype=torch.complex64).conj()
    assert not to_device(tc, torch.device("cpu")).is_conj()


def test_bench_algorithmic_bytes_match_survey():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ab = bench.algorithmic_bytes(256**3, 512**3, 2**23, 3, 32)
    assert ab["per_transform_per_coil"] == 4_731_174_912
    assert abs(ab["pair_all_coils"] / 1e9 - 302.8) < 0.1
    # the per-kernel figures add up to the per-transform figure (coordinates/image counted once)
    parts = ab["pad"] + ab["fft"] + ab["interp"]
    assert abs(parts - 32 * ab["per_transform_per_coil"]) / parts < 0.06


def test_numa_binding_is_best_effort_without_a_gpu():
    """`bind_to_gpu_numa_node` never raises and leaves the affinity alone when the topology is not visible."""
    import os

    from mrinufft_b200.dist import bind_to_gpu_numa_node

    before = os.sched_getaffinity(0)
    res = bind_to_gpu_numa_node(0)
    assert res is None or (isinstance(res, tuple) and len(res) == 2)
    if res is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
