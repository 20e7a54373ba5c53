"""Array plumbing: numpy / torch (cpu, cuda) / cupy in -> torch CUDA tensors inside -> same type out.

Mirrors the contract of the reference's ``with_numpy`` / ``with_numpy_cupy`` decorators
(``src/mrinufft/_array_compat.py:183-227, 367-492``): the array module and device of the leading
argument decide the type of the result.  cupy is not required: anything exposing
``__cuda_array_interface__`` is wrapped zero-copy through ``torch.as_tensor``.
"""

from __future__ import annotations

import warnings
import weakref

import numpy as np
import torch

NP2TORCH = {
    np.dtype("float32"): torch.float32,
    np.dtype("float64"): torch.float64,
    np.dtype("complex64"): torch.complex64,
    np.dtype("complex128"): torch.complex128,
    np.dtype("int32"): torch.int32,
    np.dtype("int64"): torch.int64,
}


# ---------------------------------------------------------------- page-locking caller memory in place
# Large pageable numpy arrays cross PCIe through the driver's staging buffer at a fraction of the link
# speed and cannot overlap with compute.  Like the reference's GPU peer
# (``_host_register``, src/mrinufft/operators/interfaces/cufinufft.py:421-458) the caller's buffer is
# page-locked in place with ``cudaHostRegister`` -- no copy -- and the registration is cached by address
# for as long as the owning array lives (registering costs about as much as one staged copy; a
# reconstruction loop hands in the same buffers again and again).
PIN_MIN_BYTES = 8 << 20
_pinned = {}


def _owner(a: np.ndarray) -> np.ndarray:
    while isinstance(a.base, np.ndarray):
        a = a.base
    return a


def pin_in_place(a: np.ndarray) -> bool:
    """Page-lock the memory of the contiguous array ``a`` (cached); False if that is not possible."""
    if a.nbytes < PIN_MIN_BYTES or not a.flags.c_contiguous or not torch.cuda.is_available():
        return False
    key = (a.ctypes.data, a.nbytes)
    fin = _pinned.get(key)
    if fin is not None and fin.alive:
        return True
    try:
        if torch.from_numpy(a.view(np.uint8).reshape(-1)[:1]).is_pinned():
            return True  # already page-locked (e.g. a view of a pinned torch tensor)
    except Exception:  # noqa: BLE001  (read-only arrays warn, exotic dtypes raise: fall through)
        pass
    rt = torch.cuda.cudart()
    # ranges that overlap an older registration (a parent buffer, a stale entry) cannot be registered twice
    err = rt.cudaHostRegister(key[0], key[1], 0)
    if int(err) != 0:
        return False

    def _release(ptr=key[0], key=key):
        _pinned.pop(key, None)
        try:
            rt.cudaHostUnregister(ptr)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    try:
        _pinned[key] = weakref.finalize(_owner(a), _release)
    except TypeError:
        rt.cudaHostUnregister(key[0])
        return False
    return True


def module_name(arr) -> str:
    return type(arr).__module__.partition(".")[0]


def describe(arr):
    """(kind, device) of an input array; kind in {'numpy', 'torch', 'cupy'}."""
    root = module_name(arr)
    if root == "torch":
        return ("torch", arr.device)
    if root == "cupy":
        return ("cupy", int(arr.device.id))
    if isinstance(arr, np.ndarray) or np.isscalar(arr):
        return ("numpy", None)
    if hasattr(arr, "__cuda_array_interface__"):
        return ("cupy", None)
    if hasattr(arr, "__array__"):
        return ("numpy", None)
    raise ValueError(f"Unknown array library (={type(arr)}).")


def to_device(arr, device: torch.device, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Convert any supported array to a contiguous torch tensor on ``device`` (never mutates ``arr``).

    Host arrays that already live in page-locked memory (e.g. numpy views of pinned torch tensors)
    are copied asynchronously on the current stream; large pageable ones are page-locked in place
    first (``pin_in_place``), small ones go through the driver's staged copy.
    """
    root = module_name(arr)
    if root == "torch":
        t = arr
        if t.is_conj():
            t = t.resolve_conj()
        t = t.detach()
        t = t.to(device=device, dtype=dtype if dtype is not None else t.dtype, non_blocking=True)
        return t.contiguous()
    if isinstance(arr, np.ndarray) or not hasattr(arr, "__cuda_array_interface__"):
        a = np.asarray(arr)
        if dtype is not None:
            want = {v: k for k, v in NP2TORCH.items()}[dtype]
            a = a.astype(want, copy=False)
        a = np.ascontiguousarray(a)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)  # read-only inputs are never written to
            t = torch.from_numpy(a)
        if device.type == "cuda" and t.numel() > 0:
            # (a converted temporary is not worth a registration: only the caller's own buffer is pinned)
            own = isinstance(arr, np.ndarray) and a.ctypes.data == arr.ctypes.data
            if t.is_pinned() or (own and pin_in_place(a)):
                return t.to(device, non_blocking=True)
        return t.to(device)
    # __cuda_array_interface__ (cupy, numba, ...): zero-copy view
    t = torch.as_tensor(arr, device=device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def from_device(t: torch.Tensor, kind, device):
    """Convert a result tensor back to the leading argument's array type / device.

    numpy results are backed by page-locked memory from torch's caching host allocator, so the
    device->host copy runs at full PCIe speed and later calls can feed them back without staging.
    """
    if kind == "numpy":
        if t.is_cuda and t.numel() > 0:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            return host.numpy()
        return t.cpu().numpy()
    if kind == "torch":
        return t if t.device == device else t.to(device)
    if kind == "cupy":
        import cupy as cp  # only reachable when the caller handed us a cupy array

        return cp.from_dlpack(t)
    raise ValueError(kind)
