"""``get_operator("b200")``: the B200-native NUFFT backend for mri-nufft.

Host-side mirror of the reference's backend interface for the type-2 ``op`` / type-1 ``adj_op`` hot
path.  It is modelled on

* ``MRIfinufft``      (``src/mrinufft/operators/interfaces/finufft.py:83-245``) -- the parity target:
  same constructor signature and defaults, ``samples`` kept as a host array in radians, ``pipe``;
* ``FourierOperatorSimple`` (``src/mrinufft/operators/base.py:880-1139``) -- shapes, squeeze rules,
  SENSE / calibrationless semantics, ``norm_factor`` on both op and adjoint, density before the
  adjoint only;
* ``MRICufiNUFFT``    (``src/mrinufft/operators/interfaces/cufinufft.py:140-1331``) -- device-resident
  smaps / density, ``n_trans`` must divide ``n_batchs * n_coils``.

All arithmetic runs in ``libb200nufft.so`` (hand-written sm_100a CUDA + cuFFT) through the C ABI in
``include/b200nufft.h``.  torch is used for device memory and streams only.  There is no CPU
fallback: without the library or without a GPU the backend registers as unavailable and the
constructor raises.
"""

from __future__ import annotations

import logging
import warnings
from contextlib import contextmanager

import numpy as np
import torch

from mrinufft._utils import proper_trajectory
from mrinufft.operators.base import FourierOperatorBase

from . import _lib
from ._arrays import describe, from_device, module_name, pin_in_place, to_device
from .toeplitz import assemble_toeplitz_kernel, modulated_weights

log = logging.getLogger("mrinufft_b200")

_IGNORED_FINUFFT_KWARGS = {
    "nthreads", "debug", "spread_debug", "showwarn", "spread_sort", "spread_kerevalmeth",
    "spread_kerpad", "chkbnds", "fftw", "modeord", "spread_thread", "maxbatchsize",
    "spread_nthr_atomic", "spread_max_sp_size", "allow_eps_too_small", "gpu_method", "gpu_sort",
    "gpu_kerevalmeth", "gpu_maxsubprobsize", "gpu_maxbatchsize", "gpu_spreadinterponly",
}


def _gpu_available() -> bool:
    try:
        return bool(_lib.library_built() and torch.cuda.is_available())
    except Exception:  # pragma: no cover
        return False


class RawB200Plan:
    """Device plan: one trajectory, type 1 and type 2, both signs (role of ``RawFinufftPlan``,
    ``finufft.py:23-80``, and ``RawCufinufftPlan``, ``cufinufft.py:51-137``)."""

    def __init__(self, samples, shape, n_trans=1, eps=1e-6, upsampfac=2.0, spread_only=False,
                 device=None, double=False, exact_grid=False):
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.eps = float(eps)
        self.upsampfac = float(upsampfac)
        self.n_trans = int(n_trans)
        self.spread_only = bool(spread_only)
        self.isign_flip = False  # toggle_grad_traj: e^{-i} <-> e^{+i}
        self.double = bool(double)
        self.plan = _lib.Plan(self.shape, n_trans_max=self.n_trans, eps=eps, upsampfac=upsampfac,
                              spread_only=spread_only, device=self.device.index, double=self.double,
                              exact_grid=exact_grid)
        self.n_samples = 0
        self._pts = None
        self.pts_version = 0  # bumped by every _set_pts (object ids are recycled, a counter is not)
        self._set_pts(samples)

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _set_pts(self, samples):
        """``Plan.setpts`` (finufft.py:55-62): fold + bin-sort on the device."""
        pts = to_device(samples, self.device, torch.float64 if self.double else torch.float32)
        if pts.ndim != 2 or pts.shape[1] != self.ndim:
            raise ValueError(f"samples should have shape (M, {self.ndim}), got {tuple(pts.shape)}")
        self._pts = pts
        self.pts_version += 1
        self.n_samples = int(pts.shape[0])
        self.plan.setpts(pts.data_ptr(), self.n_samples, self.stream)

    def toggle_grad_traj(self):
        """Role of ``RawFinufftPlan.toggle_grad_traj`` (finufft.py:78-80): the same plan serves the
        opposite sign, no second plan is needed."""
        self.isign_flip = not self.isign_flip

    def sort_indices(self):
        """(origin (d,M) int32, x1 (d,M) f32, key (M,) int32, perm (M,) int32) as numpy."""
        M, d = self.n_samples, self.ndim
        origin = torch.empty((d, M), dtype=torch.int32, device=self.device)
        x1 = torch.empty((d, M), dtype=torch.float32, device=self.device)
        key = torch.empty(M, dtype=torch.int32, device=self.device)
        perm = torch.empty(M, dtype=torch.int32, device=self.device)
        self.plan.get_sort(origin.data_ptr(), x1.data_ptr(), key.data_ptr(), perm.data_ptr(),
                           self.stream)
        torch.cuda.synchronize(self.device)
        return origin.cpu().numpy(), x1.cpu().numpy(), key.cpu().numpy(), perm.cpu().numpy()

    # raw transforms on device tensors -------------------------------------------------------
    def type2(self, img, smaps, ksp, scale=1.0, conj_smaps=False):
        T = ksp.shape[0]
        isign = +1 if self.isign_flip else -1
        self.plan.type2(img.data_ptr(), 0 if smaps is None else smaps.data_ptr(), ksp.data_ptr(),
                        T, isign, scale, int(conj_smaps), self.stream)

    def type1(self, ksp, density, smaps, img, accumulate=False, scale=1.0, conj_smaps=False):
        T = ksp.shape[0]
        isign = -1 if self.isign_flip else +1
        self.plan.type1(ksp.data_ptr(), 0 if density is None else density.data_ptr(),
                        0 if smaps is None else smaps.data_ptr(), img.data_ptr(), T,
                        int(accumulate), isign, scale, int(conj_smaps), self.stream)


class MRIB200NUFFT(FourierOperatorBase):
    """MRI non-Cartesian Fourier operator running on one NVIDIA B200.

    Parameters
    ----------
    samples: array
        Sample locations ``(M, d)`` (or ``(Nc, Ns, d)``), in [-0.5, 0.5) or radians
        (rescaled like ``proper_trajectory(..., "pi")``).
    shape: tuple
        Image shape (2D or 3D).
    density: bool, array, str, dict or callable
        Density compensation (same grammar as ``FourierOperatorBase.compute_density``).
    n_coils, n_batchs, n_trans, smaps, squeeze_dims:
        As for ``MRIfinufft`` (finufft.py:116-127).  ``n_trans`` is the number of transforms
        batched in one device call and must divide ``n_batchs * n_coils``; with the default
        ``n_trans=1`` the batch size is chosen automatically from the free device memory
        (results do not depend on it).
    eps: float
        Requested tolerance (default 1e-6 -> kernel width 7 at upsampfac 2).
    upsampfac: float
        Oversampling factor sigma (default 2.0).
    gpu_device_id: int, optional
        CUDA device ordinal (default: current torch device).
    """

    backend = "b200"
    available = _gpu_available()
    autograd_available = True

    def __init__(
        self,
        samples,
        shape,
        density=False,
        n_coils=1,
        n_batchs=1,
        n_trans=1,
        smaps=None,
        squeeze_dims=True,
        eps=1e-6,
        upsampfac=2.0,
        gpu_device_id=None,
        coil_chunk=None,
        host_chunk=None,
        spreadinterponly=0,
        isign=None,
        precision="single",
        exact_grid=False,
        **kwargs,
    ):
        if not _lib.library_built():
            raise RuntimeError(
                f"'{self.backend}' backend is not available: {_lib.LIB_PATH} has not been built."
            )
        if not torch.cuda.is_available():
            raise RuntimeError(f"'{self.backend}' backend is not available: no CUDA device.")
        unknown = set(kwargs) - _IGNORED_FINUFFT_KWARGS
        if unknown:
            raise TypeError(f"MRIB200NUFFT got unexpected keyword arguments {sorted(unknown)}")
        super().__init__()
        self._smaps_d = None
        self._density_d = None
        self.shape = shape
        if len(self.shape) not in (1, 2, 3):
            raise ValueError("b200 backend supports 1D, 2D and 3D transforms")
        dev_index = torch.cuda.current_device() if gpu_device_id is None else int(gpu_device_id)
        self.device = torch.device("cuda", dev_index)

        if precision not in ("single", "double"):
            raise ValueError("precision should be 'single' or 'double'")
        # single precision is the native, tuned path (cufinufft precedent: float64 samples are cast,
        # cufinufft.py:338-340); precision="double" opts into the complex128 kernels
        self._double = precision == "double"
        if self._double and spreadinterponly:
            raise ValueError("precision='double' does not support spreadinterponly plans")
        self._rdt = torch.float64 if self._double else torch.float32
        self._cdt = torch.complex128 if self._double else torch.complex64
        samples = self._normalize_samples(samples)
        self._samples = samples
        self.dtype = np.float64 if self._double else np.float32
        if n_coils < 1:
            raise ValueError("n_coils should be ≥ 1")
        self.n_coils = n_coils
        self.n_batchs = n_batchs
        self.n_trans = int(n_trans)
        if (self.n_batchs * self.n_coils) % self.n_trans != 0:
            raise ValueError("n_batchs * n_coils should be a multiple of n_transf")
        self.squeeze_dims = squeeze_dims
        self.eps = float(eps)
        self.upsampfac = float(upsampfac) if upsampfac else 2.0
        self._spread_only = bool(spreadinterponly)
        self._exact_grid = bool(exact_grid)  # keep finufft's next235even(sigma N) grid (B200_EXACT_GRID)
        self._user_isign_flip = isign is not None and int(isign) > 0
        self._conj_smaps = False
        self._coil_chunk = coil_chunk
        self._host_chunk = int(host_chunk) if host_chunk else None

        self.raw_op = RawB200Plan(
            samples, self.shape, n_trans=self._pick_chunk(), eps=self.eps, upsampfac=self.upsampfac,
            spread_only=self._spread_only, device=dev_index, double=self._double, exact_grid=self._exact_grid,
        )
        if self._user_isign_flip:
            self.raw_op.toggle_grad_traj()
        # Density compensation, then multi coil setup (order of base.py:942-945).
        self.compute_density(density)
        self.compute_smaps(smaps)

    # ------------------------------------------------------------------ helpers
    def _normalize_samples(self, samples):
        if module_name(samples) == "torch":
            samples = samples.detach().cpu().numpy()
        elif not isinstance(samples, np.ndarray):
            if hasattr(samples, "get"):
                samples = samples.get()
            samples = np.asarray(samples)
        samples = proper_trajectory(samples, normalize="pi")
        want = np.float64 if getattr(self, "_double", False) else np.float32
        if samples.dtype != want:
            if samples.dtype == np.float64:
                self.log.info("b200 backend computes in float32/complex64 unless precision='double': "
                              "casting float64 samples.")
            samples = samples.astype(want)
        return np.ascontiguousarray(samples.reshape(-1, len(self.shape)))

    def _pick_chunk(self) -> int:
        """Number of transforms batched per device call (workspace = chunk oversampled grids)."""
        total = self.n_batchs * self.n_coils
        if self._coil_chunk:
            return max(1, min(int(self._coil_chunk), self.n_coils))
        if self.n_trans > 1:
            return self.n_trans
        if self._spread_only:
            return min(self.n_coils, 8)
        grid_bytes = (16 if self._double else 8) * float(np.prod(
            _lib.grid_size(self.shape, self.eps, self.upsampfac, self._double, self._exact_grid)))
        free, _ = torch.cuda.mem_get_info(self.device)
        # grids + cuFFT work area (same order) may take at most ~45% of the free memory
        cap = int(0.45 * free / (2.0 * grid_bytes))
        cap = max(1, min(cap, 32))
        # largest chunk <= cap that divides n_coils (keeps chunks inside one batch volume)
        best = 1
        for c in range(1, min(cap, self.n_coils) + 1):
            if self.n_coils % c == 0:
                best = c
        _ = total
        return best

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _in(self, arr, dtype=None):
        kind, dev = describe(arr)
        t = to_device(arr, self.device, self._cdt if dtype is None else dtype)
        return t, kind, dev

    def _out(self, t, kind, dev):
        return from_device(t, kind, dev)

    # ------------------------------------------------------------------ smaps / density / samples
    @property
    def smaps(self):
        """Sensitivity maps (as given; a complex64 copy lives on the device)."""
        if getattr(self, "_smaps_lazy_host", False) and self._smaps is not None:
            self._smaps = self._smaps.cpu().numpy()
            self._smaps_lazy_host = False
        return self._smaps

    @smaps.setter
    def smaps(self, new_smaps):
        # accepts numpy, torch (cpu/cuda) and cupy arrays; the base setter only takes numpy
        # (base.py:759-760) -- same override as cufinufft.py:277-315.
        self._smaps_lazy_host = False
        self._smaps_version = getattr(self, "_smaps_version", 0) + 1
        if new_smaps is None:
            self._smaps = None
            self._smaps_d = None
            return
        shape = tuple(new_smaps.shape)
        C, XYZ = shape[0], shape[1:]
        if XYZ != self.shape:
            raise ValueError("Smaps should match image shape.")
        if C != self._n_coils:
            self._n_coils = int(C)
            self.log.warning("updating number of coils via Smaps.")
        self._smaps = new_smaps
        # a contiguous complex64 CUDA tensor on this device is used zero-copy (never written to)
        self._smaps_d = to_device(new_smaps, self.device, self._cdt)

    def compute_smaps(self, method=None):
        """Same grammar as ``FourierOperatorBase.compute_smaps`` (base.py:407-448) + device arrays."""
        if method is not None and not isinstance(method, (str, dict)) and not callable(method):
            if hasattr(method, "shape"):
                self.smaps = method.reshape(self.n_coils, *self.shape)
                return
        if method is None:
            self.smaps = None
            return
        name, kw = None, {}
        if isinstance(method, str):
            name = method
        elif isinstance(method, dict) and isinstance(method.get("name"), str):
            name = method["name"]
            kw = {k: v for k, v in method.items() if k != "name"}
        if name == "low_frequency" and not np.any(kw.get("mask", False)) and np.sum(kw.get("blurr_factor", 0.0)) == 0:
            self.smaps = self._low_frequency_smaps(**kw)
            self._smaps_lazy_host = True  # `.smaps` hands out numpy, like the reference; copied on first use
            return
        super().compute_smaps(method)

    def _low_frequency_smaps(self, kspace_data, threshold=0.1, max_iter=10, window_fun="ellipse",
                             mask=False, blurr_factor=0.0):
        """Low-frequency sensitivity maps (``low_frequency``, extras/smaps.py:220-306) for the default
        ``mask=False, blurr_factor=0``: centre of k-space -> ``pinv_solver`` on a throw-away
        calibrationless operator (device-resident lsqr) -> divide by the root sum of squares, all on
        the device.  The reference function imports scikit-image up front, even when neither the mask
        nor the blur is requested; those two options still go through it.
        """
        from mrinufft.extras.smaps import _extract_kspace_center

        ksp = kspace_data
        if module_name(ksp) == "torch":
            ksp = ksp.detach().cpu().numpy()
        elif not isinstance(ksp, np.ndarray):
            ksp = ksp.get() if hasattr(ksp, "get") else np.asarray(ksp)
        k_space, samples, _ = _extract_kspace_center(
            kspace_data=ksp, kspace_loc=self.samples, threshold=threshold, density=self.density,
            window_fun=window_fun,
        )
        # same constructor call as extras/smaps.py:280-285 (the centre samples go through
        # proper_trajectory again, exactly as they do there)
        centre = type(self)(samples, self.shape, n_coils=k_space.shape[-2], squeeze_dims=True,
                            gpu_device_id=self.device.index, precision="double" if self._double else "single")
        maps = centre.pinv_solver(to_device(k_space, self.device, self._cdt), max_iter=max_iter)
        maps = maps.reshape(k_space.shape[-2], *self.shape)
        sos = torch.sqrt(torch.sum(maps.real ** 2 + maps.imag ** 2, dim=0, keepdim=True))
        return maps / sos

    @property
    def density(self):
        """Density compensation weights (host numpy float32, or ``None``)."""
        return self._density

    @density.setter
    def density(self, new_density):
        self._toeplitz_kernel = None
        if new_density is None:
            self._density = None
            self._density_d = None
            return
        if len(new_density) != self.n_samples:
            raise ValueError("Density and samples should have the same length")
        d = to_device(new_density, self.device)
        if d.is_complex():
            d = d.real
        d = d.to(self._rdt).contiguous().reshape(-1)
        self._density_d = d
        self._density = new_density if isinstance(new_density, np.ndarray) else d.cpu().numpy()

    def compute_density(self, method=None):
        """``FourierOperatorBase.compute_density`` (base.py:576-624) with device-array support."""
        if method is not None and not isinstance(method, (str, dict, bool)) and not callable(method):
            if hasattr(method, "shape"):
                self.density = method
                return
        if method is None or method is False:
            self.density = None
            return
        super().compute_density(method)
        # the base class may have stored a raw array without going through the setter
        if self._density is not None and self._density_d is None:
            self.density = self._density

    @property
    def samples(self):
        """Host numpy ``(M, d)`` float32 array in radians (as ``MRIfinufft``, finufft.py:128-129)."""
        return self._samples

    @samples.setter
    def samples(self, new_samples):
        self.update_samples(new_samples)

    def update_samples(self, new_samples, *, unsafe: bool = False):
        """Update the sample locations (``MRIfinufft.update_samples``, finufft.py:150-181)."""
        if unsafe and module_name(new_samples) == "torch" and new_samples.is_cuda:
            self.raw_op._set_pts(new_samples.detach())
            self._samples = new_samples.detach().cpu().numpy()
        else:
            self._samples = self._normalize_samples(new_samples)
            self.raw_op._set_pts(self._samples)
        # like the reference (finufft.py:181): `compute_density(self._density_method)` unconditionally, so a
        # density given as an array (no method recorded) does not survive new sample locations -- its
        # weights belong to the old ones, and its length may no longer match
        self.compute_density(self._density_method)
        self._toeplitz_kernel = None

    # ------------------------------------------------------------------ toggles for the trajectory gradient
    def toggle_grad_traj(self):
        """Role of ``_ToggleGradPlanMixin.toggle_grad_traj`` (base.py:1234-1238): opposite FFT sign
        and conjugated smaps, without copying the maps or building a second plan."""
        if self.uses_sense:
            self._conj_smaps = not self._conj_smaps
        self.raw_op.toggle_grad_traj()

    @contextmanager
    def grad_traj_plan(self):
        self.toggle_grad_traj()
        try:
            yield
        finally:
            self.toggle_grad_traj()

    def _make_plan_grad(self, **kwargs):
        """Nothing to build: the plan handles both signs (cf. finufft.py:183-193)."""

    # ------------------------------------------------------------------ device-level transforms
    def _chunks(self, host=False):
        """Coil ranges of the library calls.  Device arrays: as many coils per call as the workspace holds.
        Host arrays (``host=True``): smaller chunks, so that the k-space of one chunk crosses PCIe while
        the next one is transformed -- the row kernels run a call of T coils in a coil class whose cost
        follows T (rows_common.cuh), so four calls of 8 coils cost little more than one call of 32."""
        T = self.raw_op.n_trans
        C = self.n_coils
        if host and not self._spread_only:
            T = min(T, self._host_chunk if self._host_chunk else (8 if C >= 16 else T))
        return [(c0, min(c0 + T, C)) for c0 in range(0, C, T)]

    def _op_device(self, img: torch.Tensor, ksp: torch.Tensor | None = None) -> torch.Tensor:
        """img (B, 1|C, *XYZ) complex64 on device -> ksp (B, C, K)."""
        B, C, K, XYZ = self.n_batchs, self.n_coils, self.n_samples, self.shape
        # the C ABI carries the scale as a float: double plans get 1 and are scaled here in float64
        inv_norm = 1.0 if self._double else float(self.inv_norm_factor)
        if ksp is None:
            ksp = torch.empty((B, C, K), dtype=self._cdt, device=self.device)
        ksp = ksp.view(B, C, K)
        raw = self.raw_op
        if self._spread_only:
            img = img.reshape(B, C, *XYZ)
            for b in range(B):
                for c0, c1 in self._chunks():
                    raw.plan.interp(img[b, c0:c1].data_ptr(), ksp[b, c0:c1].data_ptr(), c1 - c0,
                                    self.stream)
            ksp *= inv_norm
            return ksp
        if self.uses_sense:
            img = img.reshape(B, *XYZ)
            for b in range(B):
                for c0, c1 in self._chunks():
                    raw.type2(img[b], self._smaps_d[c0:c1], ksp[b, c0:c1], inv_norm, self._conj_smaps)
        else:
            img = img.reshape(B, C, *XYZ)
            for b in range(B):
                for c0, c1 in self._chunks():
                    raw.type2(img[b, c0:c1], None, ksp[b, c0:c1], inv_norm)
        if self._double:
            ksp *= float(self.inv_norm_factor)
        return ksp

    def _adj_device(self, ksp: torch.Tensor, img: torch.Tensor | None = None) -> torch.Tensor:
        """ksp (B, C, K) complex64 on device -> img (B, 1|C, *XYZ)."""
        B, C, K, XYZ = self.n_batchs, self.n_coils, self.n_samples, self.shape
        inv_norm = 1.0 if self._double else float(self.inv_norm_factor)
        ksp = ksp.reshape(B, C, K)
        raw = self.raw_op
        dens = self._density_d
        if self._spread_only:
            if img is None:
                img = torch.empty((B, C, *XYZ), dtype=self._cdt, device=self.device)
            img = img.view(B, C, *XYZ)
            for b in range(B):
                for c0, c1 in self._chunks():
                    k = ksp[b, c0:c1]
                    if dens is not None:
                        k = k * dens
                    raw.plan.spread(k.data_ptr(), img[b, c0:c1].data_ptr(), c1 - c0, self.stream)
            img *= inv_norm
            return img
        if self.uses_sense:
            if img is None:
                img = torch.empty((B, 1, *XYZ), dtype=self._cdt, device=self.device)
            img = img.view(B, 1, *XYZ)
            for b in range(B):
                for i, (c0, c1) in enumerate(self._chunks()):
                    raw.type1(ksp[b, c0:c1], dens, self._smaps_d[c0:c1], img[b, 0], accumulate=i > 0,
                              scale=inv_norm, conj_smaps=self._conj_smaps)
        else:
            if img is None:
                img = torch.empty((B, C, *XYZ), dtype=self._cdt, device=self.device)
            img = img.view(B, C, *XYZ)
            for b in range(B):
                for c0, c1 in self._chunks():
                    raw.type1(ksp[b, c0:c1], dens, None, img[b, c0:c1], accumulate=False,
                              scale=inv_norm)
        if self._double:
            img *= float(self.inv_norm_factor)
        return img

    def _dc_device(self, img: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
        """Fused A^H(Ax - y) (base.py:1075-1139), the k-space residual stays inside the library."""
        B, C, K, XYZ = self.n_batchs, self.n_coils, self.n_samples, self.shape
        inv_norm = float(self.inv_norm_factor)
        obs = obs.reshape(B, C, K)
        raw = self.raw_op
        dens = self._density_d
        dptr = 0 if dens is None else dens.data_ptr()
        if self._spread_only or raw.isign_flip or self._conj_smaps or self._double:
            return self._adj_device(self._op_device(img) - obs)
        if self.uses_sense:
            img = img.reshape(B, *XYZ)
            grad = torch.empty((B, 1, *XYZ), dtype=self._cdt, device=self.device)
            for b in range(B):
                for i, (c0, c1) in enumerate(self._chunks()):
                    raw.plan.data_consistency(
                        img[b].data_ptr(), self._smaps_d[c0:c1].data_ptr(), obs[b, c0:c1].data_ptr(),
                        dptr, grad[b, 0].data_ptr(), c1 - c0, int(i > 0), inv_norm, self.stream)
        else:
            img = img.reshape(B, C, *XYZ)
            grad = torch.empty((B, C, *XYZ), dtype=self._cdt, device=self.device)
            for b in range(B):
                for c0, c1 in self._chunks():
                    raw.plan.data_consistency(
                        img[b, c0:c1].data_ptr(), 0, obs[b, c0:c1].data_ptr(), dptr,
                        grad[b, c0:c1].data_ptr(), c1 - c0, 0, inv_norm, self.stream)
        return grad

    # ------------------------------------------------------------------ host arrays, several device calls
    # When a batch of host (numpy) data runs as several library calls -- more coils than fit one
    # workspace, or several batch volumes -- the k-space chunks cross PCIe on dedicated copy streams
    # while the previous / next chunk is being transformed: the role of cufinufft's `async_transfer`
    # pipelines (cufinufft.py:538-606, 647-700, 816-900), always on.  A single call (the 32 coils of one
    # volume fit one workspace on a B200) has nothing to overlap with and takes the plain path.
    def _host_pipeline_applies(self, arr) -> bool:
        if not isinstance(arr, np.ndarray) or self._spread_only:
            return False
        return self.n_batchs * len(self._chunks(host=True)) > 1

    def _copy_streams(self):
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(self.device)
            self._d2h_stream = torch.cuda.Stream(self.device)
        return self._h2d_stream, self._d2h_stream

    @staticmethod
    def _host_tensor(arr, dtype):
        a = np.ascontiguousarray(np.asarray(arr).astype(
            {torch.complex64: np.complex64, torch.complex128: np.complex128}[dtype], copy=False))
        if isinstance(arr, np.ndarray) and a.ctypes.data == arr.ctypes.data:
            pin_in_place(a)  # pageable caller memory: page-locked in place (cached), copies become asynchronous
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)  # read-only inputs are never written to
            return torch.from_numpy(a)

    def _op_host_pipelined(self, data) -> np.ndarray:
        """``op`` on a host array: type 2 per chunk, device -> host copies of the finished chunk's
        k-space overlapped with the next chunk's transform (calibrationless: its image chunk comes in
        on the other copy stream meanwhile).  Returns the full ``(B, C, K)`` numpy array."""
        B, C, K, XYZ = self.n_batchs, self.n_coils, self.n_samples, self.shape
        dev, cdt, raw = self.device, self._cdt, self.raw_op
        inv_norm = 1.0 if self._double else float(self.inv_norm_factor)
        cur = torch.cuda.current_stream(dev)
        h2d, d2h = self._copy_streams()
        h2d.wait_stream(cur)   # the buffers below may be recycled blocks with work pending on `cur`
        d2h.wait_stream(cur)
        T = max(c1 - c0 for c0, c1 in self._chunks(host=True))
        src = self._host_tensor(data, cdt)
        out_h = torch.empty((B, C, K), dtype=cdt, pin_memory=True)
        kbuf = [torch.empty((T, K), dtype=cdt, device=dev) for _ in range(2)]
        if self.uses_sense:
            img_d = src.reshape(B, *XYZ).to(dev, non_blocking=True)
        else:
            src = src.reshape(B, C, *XYZ)
            ibuf = [torch.empty((T, *XYZ), dtype=cdt, device=dev) for _ in range(2)]
        d2h_done, comp_done = [None, None], [None, None]
        i = 0
        for b in range(B):
            for c0, c1 in self._chunks(host=True):
                k, n = i % 2, c1 - c0
                if d2h_done[k] is not None:
                    cur.wait_event(d2h_done[k])          # kbuf[k] has been copied out
                if self.uses_sense:
                    raw.type2(img_d[b], self._smaps_d[c0:c1], kbuf[k][:n], inv_norm, self._conj_smaps)
                else:
                    with torch.cuda.stream(h2d):
                        if comp_done[k] is not None:
                            h2d.wait_event(comp_done[k])  # ibuf[k] has been consumed
                        ibuf[k][:n].copy_(src[b, c0:c1], non_blocking=True)
                        arrived = torch.cuda.Event()
                        arrived.record(h2d)
                    cur.wait_event(arrived)
                    raw.type2(ibuf[k][:n], None, kbuf[k][:n], inv_norm)
                if self._double:
                    kbuf[k][:n] *= float(self.inv_norm_factor)
                comp_done[k] = torch.cuda.Event()
                comp_done[k].record(cur)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(comp_done[k])
                    out_h[b, c0:c1].copy_(kbuf[k][:n], non_blocking=True)
                    d2h_done[k] = torch.cuda.Event()
                    d2h_done[k].record(d2h)
                i += 1
        d2h.synchronize()
        for t in kbuf:
            t.record_stream(d2h)
        if not self.uses_sense:
            for t in ibuf:
                t.record_stream(h2d)
        return out_h.numpy()

    def _adj_host_pipelined(self, coeffs, keep_on_device=False):
        """``adj_op`` on a host array: the next chunk's k-space comes in on a copy stream while the
        current chunk is transformed; SENSE accumulates the coil-combined image on the device,
        calibrationless images leave chunk by chunk on the other copy stream."""
        B, C, K, XYZ = self.n_batchs, self.n_coils, self.n_samples, self.shape
        dev, cdt, raw = self.device, self._cdt, self.raw_op
        inv_norm = 1.0 if self._double else float(self.inv_norm_factor)
        cur = torch.cuda.current_stream(dev)
        h2d, d2h = self._copy_streams()
        h2d.wait_stream(cur)
        d2h.wait_stream(cur)
        T = max(c1 - c0 for c0, c1 in self._chunks(host=True))
        src = self._host_tensor(coeffs, cdt).reshape(B, C, K)
        kbuf = [torch.empty((T, K), dtype=cdt, device=dev) for _ in range(2)]
        sense = self.uses_sense
        if sense:
            img_d = torch.empty((B, 1, *XYZ), dtype=cdt, device=dev)
        else:
            out_h = torch.empty((B, C, *XYZ), dtype=cdt, pin_memory=True)
            obuf = [torch.empty((T, *XYZ), dtype=cdt, device=dev) for _ in range(2)]
        comp_done, d2h_done = [None, None], [None, None]
        i = 0
        for b in range(B):
            for j, (c0, c1) in enumerate(self._chunks(host=True)):
                k, n = i % 2, c1 - c0
                with torch.cuda.stream(h2d):
                    if comp_done[k] is not None:
                        h2d.wait_event(comp_done[k])      # kbuf[k] has been consumed
                    kbuf[k][:n].copy_(src[b, c0:c1], non_blocking=True)
                    arrived = torch.cuda.Event()
                    arrived.record(h2d)
                cur.wait_event(arrived)
                if sense:
                    raw.type1(kbuf[k][:n], self._density_d, self._smaps_d[c0:c1], img_d[b, 0],
                              accumulate=j > 0, scale=inv_norm, conj_smaps=self._conj_smaps)
                else:
                    if d2h_done[k] is not None:
                        cur.wait_event(d2h_done[k])       # obuf[k] has been copied out
                    raw.type1(kbuf[k][:n], self._density_d, None, obuf[k][:n], accumulate=False,
                              scale=inv_norm)
                    if self._double:
                        obuf[k][:n] *= float(self.inv_norm_factor)
                comp_done[k] = torch.cuda.Event()
                comp_done[k].record(cur)
                if not sense:
                    with torch.cuda.stream(d2h):
                        d2h.wait_event(comp_done[k])
                        out_h[b, c0:c1].copy_(obuf[k][:n], non_blocking=True)
                        d2h_done[k] = torch.cuda.Event()
                        d2h_done[k].record(d2h)
                i += 1
        for t in kbuf:
            t.record_stream(h2d)
        if sense:
            if self._double:
                img_d *= float(self.inv_norm_factor)
            if keep_on_device:  # (the coil-sharded operator sums over ranks before the image leaves the device)
                return img_d
            return from_device(img_d, "numpy", None)
        d2h.synchronize()
        for t in obuf:
            t.record_stream(d2h)
        return out_h.numpy()

    # ------------------------------------------------------------------ public API
    def op(self, data, ksp=None):
        r"""Non-Cartesian MRI forward operator :math:`\mathcal{F}\mathcal{S}_\ell x` (base.py:949-977).

        Accepts numpy / torch / cupy arrays; the result has the type and device of ``data``.
        """
        self.check_shape(image=data, ksp=ksp)
        if ksp is None and self._host_pipeline_applies(data):
            return self._safe_squeeze(self._op_host_pipelined(data))
        img, kind, dev = self._in(data)
        out = None
        if ksp is not None and module_name(ksp) == "torch" and ksp.is_cuda and ksp.is_contiguous() \
                and ksp.dtype == self._cdt and ksp.device == self.device:
            out = ksp
        res = self._op_device(img, out)
        if out is not None:
            return ksp  # written in place: hand the caller's buffer back, whatever the input type was
        res = self._safe_squeeze(res)
        if ksp is not None:
            _copy_into(ksp, res)
            return ksp
        return self._out(res, kind, dev)

    def adj_op(self, coeffs, img=None):
        """Non-Cartesian MRI adjoint operator (base.py:1015-1035)."""
        self.check_shape(image=img, ksp=coeffs)
        if img is None and self._host_pipeline_applies(coeffs):
            return self._safe_squeeze(self._adj_host_pipelined(coeffs))
        ksp, kind, dev = self._in(coeffs)
        out = None
        if img is not None and module_name(img) == "torch" and img.is_cuda and img.is_contiguous() \
                and img.dtype == self._cdt and img.device == self.device:
            out = img
        res = self._adj_device(ksp, out)
        if out is not None:
            return img
        res = self._safe_squeeze(res)
        if img is not None:
            _copy_into(img, res)
            return img
        return self._out(res, kind, dev)

    def data_consistency(self, image_data, obs_data):
        """Gradient of the data-consistency term ``A^H (A x - y)`` (base.py:1075-1139)."""
        self.check_shape(image=image_data, ksp=obs_data)
        img, kind, dev = self._in(image_data)
        obs, _, _ = self._in(obs_data)
        return self._out(self._safe_squeeze(self._dc_device(img, obs)), kind, dev)

    def _op(self, image, coeffs):
        """Raw batched type 2 without smaps / normalisation; writes into ``coeffs`` (base.py:1013)."""
        img, _, _ = self._in(image)
        T = int(np.prod(img.shape[: img.ndim - len(self.shape)])) if img.ndim > len(self.shape) else 1
        img = img.reshape(T, *self.shape)
        out = torch.empty((T, self.n_samples), dtype=self._cdt, device=self.device)
        step = self.raw_op.n_trans
        for t0 in range(0, T, step):
            if self._spread_only:
                self.raw_op.plan.interp(img[t0:t0 + step].data_ptr(), out[t0:t0 + step].data_ptr(),
                                        min(step, T - t0), self.stream)
            else:
                self.raw_op.type2(img[t0:t0 + step], None, out[t0:t0 + step], 1.0)
        _copy_into(coeffs, out.reshape(coeffs.shape))
        return coeffs

    def _adj_op(self, coeffs, image):
        """Raw batched type 1 (density applied, no smaps / normalisation); writes into ``image``
        (base.py:1068-1073).  Used in place by the Toeplitz kernel builder (toeplitz.py:143-149)."""
        ksp, _, _ = self._in(coeffs)
        ksp = ksp.reshape(-1, self.n_samples)
        T = ksp.shape[0]
        out = torch.empty((T, *self.shape), dtype=self._cdt, device=self.device)
        step = self.raw_op.n_trans
        for t0 in range(0, T, step):
            k = ksp[t0:t0 + step]
            if self._spread_only:
                if self._density_d is not None:
                    k = k * self._density_d
                self.raw_op.plan.spread(k.data_ptr(), out[t0:t0 + step].data_ptr(), k.shape[0],
                                        self.stream)
            else:
                self.raw_op.type1(k, self._density_d, None, out[t0:t0 + step], False, 1.0)
        _copy_into(image, out.reshape(image.shape))
        return image

    # ------------------------------------------------------------------ Lipschitz constant / solvers
    def get_lipschitz_cst(self, max_iter=10):
        """Power method on the single-coil ``A^H A`` (density included), device resident.

        Mirrors ``FourierOperatorBase.get_lipschitz_cst`` + ``power_method``
        (base.py:626-665, 1155-1221): random start from ``np.random.random``, stop when the norm
        changes by less than 1e-6.  Returns a numpy float32 scalar.
        """
        saved = (self._n_coils, self._n_batchs, self._smaps, self._smaps_d, self.squeeze_dims)
        self._smaps = None
        self._smaps_d = None
        self._n_coils = 1
        self._n_batchs = 1
        try:
            x = np.random.random(self.shape).astype(np.complex64)
            x = to_device(x, self.device, self._cdt)
            x_norm = torch.linalg.norm(x)
            x = x / x_norm
            x_new_norm = x_norm
            i = 0
            for i in range(max_iter):  # noqa: B007
                x_new = self._adj_device(self._op_device(x.reshape(1, 1, *self.shape)))
                x_new_norm = torch.linalg.norm(x_new)
                x_new = x_new / x_new_norm
                if torch.abs(x_norm - x_new_norm) < 1e-6:
                    break
                x_norm = x_new_norm
                x = x_new
            if i == max_iter - 1:
                warnings.warn("Lipschitz constant did not converge")
            return (np.float64 if self._double else np.float32)(x_new_norm.item())
        finally:
            self._n_coils, self._n_batchs, self._smaps, self._smaps_d, self.squeeze_dims = saved

    def pinv_solver(self, kspace_data, optim="lsqr", **kwargs):
        """Solve ``A x = y`` (base.py:667-690).  ``cg``, ``lsqr`` and ``lsmr`` run device resident
        (``mrinufft_b200.solvers``, iterate-level mirrors of ``extras/optim.py``); any other
        registered optimiser uses the reference implementation on top of ``op`` / ``adj_op``."""
        from .solvers import SOLVERS

        if isinstance(optim, str) and optim in SOLVERS:
            return SOLVERS[optim](self, kspace_data, **kwargs)
        return super().pinv_solver(kspace_data, optim=optim, **kwargs)

    # ------------------------------------------------------------------ Toeplitz Gram operator
    def compute_toeplitz_kernel(self, weights=None):
        """Spectrum of the Toeplitz embedding of ``A^H W A`` on the ``2N`` grid, device resident.

        Same construction as ``compute_toeplitz_kernel`` / ``_compute_toep_2d`` / ``_compute_toep_3d``
        (src/mrinufft/operators/toeplitz.py:35-200): ``2^(d-1)`` raw adjoints of phase-modulated
        weights give the outer lags, Hermitian symmetry the rest, and one real inverse FFT the (real)
        circulant spectrum.  The adjoints run in ``libb200nufft.so``; the assembly is a handful of
        torch slicing ops at plan time.  Returns (and caches) a float32 CUDA tensor of shape ``2N``.
        """
        if self.ndim not in (2, 3):
            raise ValueError(f"Toeplitz kernel calculation not implemented for ndim={self.ndim}")
        if self._double:
            raise ValueError("the device Toeplitz path is single precision (gram_op falls back to the "
                             "reference construction for precision='double')")
        if any(s % 2 for s in self.shape):
            raise ValueError(f"Toeplitz kernel computation only supports even grid sizes, got {self.shape}.")
        if self._spread_only:
            raise ValueError("Toeplitz kernel needs a full NUFFT plan")
        dev, M = self.device, self.n_samples
        if weights is None:
            w = self._density_d if self._density_d is not None else torch.ones(M, dtype=torch.float32, device=dev)
        else:
            w = to_device(weights, dev)
            w = (w.real if w.is_complex() else w).to(torch.float32).reshape(-1)
        omega = to_device(self._samples, dev, torch.float32)  # (M, d) radians

        def adj(signs):
            ksp = modulated_weights(w, omega, signs, self.shape)
            out = torch.empty((1, *self.shape), dtype=torch.complex64, device=dev)
            self.raw_op.type1(ksp.reshape(1, M).contiguous(), None, None, out, accumulate=False, scale=1.0)
            return out[0]

        kern = assemble_toeplitz_kernel(adj, self.shape, 1.0 / float(self.norm_factor))
        self._toeplitz_kernel = kern.to(torch.float32).contiguous()
        return self._toeplitz_kernel

    def _toeplitz_plan(self):
        """The plan whose grid is exactly ``2N`` (what ``b200_toeplitz_apply`` needs), or ``None``.  Usually the
        operator's own; a 3-D operator that took a power-of-two grid (B200_EXACT_GRID in b200nufft.h) gets a
        second, trajectory-less plan for the Gram operator on first use."""
        want = tuple(2 * s for s in self.shape)
        if self.ndim not in (2, 3) or self._double or self._spread_only:
            return None
        if tuple(self.raw_op.plan.nf) == want:
            return self.raw_op.plan
        if getattr(self, "_toep_plan", None) is None:
            raw = self.raw_op
            plan = _lib.Plan(self.shape, n_trans_max=raw.n_trans, eps=raw.eps, upsampfac=raw.upsampfac,
                             device=self.device.index, exact_grid=True)
            self._toep_plan = plan if tuple(plan.nf) == want else False
            if self._toep_plan is False:
                plan.close()
        return self._toep_plan or None

    def _gram_device(self, img: torch.Tensor) -> torch.Tensor:
        """Toeplitz ``A^H A x`` on device tensors: img (B, 1|C, *XYZ) -> same shape."""
        B, C, XYZ = self.n_batchs, self.n_coils, self.shape
        if getattr(self, "_toeplitz_kernel", None) is None or not torch.is_tensor(self._toeplitz_kernel):
            self.compute_toeplitz_kernel()
        kern = self._toeplitz_kernel
        scale = 1.0 / float(np.prod([2 * s for s in XYZ]))
        plan = self._toeplitz_plan()
        if self.uses_sense:
            img = img.reshape(B, *XYZ)
            out = torch.empty((B, 1, *XYZ), dtype=self._cdt, device=self.device)
            for b in range(B):
                for i, (c0, c1) in enumerate(self._chunks()):
                    plan.toeplitz_apply(img[b].data_ptr(), self._smaps_d[c0:c1].data_ptr(), kern.data_ptr(),
                                        out[b, 0].data_ptr(), c1 - c0, int(i > 0), scale, self.stream)
        else:
            img = img.reshape(B, C, *XYZ)
            out = torch.empty((B, C, *XYZ), dtype=self._cdt, device=self.device)
            for b in range(B):
                for c0, c1 in self._chunks():
                    plan.toeplitz_apply(img[b, c0:c1].data_ptr(), 0, kern.data_ptr(),
                                        out[b, c0:c1].data_ptr(), c1 - c0, 0, scale, self.stream)
        return out

    def gram_op(self, data, toeplitz=True):
        """Gram operator ``A^H A`` (base.py:316-342).  ``toeplitz=True`` applies the Toeplitz embedding
        with two zero-padding-aware FFTs per coil on the device (no spreading / interpolation)."""
        self.check_shape(image=data)
        if not toeplitz:
            return self.adj_op(self.op(data))
        if self._toeplitz_plan() is None:
            # oversampled grid is not exactly 2N: reference construction on top of op / adj_op
            from mrinufft.operators.toeplitz import compute_toeplitz_kernel as _ref_kernel

            if not isinstance(getattr(self, "_toeplitz_kernel", None), np.ndarray):
                self._toeplitz_kernel = _ref_kernel(self, self.density)
            return super().gram_op(data, toeplitz=True)
        img, kind, dev = self._in(data)
        return self._out(self._safe_squeeze(self._gram_device(img)), kind, dev)

    # ------------------------------------------------------------------ off-resonance correction
    def with_off_resonance_correction(self, readout_time, b0_map=None, r2star_map=None, mask=None,
                                      interpolator="svd"):
        """Operator with off-resonance correction (base.py:385-398); the L interpolators ride the coil
        batch of the device transforms (``mrinufft_b200.off_resonance``)."""
        from .off_resonance import MRIB200FourierCorrected

        return MRIB200FourierCorrected(self, b0_map, readout_time, r2star_map, mask, interpolator)

    # ------------------------------------------------------------------ autodiff
    def make_autograd(self, *, wrt_data=True, wrt_traj=False, paired_batch=False):
        """Torch autograd wrapper (role of base.py:536-574).  The reference module hard-imports
        ``deepinv`` (autodiff.py:11); ours is the same wrapper without that dependency."""
        if not self.autograd_available:
            raise ValueError("Backend does not support auto-differentiation.")
        from .autodiff import MRINufftAutoGrad

        return MRINufftAutoGrad(self, wrt_data=wrt_data, wrt_traj=wrt_traj, paired_batch=paired_batch)

    # ------------------------------------------------------------------ density compensation
    @classmethod
    def pipe(cls, kspace_loc, volume_shape, max_iter=10, osf=2, normalize=True, **kwargs):
        """Pipe's iterative density compensation, device resident (``MRIfinufft.pipe``,
        finufft.py:195-245): ``d <- d / |G G^H d|`` with a spread/interp-only plan on a grid of
        size ``volume_shape`` whose kernel shape is set by ``osf``; optional PSF normalisation
        with one full op + adj_op."""
        kwargs.pop("backend", None)
        grid_op = cls(samples=kspace_loc, shape=volume_shape, upsampfac=osf, spreadinterponly=1,
                      **kwargs)
        M = grid_op.n_samples
        d = torch.ones(M, dtype=torch.float32, device=grid_op.device)
        norm2 = float(grid_op.norm_factor) ** 2
        for _ in range(max_iter):
            grid_op.raw_op.plan.pipe_iteration(d.data_ptr(), grid_op.stream)
            d *= norm2  # the reference applies 1/norm in both op and adj_op of grid_op
        if normalize:
            test_op = cls(samples=kspace_loc, shape=volume_shape, **kwargs)
            test_im = torch.ones((1, 1, *test_op.shape), dtype=torch.complex64, device=test_op.device)
            ksp = test_op._op_device(test_im)
            ksp = ksp * d.to(test_op.device)
            recon = test_op._adj_device(ksp)
            d = d / torch.mean(torch.abs(recon))
        return torch.abs(d).cpu().numpy()

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(\n"
            f"  shape: {self.shape}\n"
            f"  n_coils: {self.n_coils}\n"
            f"  n_samples: {self.n_samples}\n"
            f"  uses_sense: {self.uses_sense}\n"
            f"  device: {self.device}, eps: {self.eps}, upsampfac: {self.upsampfac},"
            f" kernel width: {self.raw_op.plan.w}, fine grid: {self.raw_op.plan.nf}\n"
            ")"
        )


def _copy_into(dst, src: torch.Tensor):
    """Write the device result ``src`` into the caller-provided buffer ``dst`` (any array type)."""
    root = module_name(dst)
    if root == "torch":
        dst.copy_(src.reshape(dst.shape))
    elif isinstance(dst, np.ndarray):
        dst[...] = src.reshape(dst.shape).cpu().numpy()
    else:
        t = torch.as_tensor(dst, device=src.device)
        t.copy_(src.reshape(t.shape))

