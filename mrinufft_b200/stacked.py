"""``get_operator("stacked-b200")``: stack-of-trajectories (2.5-D) operator, device resident.

Mirror of ``MRIStackedNUFFT`` / ``MRIStackedNUFFTGPU`` (``src/mrinufft/operators/stacked.py:36-357,
360-837``): a 3-D acquisition that repeats one 2-D trajectory on a set of Cartesian kz planes is
``FFT along z`` followed by one 2-D NUFFT per plane.  The planes of all coils are "virtual coils" of one
2-D b200 operator (``n_coils = C * len(z_index)``), so a single library call transforms up to 32
planes.  The z transform is the library's own kernel (``csrc/stack_fftz.cu``, any length Z): sensitivity-map
multiply, both centring shifts, the FFT, the kz-plane selection and the (coil, stack)-major plane layout are
one pass per direction, the SENSE coil sum of the adjoint included.  The reference's generic class does the
same arithmetic through host numpy arrays (``get_operator("stacked-<x>")``, ``base.py:151-158``).
"""

from __future__ import annotations

import numpy as np
import torch

from mrinufft.operators.stacked import MRIStackedNUFFT

from . import _lib
from ._arrays import to_device
from .operator import MRIB200NUFFT, _copy_into


class MRIB200StackedNUFFT(MRIStackedNUFFT):
    """Stacked NUFFT on one B200: ``samples`` is a 2-D trajectory (+ ``z_index``) or a stacked 3-D one."""

    backend = "stacked-b200"
    available = MRIB200NUFFT.available
    autograd_available = False

    def __init__(self, samples, shape, smaps=None, z_index="auto", n_coils=1, n_batchs=1,
                 squeeze_dims=False, **kwargs):
        kwargs.pop("backend", None)
        super().__init__(samples, shape, "b200", smaps, z_index=z_index, n_coils=n_coils,
                         n_batchs=n_batchs, squeeze_dims=squeeze_dims, **kwargs)
        self._smaps_dev = None
        self._zsel_dev = None

    # ------------------------------------------------------------------ helpers
    @property
    def smaps(self):
        return self._smaps

    @smaps.setter
    def smaps(self, new_smaps):
        # numpy, torch (cpu / cuda) and cupy arrays; the base setter only takes numpy (base.py:759-760)
        if new_smaps is not None:
            if tuple(new_smaps.shape[1:]) != tuple(self.shape):
                raise ValueError("Smaps should match image shape.")
            if new_smaps.shape[0] != self._n_coils:
                self._n_coils = int(new_smaps.shape[0])
                self.log.warning("updating number of coils via Smaps.")
        self._smaps = new_smaps
        self._smaps_dev = None

    @property
    def _dev(self):
        return self.operator.device

    def _smaps_d(self):
        if self.smaps is None:
            return None
        if self._smaps_dev is None or self._smaps_dev[0] is not self.smaps:
            self._smaps_dev = (self.smaps, to_device(self.smaps, self._dev, self.operator._cdt)
                               .reshape(self.n_coils, *self.shape))
        return self._smaps_dev[1]

    def _zsel(self):
        zi = np.asarray(self.z_index)
        if self._zsel_dev is None or self._zsel_dev[0] is not self.z_index:
            if zi.ndim != 1 or zi.min() < 0 or zi.max() >= self.shape[-1] or len(np.unique(zi)) != len(zi):
                raise ValueError("z_index must hold distinct plane indices in [0, shape[-1])")
            self._zsel_dev = (self.z_index, torch.as_tensor(zi.astype(np.int32), device=self._dev))
        return self._zsel_dev[1]

    def _fftz_call(self, adjoint, src, dst):
        """One launch of the fused z transform (`b200_stack_fftz_forward` / `_adjoint`): `_fftz` / `_ifftz` of
        stacked.py:178-195 (centred, orthonormal, and the reference's 1/sqrt(2)) with everything around them."""
        X, Y, Z = self.shape
        sm = self._smaps_d()
        with torch.cuda.device(self._dev):
            _lib.stack_fftz(adjoint, src.data_ptr(), sm.data_ptr() if sm is not None else None, dst.data_ptr(),
                            self._zsel().data_ptr(), self.n_coils, X, Y, Z, len(self.z_index),
                            1.0 / np.sqrt(2.0 * Z), torch.cuda.current_stream(self._dev).cuda_stream)

    # ------------------------------------------------------------------ device transforms
    def _op_device(self, img: torch.Tensor) -> torch.Tensor:
        """(B, 1|C, X, Y, Z) -> (B, C, NZ * NS)   (stacked.py:197-240)."""
        B, C, XYZ = self.n_batchs, self.n_coils, self.shape
        NS, NZ = len(self._samples2d), len(self.z_index)
        ksp = torch.empty((B, C * NZ, NS), dtype=self.operator._cdt, device=self._dev)
        img = img.reshape(B, 1 if self.smaps is not None else C, *XYZ).contiguous()
        planes = torch.empty((1, C * NZ, *XYZ[:2]), dtype=self.operator._cdt, device=self._dev)
        for b in range(B):
            self._fftz_call(False, img[b], planes)
            ksp[b] = self.operator._op_device(planes)[0]
        return ksp.reshape(B, C, NZ * NS)

    def _adj_device(self, ksp: torch.Tensor) -> torch.Tensor:
        """(B, C, NZ * NS) -> (B, 1|C, X, Y, Z)   (stacked.py:254-305)."""
        B, C, XYZ = self.n_batchs, self.n_coils, self.shape
        NS, NZ = len(self._samples2d), len(self.z_index)
        ksp = ksp.reshape(B, C * NZ, NS)
        out = torch.empty((B, 1 if self.smaps is not None else C, *XYZ), dtype=self.operator._cdt,
                          device=self._dev)
        for b in range(B):
            planes = self.operator._adj_device(ksp[b:b + 1].contiguous()).contiguous()  # (1, C*NZ, X, Y)
            self._fftz_call(True, planes, out[b])
        return out

    # ------------------------------------------------------------------ public API
    def op(self, data, ksp=None):
        """Forward operator (stacked.py:189-195)."""
        self.check_shape(image=data, ksp=ksp)
        img, kind, dev = self.operator._in(data)
        res = self._safe_squeeze(self._op_device(img))
        if ksp is not None:
            _copy_into(ksp, res)
            return ksp
        return self.operator._out(res, kind, dev)

    def adj_op(self, coeffs, img=None):
        """Adjoint operator (stacked.py:246-252)."""
        self.check_shape(image=img, ksp=coeffs)
        ksp, kind, dev = self.operator._in(coeffs)
        res = self._safe_squeeze(self._adj_device(ksp))
        if img is not None:
            _copy_into(img, res)
            return img
        return self.operator._out(res, kind, dev)

    def data_consistency(self, image_data, obs_data):
        """``A^H (A x - y)`` (base.py:377-383), device resident."""
        self.check_shape(image=image_data, ksp=obs_data)
        img, kind, dev = self.operator._in(image_data)
        obs, _, _ = self.operator._in(obs_data)
        res = self._adj_device(self._op_device(img) - obs.reshape(self.n_batchs, self.n_coils, -1))
        return self.operator._out(self._safe_squeeze(res), kind, dev)
