"""``get_operator("stacked-b200")``: stack-of-trajectories (2.5-D) operator, device resident.

Mirror of ``MRIStackedNUFFT`` / ``MRIStackedNUFFTGPU`` (``src/mrinufft/operators/stacked.py:36-357,
360-837``): a 3-D acquisition that repeats one 2-D trajectory on a set of Cartesian kz planes is
``FFT along z`` followed by one 2-D NUFFT per plane.  The planes of all coils are "virtual coils" of one
2-D b200 operator (``n_coils = C * len(z_index)``), so a single library call transforms up to 32
planes; the z transform (``torch.fft``: cuFFT, a plain library FFT over contiguous rows), the
sensitivity-map multiply and the plane selection stay on the device.  The reference's generic class does the
same arithmetic through host numpy arrays (``get_operator("stacked-<x>")``, ``base.py:151-158``).
"""

from __future__ import annotations

import numpy as np
import torch

from mrinufft.operators.stacked import MRIStackedNUFFT

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

    # ------------------------------------------------------------------ helpers
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

    @staticmethod
    def _fftz_d(x):  # stacked.py:177-185 (`_fftz`): centred, orthonormal, and the reference's 1/sqrt(2)
        y = torch.fft.fftshift(torch.fft.fft(torch.fft.ifftshift(x, dim=-1), dim=-1, norm="ortho"), dim=-1)
        return y * float(1.0 / np.sqrt(2.0))

    @staticmethod
    def _ifftz_d(x):  # stacked.py:187-195
        y = torch.fft.fftshift(torch.fft.ifft(torch.fft.ifftshift(x, dim=-1), dim=-1, norm="ortho"), dim=-1)
        return y * float(1.0 / np.sqrt(2.0))

    def _zsel(self):
        return torch.as_tensor(np.asarray(self.z_index), device=self._dev, dtype=torch.long)

    # ------------------------------------------------------------------ device transforms
    def _op_device(self, img: torch.Tensor) -> torch.Tensor:
        """(B, 1|C, X, Y, Z) -> (B, C, NZ * NS)   (stacked.py:197-240)."""
        B, C, XYZ = self.n_batchs, self.n_coils, self.shape
        NS, NZ = len(self._samples2d), len(self.z_index)
        zsel = self._zsel()
        ksp = torch.empty((B, C * NZ, NS), dtype=self.operator._cdt, device=self._dev)
        sm = self._smaps_d()
        img = img.reshape(B, 1 if sm is not None else C, *XYZ)
        for b in range(B):
            coil = img[b] * sm if sm is not None else img[b]            # (C, X, Y, Z)
            kz = self._fftz_d(coil).index_select(-1, zsel)               # (C, X, Y, NZ)
            planes = kz.permute(0, 3, 1, 2).reshape(1, C * NZ, *XYZ[:2]).contiguous()
            ksp[b] = self.operator._op_device(planes)[0]
        return ksp.reshape(B, C, NZ * NS)

    def _adj_device(self, ksp: torch.Tensor) -> torch.Tensor:
        """(B, C, NZ * NS) -> (B, 1|C, X, Y, Z)   (stacked.py:254-305)."""
        B, C, XYZ = self.n_batchs, self.n_coils, self.shape
        NS, NZ = len(self._samples2d), len(self.z_index)
        zsel = self._zsel()
        sm = self._smaps_d()
        ksp = ksp.reshape(B, C * NZ, NS)
        out = torch.empty((B, 1 if sm is not None else C, *XYZ), dtype=self.operator._cdt, device=self._dev)
        for b in range(B):
            planes = self.operator._adj_device(ksp[b:b + 1].contiguous())  # (1, C*NZ, X, Y)
            imgz = torch.zeros((C, *XYZ), dtype=self.operator._cdt, device=self._dev)
            imgz.index_copy_(-1, zsel, planes.reshape(C, NZ, *XYZ[:2]).permute(0, 2, 3, 1))
            imgc = self._ifftz_d(imgz)
            out[b] = torch.sum(imgc * torch.conj(sm), dim=0, keepdim=True) if sm is not None else imgc
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
