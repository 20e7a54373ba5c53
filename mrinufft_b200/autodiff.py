"""Torch autograd wrapper for the b200 operator.

Restates ``mrinufft.operators.autodiff`` (``src/mrinufft/operators/autodiff.py:14-456``) without its
hard ``deepinv`` import (autodiff.py:11), which is absent from this image.  Forward / backward
formulas are the reference's:

* d op / d data      = adj_op(dy)                                   (autodiff.py:14-19)
* d op / d samples   = sum_{b,c} -i conj(dy) op(x r_d), per axis d  (autodiff.py:22-41)
* d adj / d data     = op(dx)                                       (autodiff.py:64-67)
* d adj / d samples  = sum_{b,c} i y op_{+}(conj(dx) r_d)           (autodiff.py:69-86) with the
  opposite-sign plan and conjugated smaps (``grad_traj_plan``), which ``MRIB200NUFFT`` provides by
  flipping the sign of the same plan.
* d op / d field_map  = conj(x) adj_op(dy t)                        (autodiff.py:44-55, with the
  conjugate the reference omits -- see ``_backward_op_field_map``)
* d adj / d field_map = conj(dx) adj_op(y t)                        (autodiff.py:89-100), ``t`` the
  operator's ``full_readout_time``, for off-resonance-corrected operators (``MRIFourierCorrected``
  and its batched b200 subclass) only.
"""

from __future__ import annotations

import numpy as np
import torch

from mrinufft.operators.off_resonance import MRIFourierCorrected

from ._arrays import NP2TORCH


def _grid_r(shape, like):
    r = [torch.linspace(-s / 2, s / 2 - 1, s) for s in shape]
    grid_r = torch.meshgrid(*r, indexing="ij")
    return torch.stack(grid_r, dim=0).type_as(like)


def _backward_op_data(nufft, x, dy):
    if not nufft._grad_wrt_data:
        return None
    return nufft.adj_op(dy)


def _backward_op_samples(nufft, x, dy):
    if not nufft._grad_wrt_traj:
        return None
    grid_r = _grid_r(nufft.shape, x)
    rows = [
        torch.sum(-1j * torch.conj(dy) * nufft.op(x * grid_r[i]), dim=(0, 1))
        for i in range(grid_r.size(0))
    ]
    return torch.stack(rows, dim=0).transpose(0, 1).to(NP2TORCH[np.dtype(nufft.dtype)])


def _readout_time(nufft, like):
    t = nufft.full_readout_time
    if not torch.is_tensor(t):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(t)))
    return t.to(like.device)


def _backward_op_field_map(nufft, x, dy):
    if not nufft._grad_wrt_field_map or not isinstance(nufft, MRIFourierCorrected):
        return None
    # conj(x): y_m = sum_n A_mn exp(f_n t_m) x_n is holomorphic in f, so the VJP is conj(dy_m/df_n) dy_m.
    # The reference multiplies by x itself (autodiff.py:54), which does not agree with torch's autograd
    # through the dense model -- the check its own test intends (tests/operators/test_autodiff.py:248-259,
    # whose atol = 1e-1 is far above the gradient's magnitude); see tests/test_autodiff_cpu.py.
    return x.conj() * nufft.adj_op(dy * _readout_time(nufft, dy))


def _backward_adj_field_map(nufft, y, dx):
    if not nufft._grad_wrt_field_map or not isinstance(nufft, MRIFourierCorrected):
        return None
    return dx.conj() * nufft.adj_op(y * _readout_time(nufft, y))


def _backward_adj_data(nufft, y, dx):
    if not nufft._grad_wrt_data:
        return None
    return nufft.op(dx)


def _backward_adj_samples(nufft, y, dx):
    if not nufft._grad_wrt_traj:
        return None
    with nufft.grad_traj_plan():
        grid_r = _grid_r(nufft.shape, dx)
        rows = [
            torch.sum(1j * y * nufft.op(torch.conj(dx) * grid_r[i]), dim=(0, 1))
            for i in range(grid_r.size(0))
        ]
        grad_traj = torch.stack(rows, dim=0).transpose(0, 1).to(NP2TORCH[np.dtype(nufft.dtype)])
    return grad_traj


class _NUFFT_OP(torch.autograd.Function):
    """Autograd support for ``op`` (autodiff.py:103-132)."""

    @staticmethod
    def forward(ctx, x, traj, field_map, nufft_op):
        ctx.save_for_backward(x)
        ctx.nufft = nufft_op
        return nufft_op.op(x)

    @staticmethod
    def backward(ctx, dy):
        x = ctx.saved_tensors[0]
        gt = _backward_op_samples(ctx.nufft, x, dy)
        if gt is not None and traj_device(ctx) is not None:
            gt = gt.to(traj_device(ctx))
        return (_backward_op_data(ctx.nufft, x, dy), gt, _to_leaf(_backward_op_field_map(ctx.nufft, x, dy), ctx), None)


class _NUFFT_ADJOP(torch.autograd.Function):
    """Autograd support for ``adj_op`` (autodiff.py:135-154)."""

    @staticmethod
    def forward(ctx, y, traj, field_map, nufft_op):
        ctx.save_for_backward(y)
        ctx.nufft = nufft_op
        return nufft_op.adj_op(y)

    @staticmethod
    def backward(ctx, dx):
        y = ctx.saved_tensors[0]
        gt = _backward_adj_samples(ctx.nufft, y, dx)
        if gt is not None and traj_device(ctx) is not None:
            gt = gt.to(traj_device(ctx))
        return (_backward_adj_data(ctx.nufft, y, dx), gt, _to_leaf(_backward_adj_field_map(ctx.nufft, y, dx), ctx), None)


def traj_device(ctx):
    return getattr(ctx.nufft, "_traj_grad_device", None)


def _to_leaf(grad, ctx):
    """Field-map gradient in the shape / on the device of the field-map tensor (it may live on the host
    while the data are CUDA tensors, like the samples)."""
    if grad is None:
        return None
    dev = getattr(ctx.nufft, "_field_map_grad_device", None)
    grad = grad.reshape(ctx.nufft.shape) if grad.numel() == int(np.prod(ctx.nufft.shape)) else grad.sum(dim=(0, 1))
    return grad if dev is None else grad.to(dev)


class MRINufftAutoGrad(torch.nn.Module):
    """Wraps the NUFFT operator to support torch autodiff (autodiff.py:157-456).

    Same constructor, ``op`` / ``adj_op`` (incl. ``paired_batch`` mode with per-item smaps and
    samples), ``samples`` property and ``update_samples`` as the reference class.
    """

    def __init__(self, nufft_op, wrt_data=True, wrt_traj=False, wrt_field_map=False,
                 paired_batch=False):
        if any((wrt_data, wrt_traj, wrt_field_map)) and nufft_op.squeeze_dims:
            raise ValueError("Squeezing dimensions is not supported for autodiff.")
        super().__init__()
        self.nufft_op = nufft_op
        self.nufft_op._grad_wrt_traj = wrt_traj
        self.nufft_op._grad_wrt_data = wrt_data
        self.nufft_op._grad_wrt_field_map = wrt_field_map
        if wrt_traj:
            self.nufft_op._make_plan_grad()
            self._samples_torch = torch.from_numpy(np.array(self.nufft_op.samples, copy=True))
            self._samples_torch.requires_grad = True
            self.nufft_op._traj_grad_device = self._samples_torch.device
        self._field_map_torch = None
        if wrt_field_map and isinstance(self.nufft_op, MRIFourierCorrected):
            fm = self.nufft_op.field_map
            fm = fm.detach().clone() if torch.is_tensor(fm) else torch.from_numpy(np.array(fm, copy=True))
            self._field_map_torch = fm
            self._field_map_torch.requires_grad = True
            self.nufft_op._field_map_grad_device = fm.device
        self.paired_batch = paired_batch

    def _field_map_arg(self, field_map):
        corrected = isinstance(self.nufft_op, MRIFourierCorrected)
        if field_map is not None and not corrected:
            raise ValueError("Underlying nufft operator does not support field map.")
        if corrected and field_map is None:
            field_map = self.field_map
        return field_map

    def op(self, x, smaps=None, samples=None, field_map=None):
        """Forward image -> k-space (autodiff.py:222-253)."""
        if self.paired_batch:
            return self._op_batched(x, smaps, samples)
        return _NUFFT_OP.apply(x, self.samples, self._field_map_arg(field_map), self.nufft_op)

    def adj_op(self, kspace, smaps=None, samples=None, field_map=None):
        """Adjoint k-space -> image (autodiff.py:255-289)."""
        if self.paired_batch:
            return self._adj_op_batched(kspace, smaps, samples)
        return _NUFFT_ADJOP.apply(kspace, self.samples, self._field_map_arg(field_map), self.nufft_op)

    def _op_batched(self, batched_imgs, batched_smaps=None, batched_samples=None):
        self._check_input_shape(smaps=batched_smaps, imgs=batched_imgs, samples=batched_samples)
        out = []
        for i in range(len(batched_imgs)):
            try:
                if batched_smaps is not None:
                    self.nufft_op.smaps = batched_smaps[i]
                if batched_samples is not None:
                    self.samples = batched_samples[i]
                out.append(_NUFFT_OP.apply(batched_imgs[i], self.samples, None, self.nufft_op))
            except Exception as e:
                raise RuntimeError(f"Failed at batch index {i}") from e
        return torch.stack(out, dim=0)

    def _adj_op_batched(self, batched_kspace, batched_smaps=None, batched_samples=None):
        self._check_input_shape(smaps=batched_smaps, kspace=batched_kspace, samples=batched_samples)
        out = []
        for i in range(len(batched_kspace)):
            try:
                if batched_smaps is not None:
                    self.nufft_op.smaps = batched_smaps[i]
                if batched_samples is not None:
                    self.samples = batched_samples[i]
                out.append(_NUFFT_ADJOP.apply(batched_kspace[i], self.samples, None, self.nufft_op))
            except Exception as e:
                raise RuntimeError(f"Failed at batch index {i}") from e
        return torch.stack(out, dim=0)

    @property
    def samples(self):
        try:
            return self._samples_torch
        except AttributeError:
            return self.nufft_op.samples

    @samples.setter
    def samples(self, value):
        self.update_samples(value, unsafe=False)

    def update_samples(self, new_samples, *, unsafe: bool = False):
        """Update the samples of the underlying operator (autodiff.py:362-383)."""
        self._samples_torch = new_samples
        self.nufft_op._traj_grad_device = new_samples.device
        self.nufft_op.update_samples(new_samples.detach(), unsafe=unsafe)

    @property
    def field_map(self):
        """The field map as a torch tensor (autodiff.py:385-393)."""
        if not isinstance(self.nufft_op, MRIFourierCorrected):
            raise ValueError("Underlying nufft operator does not support field map.")
        if self._field_map_torch is not None:
            return self._field_map_torch
        return self.nufft_op.field_map

    @field_map.setter
    def field_map(self, value):
        self.update_field_map(value)

    def update_field_map(self, new_field_map):
        """Update the field map and recompute the interpolators (autodiff.py:399-406)."""
        if not isinstance(self.nufft_op, MRIFourierCorrected):
            raise ValueError("Underlying nufft operator does not support field map.")
        self._field_map_torch = new_field_map
        self.nufft_op._field_map_grad_device = new_field_map.device
        self.nufft_op.update_field_map(new_field_map.detach())

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.nufft_op, name)

    def _check_input_shape(self, *, imgs=None, kspace=None, smaps=None, samples=None) -> bool:
        """Batch-size validation of the paired-batch mode (autodiff.py:408-456)."""
        if imgs is not None and smaps is not None:
            D, B, C, *XYZ = imgs.shape
            D2, C2, *XYZ2 = smaps.shape
            if D != D2 or XYZ != XYZ2 or C != 1:
                raise ValueError("Shape mismatch between smaps and image")
        if kspace is not None and smaps is not None:
            D = kspace.shape[0]
            if D != smaps.shape[0]:
                raise ValueError("Shape mismatch between smaps and k-space")
        if kspace is not None and samples is not None:
            D, B, C, NS = kspace.shape
            D2, NS2, N = samples.shape
            if D != D2 or NS != NS2:
                raise ValueError("Shape mismatch between k-space and samples loc")
        if imgs is not None and samples is not None:
            D, B, C, *XYZ = imgs.shape
            D2, NS2, N = samples.shape
            if D != D2 or N != len(XYZ):
                raise ValueError("Shape mismatch between samples loc and image")
        if samples is not None and smaps is not None:
            D, NS, N = samples.shape
            D2, C2, *XYZ2 = smaps.shape
            if D != D2 or N != len(XYZ2):
                raise ValueError("Shape mismatch between samples loc and smaps")
        if imgs is not None and kspace is not None:
            raise ValueError("Input shape should not compare batched_img and batched_kspace")
        return True
