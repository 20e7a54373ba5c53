"""Torch autograd around the b200 operator (and its off-resonance-corrected wrapper).

Same public surface as ``mrinufft.operators.autodiff.MRINufftAutoGrad``
(``src/mrinufft/operators/autodiff.py:157-456``) -- ``op`` / ``adj_op`` (with the ``paired_batch`` mode
of per-item sensitivity maps and trajectories), ``samples`` / ``update_samples``, ``field_map`` /
``update_field_map`` -- without that module's hard ``deepinv`` import (autodiff.py:11), which this image
does not have.  One ``torch.autograd.Function`` serves both directions; the vector-Jacobian products
are those of Wang & Fessler (IEEE TCI 2023), as the reference states them:

====================  ==========================================================  ===================
quantity              VJP (``g`` = incoming gradient, ``r_d`` = image coordinate)   reference
====================  ==========================================================  ===================
op,  data             ``adj_op(g)``                                                autodiff.py:14-19
op,  samples[:, d]    ``sum_{b,c} -i conj(g) op(x r_d)``                           autodiff.py:22-41
op,  field map        ``conj(x) adj_op(g t)``                                      autodiff.py:44-55 (*)
adj, data             ``op(g)``                                                    autodiff.py:64-67
adj, samples[:, d]    ``sum_{b,c} i y op_+(conj(g) r_d)``  (opposite-sign plan)   autodiff.py:69-86
adj, field map        ``conj(g) adj_op(y t)``                                      autodiff.py:89-100
====================  ==========================================================  ===================

``t`` is the operator's ``full_readout_time``.  (*) The reference multiplies by ``x`` instead of
``conj(x)``: ``y_m = sum_n A_mn exp(f_n t_m) x_n`` is holomorphic in the field map ``f``, so the VJP is
``conj(dy_m / df_n) g_m``; with ``x`` it disagrees with torch's autograd through the dense model -- the
very check the reference's test intends (tests/operators/test_autodiff.py:248-259; its atol of 0.1 is far
above the gradient's magnitude).  ``tests/test_autodiff_cpu.py`` holds the dense-model comparison.

The opposite-sign plan of the adjoint's trajectory gradient is ``MRIB200NUFFT.grad_traj_plan()``: the same
device plan with the sign flipped and the maps conjugated, no second plan.
"""

from __future__ import annotations

import numpy as np
import torch
from mrinufft.operators.off_resonance import MRIFourierCorrected

from ._arrays import NP2TORCH

FORWARD, ADJOINT = 0, 1


class _Jacobians:
    """The three vector-Jacobian products of one operator, for either direction."""

    def __init__(self, nufft):
        self.nufft = nufft
        self.corrected = isinstance(nufft, MRIFourierCorrected)

    def _coords(self, like):
        axes = [torch.linspace(-s / 2, s / 2 - 1, s) for s in self.nufft.shape]
        return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=0).type_as(like)

    def _times(self, like):
        t = self.nufft.full_readout_time
        if not torch.is_tensor(t):
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(t)))
        return t.to(like.device)

    def data(self, direction, g):
        if not self.nufft._grad_wrt_data:
            return None
        return self.nufft.adj_op(g) if direction == FORWARD else self.nufft.op(g)

    def samples(self, direction, saved, g):
        nufft = self.nufft
        if not nufft._grad_wrt_traj:
            return None
        r = self._coords(saved if direction == FORWARD else g)
        if direction == FORWARD:
            cols = [torch.sum(-1j * torch.conj(g) * nufft.op(saved * r[d]), dim=(0, 1)) for d in range(len(r))]
        else:
            with nufft.grad_traj_plan():
                cols = [torch.sum(1j * saved * nufft.op(torch.conj(g) * r[d]), dim=(0, 1)) for d in range(len(r))]
        grad = torch.stack(cols, dim=1).to(NP2TORCH[np.dtype(nufft.dtype)])  # (K, ndim); keeps the real part
        where = getattr(nufft, "_traj_grad_device", None)  # the samples tensor may live on the host
        return grad if where is None else grad.to(where)

    def field_map(self, direction, saved, g):
        nufft = self.nufft
        if not (self.corrected and nufft._grad_wrt_field_map):
            return None
        if direction == FORWARD:
            grad = saved.conj() * nufft.adj_op(g * self._times(g))
        else:
            grad = g.conj() * nufft.adj_op(saved * self._times(saved))
        # one value per voxel: fold batch / coil axes, land where the field-map tensor lives
        n_vox = int(np.prod(nufft.shape))
        grad = grad.reshape(nufft.shape) if grad.numel() == n_vox else grad.reshape(-1, *nufft.shape).sum(dim=0)
        where = getattr(nufft, "_field_map_grad_device", None)
        return grad if where is None else grad.to(where)


class _ItemState:
    """The maps / trajectory one item of a paired batch ran with.

    All items share ONE operator object whose ``smaps`` / samples are overwritten item after item, and
    ``backward`` runs after the last item's forward: without putting an item's own state back, every
    item's VJPs would be taken with the LAST item's maps.  (The reference stores the shared operator in
    ``ctx`` and relies on it, autodiff.py:300-330 -- its paired-batch gradients are those of the last item's
    maps; tests/test_autodiff_cpu.py checks ours item by item.)
    """

    def __init__(self, smaps=None, samples=None):
        self.smaps, self.samples = smaps, samples

    def __call__(self, nufft):
        return _Swapped(nufft, self)


class _Swapped:
    def __init__(self, nufft, state):
        self.nufft, self.state = nufft, state

    def __enter__(self):
        nufft, st = self.nufft, self.state
        self.old_smaps = self.old_samples = None
        if st.smaps is not None and nufft.smaps is not st.smaps:
            self.old_smaps = nufft.smaps
            nufft.smaps = st.smaps
        if st.samples is not None:
            self.old_samples = np.array(nufft.samples, copy=True)
            nufft.update_samples(st.samples)

    def __exit__(self, *exc):
        if self.old_smaps is not None:
            self.nufft.smaps = self.old_smaps
        if self.old_samples is not None:
            self.nufft.update_samples(self.old_samples)
        return False


class _Transform(torch.autograd.Function):
    """``op`` (direction 0) or ``adj_op`` (direction 1) of ``nufft`` with the VJPs above."""

    @staticmethod
    def forward(ctx, inp, samples, field_map, nufft, direction, item_state=None):
        ctx.save_for_backward(inp)
        ctx.jac = _Jacobians(nufft)
        ctx.direction = direction
        ctx.item_state = item_state
        return nufft.op(inp) if direction == FORWARD else nufft.adj_op(inp)

    @staticmethod
    def backward(ctx, g):
        (saved,) = ctx.saved_tensors
        jac, direction = ctx.jac, ctx.direction

        def vjps():
            return (jac.data(direction, g), jac.samples(direction, saved, g),
                    jac.field_map(direction, saved, g), None, None, None)

        if ctx.item_state is None:
            return vjps()
        with ctx.item_state(jac.nufft):
            return vjps()


def _pairing_errors(imgs, kspace, smaps, samples):
    """What must agree between the arguments of the paired-batch mode (autodiff.py:408-456)."""
    if imgs is not None and kspace is not None:
        yield "Input shape should not compare batched_img and batched_kspace"
    if imgs is not None and smaps is not None:
        if imgs.shape[0] != smaps.shape[0] or imgs.shape[2] != 1 or tuple(imgs.shape[3:]) != tuple(smaps.shape[2:]):
            yield "Shape mismatch between smaps and image"
    if kspace is not None and smaps is not None and kspace.shape[0] != smaps.shape[0]:
        yield "Shape mismatch between smaps and k-space"
    if kspace is not None and samples is not None:
        if kspace.shape[0] != samples.shape[0] or kspace.shape[3] != samples.shape[1]:
            yield "Shape mismatch between k-space and samples loc"
    if imgs is not None and samples is not None:
        if imgs.shape[0] != samples.shape[0] or samples.shape[2] != imgs.ndim - 3:
            yield "Shape mismatch between samples loc and image"
    if samples is not None and smaps is not None:
        if samples.shape[0] != smaps.shape[0] or samples.shape[2] != smaps.ndim - 2:
            yield "Shape mismatch between samples loc and smaps"


class MRINufftAutoGrad(torch.nn.Module):
    """Differentiable view of a NUFFT operator (drop-in for the reference class of the same name).

    Parameters
    ----------
    nufft_op: the b200 operator, or an off-resonance-corrected operator built on it.
    wrt_data, wrt_traj, wrt_field_map: which gradients ``backward`` produces.
    paired_batch: an extra leading axis pairs every data item with its own maps / trajectory; the items
        run one after the other (autodiff.py:291-360).
    """

    def __init__(self, nufft_op, wrt_data=True, wrt_traj=False, wrt_field_map=False, paired_batch=False):
        if (wrt_data or wrt_traj or wrt_field_map) and nufft_op.squeeze_dims:
            raise ValueError("Squeezing dimensions is not supported for autodiff.")
        super().__init__()
        self.nufft_op = nufft_op
        self.paired_batch = paired_batch
        nufft_op._grad_wrt_data = wrt_data
        nufft_op._grad_wrt_traj = wrt_traj
        nufft_op._grad_wrt_field_map = wrt_field_map
        if wrt_traj:
            nufft_op._make_plan_grad()
            self._samples_torch = torch.from_numpy(np.array(nufft_op.samples, copy=True)).requires_grad_(True)
            nufft_op._traj_grad_device = self._samples_torch.device
        self._field_map_torch = None
        if wrt_field_map and isinstance(nufft_op, MRIFourierCorrected):
            fm = nufft_op.field_map
            fm = fm.detach().clone() if torch.is_tensor(fm) else torch.from_numpy(np.array(fm, copy=True))
            self._field_map_torch = fm.requires_grad_(True)
            nufft_op._field_map_grad_device = fm.device

    def _own(self, name):
        """A tensor attribute of this module, whether it is a plain tensor or was made an nn.Parameter."""
        if name in self.__dict__:
            return self.__dict__[name]
        return self.__dict__.get("_parameters", {}).get(name)

    # -- transforms ---------------------------------------------------------------------------
    def _run(self, direction, inp, smaps, samples, field_map):
        if self.paired_batch:
            return self._run_paired(direction, inp, smaps, samples)
        corrected = isinstance(self.nufft_op, MRIFourierCorrected)
        if field_map is not None and not corrected:
            raise ValueError("Underlying nufft operator does not support field map.")
        if corrected and field_map is None:
            field_map = self.field_map
        return _Transform.apply(inp, self.samples, field_map, self.nufft_op, direction)

    def _run_paired(self, direction, batch, smaps, samples):
        which = {"imgs": batch} if direction == FORWARD else {"kspace": batch}
        self._check_input_shape(smaps=smaps, samples=samples, **which)
        out = []
        for i, item in enumerate(batch):
            try:
                if smaps is not None:
                    self.nufft_op.smaps = smaps[i]
                if samples is not None:
                    self.samples = samples[i]
                state = _ItemState(None if smaps is None else smaps[i],
                                   None if samples is None else samples[i].detach())
                out.append(_Transform.apply(item, self.samples, None, self.nufft_op, direction, state))
            except Exception as exc:
                raise RuntimeError(f"Failed at batch index {i}") from exc
        return torch.stack(out, dim=0)

    def op(self, x, smaps=None, samples=None, field_map=None):
        """Image -> k-space, ``(B, C, K)`` (autodiff.py:222-253)."""
        return self._run(FORWARD, x, smaps, samples, field_map)

    def adj_op(self, kspace, smaps=None, samples=None, field_map=None):
        """k-space -> image, ``(B, 1 | C, *shape)`` (autodiff.py:255-289)."""
        return self._run(ADJOINT, kspace, smaps, samples, field_map)

    def _check_input_shape(self, *, imgs=None, kspace=None, smaps=None, samples=None) -> bool:
        for message in _pairing_errors(imgs, kspace, smaps, samples):
            raise ValueError(message)
        return True

    # -- differentiable parameters --------------------------------------------------------------
    @property
    def samples(self):
        own = self._own("_samples_torch")
        return own if own is not None else self.nufft_op.samples

    @samples.setter
    def samples(self, value):
        self.update_samples(value, unsafe=False)

    def update_samples(self, new_samples, *, unsafe: bool = False):
        """New sample locations for the wrapped operator (autodiff.py:362-383)."""
        self._samples_torch = new_samples
        self.nufft_op._traj_grad_device = new_samples.device
        self.nufft_op.update_samples(new_samples.detach(), unsafe=unsafe)

    def _require_corrected(self):
        if not isinstance(self.nufft_op, MRIFourierCorrected):
            raise ValueError("Underlying nufft operator does not support field map.")

    @property
    def field_map(self):
        """The field map as a torch tensor (autodiff.py:385-393)."""
        self._require_corrected()
        own = self._own("_field_map_torch")
        return own if own is not None else self.nufft_op.field_map

    @field_map.setter
    def field_map(self, value):
        self.update_field_map(value)

    def update_field_map(self, new_field_map):
        """New field map; the interpolators are recomputed (autodiff.py:399-406)."""
        self._require_corrected()
        self._field_map_torch = new_field_map
        self.nufft_op._field_map_grad_device = new_field_map.device
        self.nufft_op.update_field_map(new_field_map.detach())

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.nufft_op, name)
