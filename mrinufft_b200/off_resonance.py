"""Off-resonance-corrected operator with the interpolators riding the coil batch.

The reference's ``MRIFourierCorrected`` (``src/mrinufft/operators/off_resonance.py:232-332``) models

    op(x)      = sum_l  B[:, l] . F( C[l] . x )                (L NUFFTs per coil, Python loop)
    adj_op(y)  = sum_l  conj(C[l]) . F^H( conj(B[:, l]) . y )

on top of any backend.  On the b200 backend the spatial interpolators are folded into the
sensitivity maps -- ``V[l, c] = C[l] . S[c]`` -- so that the L x C products become "virtual coils" of
one batched SENSE transform: the multiply by ``V`` (type 2) and the conj-multiply + sum over all
(l, c) (type 1) are the fused pad / crop passes of ``libb200nufft.so``, and one library call
transforms up to 32 virtual coils.  What remains outside the library is the temporal weighting
``sum_l B[k, l] . k_v[l, c, k]`` -- one elementwise pass over k-space.

Falls back to the reference loop when the virtual maps would not fit in device memory, for
calibrationless multi-coil operators and for non-b200 operators.
"""

from __future__ import annotations

import numpy as np
import torch

from mrinufft.operators.off_resonance import MRIFourierCorrected

from . import _lib
from ._arrays import to_device


class MRIB200FourierCorrected(MRIFourierCorrected):
    """``MRIFourierCorrected`` whose ``op`` / ``adj_op`` run as one batched b200 SENSE transform."""

    def __init__(self, fourier_op, b0_map=None, readout_time=None, r2star_map=None, mask=None,
                 interpolator="svd"):
        self._fused = None
        self._fused_failed = False
        super().__init__(fourier_op, b0_map, readout_time, r2star_map, mask, interpolator)

    def compute_interpolator(self, *args, **kwargs):
        ret = super().compute_interpolator(*args, **kwargs)
        self._fused = None
        self._fused_failed = False
        return ret

    # ------------------------------------------------------------------ device tables
    def _ensure_fused(self) -> bool:
        if self._fused is not None:
            return True
        if self._fused_failed:
            return False
        from .operator import MRIB200NUFFT

        fo = self._fourier_op
        ok = isinstance(fo, MRIB200NUFFT) and not fo._spread_only and (fo.uses_sense or fo.n_coils == 1)
        ok = ok and not fo._double  # the batched path is single precision
        # samples are kept in radians; a trajectory that small would be re-scaled on re-entry
        ok = ok and float(np.abs(fo.samples).max()) - 1e-4 >= 0.5
        if ok:
            L, C = int(self.n_interpolators), int(fo.n_coils)
            need = 8.0 * L * C * float(np.prod(fo.shape))
            free, _ = torch.cuda.mem_get_info(fo.device)
            ok = need < 0.25 * free
        if not ok:
            self._fused_failed = True
            return False
        dev = fo.device
        Bd = to_device(np.asarray(self.B) if not torch.is_tensor(self.B) else self.B, dev, torch.complex64)
        Cd = to_device(np.asarray(self.C) if not torch.is_tensor(self.C) else self.C, dev, torch.complex64)
        Cd = Cd.reshape(L, *fo.shape)
        if fo.uses_sense:
            V = (Cd[:, None] * fo._smaps_d[None]).reshape(L * C, *fo.shape).contiguous()
        else:
            V = Cd.contiguous()
        vop = MRIB200NUFFT(
            fo.samples, fo.shape, n_coils=L * C, n_batchs=fo.n_batchs, smaps=V, squeeze_dims=False,
            density=False, eps=fo.eps, upsampfac=fo.upsampfac, gpu_device_id=dev.index,
            coil_chunk=min(32, L * C),
        )
        if fo.raw_op.isign_flip != vop.raw_op.isign_flip:
            vop.raw_op.toggle_grad_traj()
        self._fused = {"vop": vop, "B": Bd.reshape(-1, L).contiguous(), "L": L, "C": C,
                       "smaps_id": getattr(fo, "_smaps_version", 0), "pts_id": fo.raw_op.pts_version}
        return True

    def _vop(self):
        """The batched operator, brought in line with the inner operator's current state, or ``None``
        when the call has to take the reference loop.

        ``update_samples`` / ``grad_traj_plan`` / the ``smaps`` and ``density`` setters reach the inner
        operator through ``MRIFourierCorrected.__getattr__`` and never this class, so the state is
        compared on every call: new sample locations are handed to the batched plan (the device copy
        of the inner plan, no host round trip), a flipped FFT sign is mirrored, replaced maps rebuild
        the virtual coils.  With conjugated maps (trajectory VJP of a SENSE operator,
        ``base.py:1234-1238``) only ``S`` is conjugated, not ``C[l]``: ``conj(V)`` would be wrong, so
        that state uses the reference loop on the inner (toggled) operator.
        """
        fo = self._fourier_op
        if fo._conj_smaps:
            return None
        f = self._fused
        if fo.uses_sense and f["smaps_id"] != getattr(fo, "_smaps_version", 0):  # smaps were replaced: rebuild
            self._fused = None
            if not self._ensure_fused():
                return None
            f = self._fused
        vop = f["vop"]
        if f["pts_id"] != fo.raw_op.pts_version:
            if fo.n_samples != int(self.n_shots) * int(self.n_samples_per_shot):
                return None  # the temporal interpolator no longer matches the trajectory
            vop.raw_op._set_pts(fo.raw_op._pts)
            vop._samples = fo._samples
            vop._toeplitz_kernel = None
            f["pts_id"] = fo.raw_op.pts_version
        if vop.raw_op.isign_flip != fo.raw_op.isign_flip:
            vop.raw_op.toggle_grad_traj()
        vop._density_d = fo._density_d  # density multiplies k-space before the adjoint only
        vop._density = fo._density
        return vop

    def update_samples(self, new_samples, *, unsafe: bool = False):
        """New sample locations for the inner operator (and, through `_vop`, the batched one).  The reference
        wrapper inherits `FourierOperatorBase.update_samples` (base.py:809-830), which only stores the array
        on the wrapper and leaves the inner operator on the old trajectory."""
        self._fourier_op.update_samples(new_samples, unsafe=unsafe)
        self._samples = self._fourier_op.samples

    @property
    def samples(self):
        return self._fourier_op.samples

    @samples.setter
    def samples(self, new_samples):
        self.update_samples(new_samples)

    # ------------------------------------------------------------------ autodiff
    def make_autograd(self, *, wrt_data=True, wrt_traj=False, wrt_field_map=False, paired_batch=False):
        """Torch autograd wrapper incl. the field-map gradient (off_resonance.py:399-446); the reference's
        own wrapper hard-imports ``deepinv`` (autodiff.py:11)."""
        from .autodiff import MRINufftAutoGrad

        return MRINufftAutoGrad(self, wrt_data=wrt_data, wrt_traj=wrt_traj, wrt_field_map=wrt_field_map,
                                paired_batch=paired_batch)

    # ------------------------------------------------------------------ operators
    def _weights(self, kv, bw, y, expand):
        """Temporal weights between the virtual-coil k-space and the coils' k-space (`b200_orc_weights`)."""
        Bn, LC, K = kv.shape
        NK, L = bw.shape
        with torch.cuda.device(kv.device):
            _lib.check(_lib.load().b200_orc_weights(kv.data_ptr(), bw.data_ptr(), y.data_ptr(), Bn, L, LC // L, K, NK,
                                                    int(expand), torch.cuda.current_stream(kv.device).cuda_stream),
                       "b200_orc_weights")

    def op(self, data, *args):
        """Forward model with off-resonance (off_resonance.py:232-281)."""
        if args or not self._ensure_fused():
            return super().op(data, *args)
        fo = self._fourier_op
        vop = self._vop()
        if vop is None:
            return super().op(data)
        f = self._fused
        L, C, Bn = f["L"], f["C"], fo.n_batchs
        NS, NK = int(self.n_shots), int(self.n_samples_per_shot)
        img, kind, dev = fo._in(data)
        kv = vop._op_device(img.reshape(Bn, 1, *fo.shape)).contiguous()  # (B, L*C, K), virtual coil index l*C + c
        y = torch.empty((Bn, C, NS * NK), dtype=kv.dtype, device=kv.device)
        self._weights(kv, f["B"], y, expand=False)  # y[b, c, s, n] = sum_l kv[b, l, c, s, n] B[n, l]
        return fo._out(self._safe_squeeze(y), kind, dev)

    def adj_op(self, coeffs, *args):
        """Adjoint with off-resonance (off_resonance.py:283-332)."""
        if args or not self._ensure_fused():
            return super().adj_op(coeffs, *args)
        fo = self._fourier_op
        vop = self._vop()
        if vop is None:
            return super().adj_op(coeffs)
        f = self._fused
        L, C, Bn = f["L"], f["C"], fo.n_batchs
        NS, NK = int(self.n_shots), int(self.n_samples_per_shot)
        ksp, kind, dev = fo._in(coeffs)
        ksp = ksp.reshape(Bn, C, NS * NK).contiguous()
        kv = torch.empty((Bn, L * C, NS * NK), dtype=ksp.dtype, device=ksp.device)
        self._weights(kv, f["B"], ksp, expand=True)  # kv[b, l, c, s, n] = conj(B[n, l]) ksp[b, c, s, n]
        img = vop._adj_device(kv)  # (B, 1, *XYZ): conj(V) multiply and the sum over (l, c) are fused
        return fo._out(self._safe_squeeze(img), kind, dev)
