"""Device-resident conjugate gradient for ``MRIB200NUFFT.pinv_solver(optim="cg")``.

Mirrors ``mrinufft.extras.optim.cg`` (``src/mrinufft/extras/optim.py:801-902``) statement by
statement, including its quirks: the Polak-Ribiere ``beta`` is built from the
UN-conjugated dot product (``xp.dot``), ``max(0, beta)`` follows numpy's lexicographic ordering of
complex numbers, the first step is ``velocity = tol*velocity + grad/L``, and the stop test is on the
un-normalised ``||grad_new|| <= tol``.  The reference runs it through ``with_numpy_cupy`` (host
round trips per iteration without cupy); here the iterate, the k-space data and all reductions stay
on the device, the gradient is the fused ``b200_data_consistency`` call.

With coil-sharded operators (``mrinufft_b200.dist``) the reductions go through ``reduce_fn``.
"""

from __future__ import annotations

import numpy as np
import torch

from ._arrays import describe, from_device, to_device


def _lex_max0(beta: complex) -> complex:
    """``max(0, beta)`` with numpy's complex ordering: keep beta iff Re>0 or (Re==0 and Im>0)."""
    if beta.real > 0 or (beta.real == 0 and beta.imag > 0):
        return beta
    return 0.0


def cg(
    operator,
    kspace_data,
    damp: float = 0.0,
    x0=None,
    x_init=None,
    max_iter: int = 10,
    tol: float = 1e-4,
    progressbar=False,
    callback=None,
    reduce_fn=None,
):
    """Fixed-step Polak-Ribiere conjugate gradient on ``data_consistency`` (optim.py:801-902)."""
    dev = operator.device
    kind, kdev = describe(kspace_data)
    cdt = getattr(operator, "_cdt", torch.complex64)
    y = to_device(kspace_data, dev, cdt)
    lipschitz_cst = float(operator.get_lipschitz_cst())

    def _scaled_dcp():  # optim.py:151-165
        xi = operator._adj_device(y)
        yy = operator._op_device(xi)
        return xi * torch.linalg.norm(y) / torch.linalg.norm(yy)

    def _sum(v):
        return reduce_fn(v) if reduce_fn is not None else v

    old_density = None
    xi = None if x_init is None else to_device(x_init, dev, cdt)
    if operator.uses_density:
        if xi is None:
            xi = _scaled_dcp()
        old_density = operator.density
        old_density_d = operator._density_d
        operator.density = None
    try:
        full = operator.img_full_shape
        image = torch.zeros(full, dtype=cdt, device=dev) if xi is None else xi.reshape(full).clone()
        x0_d = None if x0 is None else to_device(x0, dev, cdt).reshape(full)
        velocity = torch.zeros_like(image)

        def _grad(img):
            g = operator._dc_device(img, y).reshape(full)
            if damp:
                g = g + damp * (img - x0_d) if x0_d is not None else g + damp * img
            return g

        grad = _grad(image)
        velocity = tol * velocity + grad / lipschitz_cst
        image = image - velocity
        callbacks_results = []
        for _ in range(max_iter):
            grad_new = _grad(image)
            gnorm = torch.sqrt(_sum(torch.sum(grad_new.real**2 + grad_new.imag**2)))
            if float(gnorm) <= tol:
                break
            gn, g = grad_new.flatten(), grad.flatten()
            num = _sum(torch.sum(gn * (gn - g)))  # un-conjugated dot, as xp.dot
            den = _sum(torch.sum(g * g))
            beta = _lex_max0(complex((num / den).item()))
            velocity = grad_new + beta * velocity
            image = image - velocity / lipschitz_cst
            grad = grad_new
            if callback:
                img_cb = from_device(image, kind, kdev)
                callbacks_results.append(callback(img_cb, operator, kspace_data, damp=damp, x0=x0))
        if operator.squeeze_dims:
            image = operator._safe_squeeze(image)
    finally:
        if old_density is not None:
            operator._density = old_density
            operator._density_d = old_density_d
    out = from_device(image, kind, kdev)
    if callbacks_results:
        return out, callbacks_results
    return out


_ = np
