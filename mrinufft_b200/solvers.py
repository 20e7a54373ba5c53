"""Device-resident iterative solvers behind ``MRIB200NUFFT.pinv_solver`` (``cg``, ``lsqr``, ``lsmr``).

The reference's solvers (``src/mrinufft/extras/optim.py``: ``lsqr`` 249-495, ``lsmr`` 498-798, ``cg``
801-902) run through ``with_numpy_cupy``: without cupy every ``op`` / ``adj_op`` of every iteration is
a host round trip of the whole k-space batch.  Here the iterates, the k-space data and all vector
updates stay on the device; only the per-batch scalars (a handful of floats per iteration) come to
the host, where the Givens-rotation recurrences run in numpy exactly as in the reference.  The vector
updates of an iteration are the library's fused kernels (``csrc/vecops.cu``: one pass over memory per group
of updates that read the same vectors, the norms / dot products accumulated in double on the way) -- e.g. the
whole image-domain update of an ``lsqr`` iteration, ``||w||``, ``x += t1 w``, ``w = v + t2 w``, is one launch.

Iterate-level parity is the contract (tests/test_solvers_cpu.py compares iterate by iterate against
the reference on its exact-NDFT backend), so the reference's quirks are kept:

* ``cg``: the Polak-Ribiere ``beta`` is built from the UN-conjugated dot product (``xp.dot``),
  ``max(0, beta)`` follows numpy's lexicographic ordering of complex numbers, the first step is
  ``velocity = tol*velocity + grad/L``, the stop test is on the un-normalised ``||grad_new|| <= tol``,
  and the step size comes from the density-weighted operator while the iteration runs un-weighted.
* ``lsqr`` / ``lsmr``: the plane rotations are decided for the whole batch at once (``any`` over the
  batch), ``lsmr`` never advances its iteration counter (so ``minrbar`` keeps its initial 1e100 and
  the condition estimate is ``mean(max(maxrbar, rhotemp) / rhotemp)``), and with a density the start
  is the scaled density-compensated adjoint while the iteration itself is un-weighted.
  NOT mirrored: the reference forgets to restore ``operator.density`` when a callback is supplied
  (early ``return`` before the restore, optim.py:491-495, 794-798).

With coil-sharded operators (``mrinufft_b200.dist``) squared norms / dot products of k-space vectors
go through ``reduce_ksp`` and those of coil-sharded (calibrationless) images through ``reduce_img``:
one small all-reduce each.
"""

from __future__ import annotations

import contextlib

import numpy as np
import torch

from . import _lib
from ._arrays import describe, from_device, to_device


# ---------------------------------------------------------------------------------------- helpers
def _lex_max0(beta: complex) -> complex:
    """``max(0, beta)`` with numpy's complex ordering: keep beta iff Re>0 or (Re==0 and Im>0)."""
    if beta.real > 0 or (beta.real == 0 and beta.imag > 0):
        return beta
    return 0.0


def _ident(v):
    return v


class _Ctx:
    """What the solvers need from an operator: device, dtypes, shapes, the device-level transforms."""

    def __init__(self, operator, kspace_data, reduce_ksp=None, reduce_img=None):
        self.op = operator
        self.dev = operator.device
        self.cdt = getattr(operator, "_cdt", torch.complex64)
        self.rnp = np.float64 if self.cdt == torch.complex128 else np.float32
        self.kind, self.kdev = describe(kspace_data)
        self.full_img = tuple(operator.img_full_shape)
        self.full_ksp = tuple(operator.ksp_full_shape)
        self.reduce_ksp = reduce_ksp or _ident
        self.reduce_img = reduce_img or _ident
        self.y = to_device(kspace_data, self.dev, self.cdt).reshape(self.full_ksp)

    def A(self, x):
        return self.op._op_device(x.reshape(self.full_img)).reshape(self.full_ksp)

    def AH(self, y):
        return self.op._adj_device(y.reshape(self.full_ksp)).reshape(self.full_img)

    def image(self, arr):
        return to_device(arr, self.dev, self.cdt).reshape(self.full_img)

    def out(self, x):
        return from_device(x, self.kind, self.kdev)

    # ---- vector updates: the library's fused kernels on the device (csrc/vecops.cu).  CPU tensors only occur
    # with the stand-in operators of the host-logic tests (tests/test_solvers_cpu.py, test_dist_cpu.py), which
    # have no CUDA device: there the same updates are written with torch.
    def _scal(self, v, B):
        """(B,) complex scalars as interleaved doubles for the C ABI."""
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.complex128), (B,)))
        return a, a.ctypes.data

    def _launch(self, fn, vecs, scalars, nred, red_arg=True):
        t0 = vecs[0]
        B, n = t0.shape[0], t0[0].numel()
        for t in vecs:
            assert t is None or (t.is_contiguous() and t.dtype == self.cdt and t.numel() == B * n)
        red = torch.empty((B, nred), dtype=torch.float64, device=t0.device) if nred else None
        keep = [self._scal(v, B) for v in scalars]
        args = [t.data_ptr() if t is not None else None for t in vecs] + [k[1] for k in keep] + [B, n]
        if red_arg:
            args.append(red.data_ptr() if nred else None)
        args += [int(self.cdt == torch.complex128), torch.cuda.current_stream(t0.device).cuda_stream]
        with torch.cuda.device(t0.device):
            _lib.check(getattr(_lib.load(), fn)(*args), fn)
        return red

    def _host_norm(self, sq, reduce):
        return np.sqrt(reduce(sq).cpu().numpy().astype(self.rnp))

    # per-batch 2-norms (optim.py:57-58) as a host (B,) array of the operator's real precision
    def _norm(self, t, reduce):
        if t.is_cuda:
            t = t.reshape(t.shape[0], -1)
            return self._host_norm(self._launch("b200_vec_cg_dots", [t, t], [], 5)[:, 0], reduce)
        sq = torch.sum(t.real.reshape(t.shape[0], -1) ** 2 + t.imag.reshape(t.shape[0], -1) ** 2, dim=1)
        return self._host_norm(sq, reduce)

    def knorm(self, t):
        return self._norm(t, self.reduce_ksp)

    def inorm(self, t):
        return self._norm(t, self.reduce_img)

    def axpby(self, out, a, x, b=None, y=None, norm=None):
        """``out = a x + b y`` with per-batch real scalars (``out`` may be ``x`` or ``y``); with ``norm`` =
        "ksp" / "img" the per-batch 2-norms of the result come back as a host array."""
        reduce = {None: None, "ksp": self.reduce_ksp, "img": self.reduce_img}[norm]
        if out.is_cuda:
            sq = self._launch("b200_vec_axpby", [out, x, y], [a, 0.0 if b is None else b], 1 if norm else 0)
            return self._host_norm(sq[:, 0], reduce) if norm else None
        r = self.bc(a, x) * x
        if y is not None:
            r = r + self.bc(b, y) * y
        out.copy_(r)
        return self._norm(out, reduce) if norm else None

    def cg_dots(self, gn, g):
        """``(||gn||^2, sum gn (gn - g), sum g g)`` over the whole batch, un-conjugated products as ``xp.dot``
        (optim.py:866-880), summed over ranks where the image is sharded."""
        if gn.is_cuda:
            d = self.reduce_img(self._launch("b200_vec_cg_dots", [gn, g], [], 5).sum(dim=0)).cpu().numpy()
            return float(d[0]), complex(d[1], d[2]), complex(d[3], d[4])
        a, b = gn.flatten(), g.flatten()
        _sum = self.reduce_img
        return (float(_sum(torch.sum(a.real ** 2 + a.imag ** 2))), complex(_sum(torch.sum(a * (a - b))).item()),
                complex(_sum(torch.sum(b * b)).item()))

    def cg_step(self, x, v, g, beta, lipschitz):
        """``v = g + beta v ; x = x - v / L`` in place (optim.py:881-883)."""
        if x.is_cuda:
            self._launch("b200_vec_cg_step", [x, v, g], [beta, -1.0 / lipschitz], 0, red_arg=False)
            return
        v.copy_(g + beta * v)
        x.sub_(v / lipschitz)

    def lsqr_step(self, x, w, v, t1, t2):
        """``||w||`` (of the incoming ``w``), then ``x += t1 w ; w = v + t2 w`` in place (optim.py:441-446)."""
        if x.is_cuda:
            return self._host_norm(self._launch("b200_vec_lsqr_step", [x, w, v], [t1, t2], 1)[:, 0], self.reduce_img)
        nw = self.inorm(w)
        x += self.bc(t1, w) * w
        w *= self.bc(t2, w)
        w += v
        return nw

    def lsmr_step(self, x, hbar, h, v, a, b, c):
        """``hbar = h + a hbar ; x += b hbar ; h = v + c h`` in place, returns ``||x||`` (optim.py:716-724, 760)."""
        if x.is_cuda:
            return self._host_norm(self._launch("b200_vec_lsmr_step", [x, hbar, h, v], [a, b, c], 1)[:, 0],
                                   self.reduce_img)
        hbar *= self.bc(a, hbar)
        hbar += h
        x += self.bc(b, hbar) * hbar
        h *= self.bc(c, h)
        h += v
        return self.inorm(x)

    # host (B,) scalars -> device tensor that broadcasts from the left (``_bc_left``, optim.py:68-86)
    def bc(self, s, like):
        s = np.array(np.broadcast_to(np.asarray(s, dtype=self.rnp), (like.shape[0],)))
        t = torch.from_numpy(s).to(like.device)
        return t.reshape(-1, *([1] * (like.ndim - 1)))

    def scaled_dcp(self):
        """Scaled density-compensated adjoint (``_scaled_dcp``, optim.py:151-165)."""
        xi = self.AH(self.y)
        yy = self.A(xi)
        num = torch.sqrt(self.reduce_ksp(torch.sum(self.y.real ** 2 + self.y.imag ** 2)))
        den = torch.sqrt(self.reduce_ksp(torch.sum(yy.real ** 2 + yy.imag ** 2)))
        return xi * (num / den)


@contextlib.contextmanager
def _density_off(operator):
    """The solvers iterate on the un-weighted operator (optim.py:838-842) -- always restored."""
    saved = operator.density if operator.uses_density else None
    if saved is not None:
        operator.density = None
    try:
        yield
    finally:
        if saved is not None:
            operator.density = saved


def _givens(a, b):
    """Plane rotation ``(c, s, r)`` with ``c a + s b = r`` in the overflow-safe form of Choi's
    SymOrtho, evaluated like ``_sym_ortho`` (optim.py:213-246): which of the four formulas is used
    is decided for the whole batch (``any`` over its entries)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if np.any(b == 0):
        return np.sign(a), 0, np.abs(a)
    if np.any(a == 0):
        return 0, np.sign(b), np.abs(b)
    if np.any(np.abs(b) > np.abs(a)):
        t = a / b
        s = np.sign(b) / np.sqrt(1 + t * t)
        return s * t, s, b / s
    t = b / a
    c = np.sign(a) / np.sqrt(1 + t * t)
    return c, c * t, a / c


class _Bidiag:
    """Golub-Kahan bidiagonalisation shared by ``lsqr`` and ``lsmr``: ``u`` (k-space) and ``v`` (image)
    live on the device, ``alpha`` / ``beta`` per batch on the host.

    ``start`` is the common prologue (optim.py:354-385 and 600-625), ``step`` the first lines of both
    loops (optim.py:402-414 and 669-680).
    """

    def __init__(self, ctx: _Ctx, x0_d):
        self.c = ctx
        self.u = ctx.y.clone(memory_format=torch.contiguous_format)
        self.bnorm = ctx.knorm(self.u)
        self.beta = self.bnorm.copy()
        if x0_d is not None:
            self.beta = ctx.axpby(self.u, 1.0, self.u, -1.0, ctx.A(x0_d).contiguous(), norm="ksp")
        self.v = None
        self.alpha = None

    def start(self, x):
        c = self.c
        if np.all(self.beta) > 0:
            c.axpby(self.u, 1 / self.beta, self.u)
            self.v = c.AH(self.u).contiguous()
            self.alpha = c.inorm(self.v)
        else:
            self.v = x.clone()
            self.alpha = np.zeros(self.v.shape[0], dtype=c.rnp)
        if np.any((self.alpha * self.beta) == 0):
            return False
        if np.all(self.alpha) > 0:
            c.axpby(self.v, 1 / self.alpha, self.v)
        return True

    def step(self):
        """Next pair of Lanczos vectors.  Returns True when the ``beta > 0`` branch ran."""
        c = self.c
        self.beta = c.axpby(self.u, 1.0, c.A(self.v).contiguous(), -self.alpha, self.u, norm="ksp")
        if not (np.all(self.beta) > 0):
            return False
        c.axpby(self.u, 1 / self.beta, self.u)
        return True

    def step_v(self):
        c = self.c
        self.alpha = c.axpby(self.v, 1.0, c.AH(self.u).contiguous(), -self.beta, self.v, norm="img")
        if np.all(self.alpha) > 0:
            c.axpby(self.v, 1 / self.alpha, self.v)


def _initial_iterate(ctx: _Ctx, x0, x_init):
    """(x, x0_d): the starting image (a fresh device tensor) and the regularisation centre."""
    x0_d = None if x0 is None else ctx.image(x0)
    if x_init is not None:
        x = ctx.image(x_init).clone(memory_format=torch.contiguous_format)
    elif x0_d is not None:
        x = x0_d.clone(memory_format=torch.contiguous_format)
    else:
        x = torch.zeros(ctx.full_img, dtype=ctx.cdt, device=ctx.dev)
    return x, x0_d


def _stop_code(test1, test2, test3, t1, rtol, atol, ctol):
    """The six stopping rules shared by lsqr / lsmr (optim.py:466-479, 769-782)."""
    if np.all(1 + test3 <= 1):
        return 6
    if np.all(1 + test2 <= 1):
        return 5
    if np.all(1 + t1 <= 1):
        return 4
    if np.all(test3 <= ctol):
        return 3
    if np.all(test2 <= atol):
        return 2
    if np.all(test1 <= rtol):
        return 1
    return 0


def _finish(ctx: _Ctx, x, callback_returns):
    if ctx.op.squeeze_dims:
        x = ctx.op._safe_squeeze(x)
    out = ctx.out(x)
    if callback_returns:
        return out, callback_returns
    return out


# ------------------------------------------------------------------------------------------- LSQR
def lsqr(
    operator,
    kspace_data,
    damp: float = 0.0,
    atol: float = 1e-6,
    btol: float = 1e-6,
    conlim: float = 1e8,
    max_iter: int = 100,
    x0=None,
    x_init=None,
    callback=None,
    progressbar=False,
    reduce_ksp=None,
    reduce_img=None,
):
    """LSQR (Paige & Saunders 1982) on ``argmin ||A x - y||^2 + damp^2 ||x - x0||^2`` (optim.py:249-495)."""
    ctx = _Ctx(operator, kspace_data, reduce_ksp, reduce_img)
    np_eps = np.finfo(ctx.rnp).eps
    ctol = 1 / conlim if conlim > 0 else 0
    with contextlib.ExitStack() as stack:
        if operator.uses_density:
            if x_init is None:
                x_init = ctx.scaled_dcp()
            stack.enter_context(_density_off(operator))
        x, x0_d = _initial_iterate(ctx, x0, x_init)
        gk = _Bidiag(ctx, x0_d)
        bnorm = gk.bnorm
        if not gk.start(x):
            return ctx.out(x)
        w = gk.v.clone()

        alpha, beta = gk.alpha, gk.beta
        rhobar = alpha
        phibar = beta
        ddnorm = res2 = xnorm = xxnorm = z = anorm = 0.0
        dampsq = damp ** 2
        cs2, sn2 = -1, 0.0
        callback_returns = []
        for _ in range(max_iter):
            if gk.step():
                beta = gk.beta
                anorm = np.sqrt(anorm ** 2 + alpha ** 2 + beta ** 2 + dampsq)
                gk.step_v()
                alpha = gk.alpha
            else:
                beta = gk.beta
            if damp:
                rhobar1 = np.sqrt(rhobar ** 2 + dampsq)
                cs1 = rhobar / rhobar1
                sn1 = damp / rhobar1
                psi = sn1 * phibar
                phibar = cs1 * phibar
            else:
                rhobar1 = rhobar
                psi = 0.0
            # rotation that removes beta from the lower-bidiagonal matrix
            cs, sn, rho = _givens(rhobar1, beta)
            theta = sn * alpha
            rhobar = -cs * alpha
            phi = cs * phibar
            phibar = sn * phibar
            tau = sn * phi
            t1 = phi / rho
            t2 = -theta / rho

            # x += (phi / rho) w ;  ||w / rho||^2 feeds the condition estimate ;  w = v - (theta / rho) w
            ddnorm = ddnorm + (ctx.lsqr_step(x, w, gk.v, t1, t2) / np.abs(rho)) ** 2

            # rotation on the right: estimate of ||x||
            delta = sn2 * rho
            gambar = -cs2 * rho
            rhs = phi - delta * z
            zbar = rhs / gambar
            xnorm = np.sqrt(xxnorm + zbar ** 2)
            gamma = np.sqrt(gambar ** 2 + theta ** 2)
            cs2 = gambar / gamma
            sn2 = theta / gamma
            z = rhs / gamma
            xxnorm = xxnorm + z ** 2

            acond = anorm * np.sqrt(ddnorm)
            res2 = res2 + psi ** 2
            rnorm = np.sqrt(phibar ** 2 + res2)
            arnorm = alpha * np.abs(tau)

            test1 = rnorm / bnorm
            test2 = arnorm / (anorm * rnorm + np_eps)
            test3 = 1 / (acond + np_eps)
            t1 = test1 / (1 + anorm * xnorm / bnorm)
            rtol = btol + atol * anorm * xnorm / bnorm

            if callback:
                callback_returns.append(callback(ctx.out(x), operator, kspace_data, damp=damp, x0=x0))
            if _stop_code(test1, test2, test3, t1, rtol, atol, ctol):
                break
    return _finish(ctx, x, callback_returns)


# ------------------------------------------------------------------------------------------- LSMR
def lsmr(
    operator,
    kspace_data,
    damp: float = 0.0,
    atol: float = 1e-6,
    btol: float = 1e-6,
    conlim: float = 1e8,
    max_iter: int = 100,
    x0=None,
    x_init=None,
    callback=None,
    progressbar=False,
    reduce_ksp=None,
    reduce_img=None,
):
    """LSMR (Fong & Saunders 2011) on the same problem as :func:`lsqr` (optim.py:498-798)."""
    ctx = _Ctx(operator, kspace_data, reduce_ksp, reduce_img)
    ctol = 1 / conlim if conlim > 0 else 0
    with contextlib.ExitStack() as stack:
        if operator.uses_density:
            if x_init is None:
                x_init = ctx.scaled_dcp()
            stack.enter_context(_density_off(operator))
        x, x0_d = _initial_iterate(ctx, x0, x_init)
        # the reference's lsmr subtracts A(x) of the STARTING iterate when x0 is given (optim.py:610-612)
        gk = _Bidiag(ctx, x if x0 is not None else None)
        normb = gk.bnorm
        if not gk.start(x):
            return ctx.out(x)
        alpha, beta = gk.alpha, gk.beta
        damp_b = np.full(x.shape[0], damp, np.float32)

        zetabar = alpha * beta
        alphabar = alpha
        rho = rhobar = cbar = 1
        sbar = 0
        h = gk.v.clone()
        hbar = torch.zeros_like(h)

        # estimate of ||r||
        betadd = beta
        betad = 0
        rhodold = 1
        tautildeold = thetatilde = zeta = d = 0
        # estimates of ||A|| and cond(A)
        normA2 = alpha * alpha
        maxrbar = 0
        callback_returns = []
        for _ in range(max_iter):
            if gk.step():
                gk.step_v()
            alpha, beta = gk.alpha, gk.beta

            chat, shat, alphahat = _givens(alphabar, damp_b)
            rhoold = rho
            c, s, rho = _givens(alphahat, beta)
            thetanew = s * alpha
            alphabar = c * alpha

            # rotation Qbar_i: R_i^T -> R_i^bar
            rhobarold = rhobar
            zetaold = zeta
            thetabar = sbar * rho
            rhotemp = cbar * rho
            cbar, sbar, rhobar = _givens(cbar * rho, thetanew)
            zeta = cbar * zetabar
            zetabar = -sbar * zetabar

            # hbar = h - (thetabar rho / (rhoold rhobarold)) hbar ;  x += (zeta / (rho rhobar)) hbar ;
            # h = v - (thetanew / rho) h
            normx = ctx.lsmr_step(x, hbar, h, gk.v, -(thetabar * rho / (rhoold * rhobarold)), zeta / (rho * rhobar),
                                  -(thetanew / rho))

            # estimate of ||r||: rotations Qhat_{k,2k+1}, Q_{k,k+1}, Qtilde_{k-1}
            betaacute = chat * betadd
            betacheck = -shat * betadd
            betahat = c * betaacute
            betadd = -s * betaacute
            thetatildeold = thetatilde
            ctildeold, stildeold, rhotildeold = _givens(rhodold, thetabar)
            thetatilde = stildeold * rhobar
            rhodold = ctildeold * rhobar
            betad = -stildeold * betad + ctildeold * betahat
            tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
            taud = (zeta - thetatilde * tautildeold) / rhodold
            d = d + betacheck * betacheck
            normr = np.sqrt(d + (betad - taud) ** 2 + betadd * betadd)

            normA2 = normA2 + beta * beta
            normA = np.sqrt(normA2)
            normA2 = normA2 + alpha * alpha

            # the reference's iteration counter stays at 0, so its minrbar keeps the initial 1e100
            # (inf once cast to float32): min(minrbar, rhotemp) is rhotemp
            maxrbar = np.max(np.maximum(maxrbar, rhobarold))
            condA = np.mean(np.maximum(maxrbar, rhotemp) / rhotemp)

            normar = np.abs(zetabar)
            test1 = normr / normb
            test2 = normar / (normA * normr) if np.all((normA * normr) != 0) else np.inf
            test3 = 1 / condA
            t1 = test1 / (1 + normA * normx / normb)
            rtol = btol + atol * normA * normx / normb

            if callback:
                callback_returns.append(callback(ctx.out(x), operator, kspace_data, damp=damp_b, x0=x0))
            if _stop_code(test1, test2, test3, t1, rtol, atol, ctol):
                break
    return _finish(ctx, x, callback_returns)


# --------------------------------------------------------------------------------------------- CG
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
    lipschitz_cst=None,
    reduce_ksp=None,
):
    """Fixed-step Polak-Ribiere conjugate gradient on ``data_consistency`` (optim.py:801-902).

    ``reduce_fn`` sums image-domain scalars over the ranks of a coil-sharded (calibrationless)
    iterate; ``lipschitz_cst`` lets a sharded caller hand in the constant agreed between ranks.
    """
    ctx = _Ctx(operator, kspace_data, reduce_ksp, reduce_fn)
    y, cdt, dev = ctx.y, ctx.cdt, ctx.dev
    if lipschitz_cst is None:
        lipschitz_cst = float(operator.get_lipschitz_cst())
    xi = None if x_init is None else ctx.image(x_init)
    with contextlib.ExitStack() as stack:
        if operator.uses_density:
            if xi is None:
                xi = ctx.scaled_dcp()
            stack.enter_context(_density_off(operator))
        full = ctx.full_img
        image = torch.zeros(full, dtype=cdt, device=dev) if xi is None else xi.reshape(full).clone()
        x0_d = None if x0 is None else ctx.image(x0)
        velocity = torch.zeros_like(image)

        def _grad(img):
            g = operator._dc_device(img, y).reshape(full)
            if damp:
                g = g + damp * (img - x0_d) if x0_d is not None else g + damp * img
            return g

        grad = _grad(image).contiguous()
        velocity = (tol * velocity + grad / lipschitz_cst).contiguous()
        image = (image - velocity).contiguous()
        callbacks_results = []
        for _ in range(max_iter):
            grad_new = _grad(image).contiguous()
            gsq, num, den = ctx.cg_dots(grad_new, grad)
            if np.sqrt(gsq) <= tol:
                break
            with np.errstate(all="ignore"):
                beta = _lex_max0(complex(np.complex128(num) / np.complex128(den)))
            ctx.cg_step(image, velocity, grad_new, beta, lipschitz_cst)
            grad = grad_new
            if callback:
                callbacks_results.append(callback(ctx.out(image.clone()), operator, kspace_data, damp=damp, x0=x0))
    return _finish(ctx, image, callbacks_results)


SOLVERS = {"cg": cg, "lsqr": lsqr, "lsmr": lsmr}
