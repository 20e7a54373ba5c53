"""Coil-sharded multi-GPU operator: one process per GPU, ``torch.distributed`` for the plumbing.

The reference has no multi-GPU path; parity is "N-GPU result == 1-GPU result".
Coils (and batched volumes) are independent transforms over the same trajectory, so the coil axis
is partitioned across ranks: every rank holds the plan + sorted points (replicated), its slice of
the sensitivity maps and of the k-space data.  Collectives (NCCL over NVLink on the GPU box, gloo in
the CPU tests) are needed only where the path has a real exchange step:

* SENSE ``adj_op`` / ``data_consistency``: one all-reduce(sum) of the coil-combined image;
* ``pinv_solver`` (``cg`` / ``lsqr`` / ``lsmr``, ``mrinufft_b200.solvers``): scalar all-reduces of the
  k-space norms, and of the image-domain inner products when the iterate is coil-sharded
  (calibrationless); the Lipschitz constant of ``cg`` is broadcast from rank 0;
* ``op`` and calibrationless ``adj_op``: no communication (results stay sharded by coil).

``local_factory`` builds the rank-local operator; the default is ``MRIB200NUFFT``.  The CPU tests
inject a numpy operator to exercise this host logic with ``gloo`` and world_size 2.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs next to its GPU (one process per GPU).

    Page-locked host buffers are placed by first touch, so with every rank's buffers on one socket the
    host<->device copies of the other socket's GPUs cross the inter-socket link.  Best effort: uses the
    PCI address of the device and ``/sys``; returns ``(numa_node, n_cpus)`` or ``None`` when the
    topology is not visible or the process is not allowed on those CPUs (then nothing changes).
    """
    import os
    from pathlib import Path

    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


def coil_slice(n_coils: int, rank: int, world: int):
    """Contiguous, balanced partition of the coil axis: rank r owns [lo, hi)."""
    if n_coils < world:
        raise ValueError(f"cannot shard {n_coils} coils over {world} ranks")
    base, rem = divmod(n_coils, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CoilShardedOperator:
    """SENSE / calibrationless operator whose coils are sharded over the ranks of ``group``.

    Parameters
    ----------
    samples, shape: as for ``MRIB200NUFFT`` (replicated on every rank).
    n_coils: int
        GLOBAL number of coils.
    smaps: array (n_coils_local, *shape) or (n_coils, *shape), optional
        Either this rank's slice or the full set (sliced here).
    group: torch.distributed process group (default: WORLD).
    local_factory: callable(samples, shape, n_coils=, smaps=, **kw) -> operator
    """

    def __init__(self, samples, shape, n_coils, smaps=None, group=None, local_factory=None, **kwargs):
        self.group = group
        if dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:  # a single process: one shard that holds every coil, no collectives
            self.rank, self.world = 0, 1
        self.n_coils = int(n_coils)
        self.lo, self.hi = coil_slice(self.n_coils, self.rank, self.world)
        self.shape = tuple(int(s) for s in shape)
        if smaps is not None and smaps.shape[0] == self.n_coils and self.n_coils != self.hi - self.lo:
            smaps = smaps[self.lo:self.hi]
        if local_factory is None:
            from .operator import MRIB200NUFFT

            local_factory = MRIB200NUFFT
        kwargs.setdefault("squeeze_dims", False)
        self.local = local_factory(samples, shape, n_coils=self.hi - self.lo, smaps=smaps, **kwargs)
        self.uses_sense = smaps is not None

    # -- collectives ------------------------------------------------------------------------
    def _allreduce(self, arr):
        """Sum over ranks, in place for torch tensors; numpy arrays go through a tensor view."""
        if isinstance(arr, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(arr))
            self._allreduce_tensor(t)
            return t.numpy()
        self._allreduce_tensor(arr)
        return arr

    def _allreduce_tensor(self, t: torch.Tensor):
        if t.is_complex():
            dist.all_reduce(torch.view_as_real(t), op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_scalar(self, v: torch.Tensor) -> torch.Tensor:
        """Sum an image-domain scalar over ranks (inner products on coil-sharded iterates)."""
        return self.reduce_img(v)

    # -- operator surface ---------------------------------------------------------------------
    def op(self, image):
        """Image (replicated for SENSE, coil slice otherwise) -> this rank's k-space slice."""
        return self.local.op(image)

    def adj_op(self, ksp_local):
        """This rank's k-space slice -> image; SENSE: all-reduced (identical on every rank).  On the b200
        operator the sum over ranks happens on the device (NCCL), before a host caller's image is copied
        out: a host array never goes back up just to be reduced."""
        loc = self.local
        if self.uses_sense and self.world > 1 and hasattr(loc, "_host_pipeline_applies"):
            loc.check_shape(ksp=ksp_local)
            if loc._host_pipeline_applies(ksp_local):
                img, kind, dev = loc._adj_host_pipelined(ksp_local, keep_on_device=True), "numpy", None
            else:
                ksp, kind, dev = loc._in(ksp_local)
                img = loc._adj_device(ksp)
            self._allreduce_tensor(img)
            return loc._out(loc._safe_squeeze(img), kind, dev)
        img = loc.adj_op(ksp_local)
        if self.uses_sense and self.world > 1:
            img = self._allreduce(img)
        return img

    def data_consistency(self, image, obs_local):
        loc = self.local
        if self.uses_sense and self.world > 1 and hasattr(loc, "_host_pipeline_applies"):
            loc.check_shape(image=image, ksp=obs_local)
            img, kind, dev = loc._in(image)
            obs, _, _ = loc._in(obs_local)
            g = loc._dc_device(img, obs)
            self._allreduce_tensor(g)
            return loc._out(loc._safe_squeeze(g), kind, dev)
        g = loc.data_consistency(image, obs_local)
        if self.uses_sense and self.world > 1:
            g = self._allreduce(g)
        return g

    # -- device-level surface and reductions used by mrinufft_b200.solvers -------------------------
    @property
    def density(self):
        return self.local.density

    @density.setter
    def density(self, value):
        self.local.density = value

    def _op_device(self, img):
        return self.local._op_device(img)

    def _adj_device(self, ksp):
        img = self.local._adj_device(ksp)
        if self.uses_sense and self.world > 1:
            self._allreduce_tensor(img)
        return img

    def _dc_device(self, img, obs):
        g = self.local._dc_device(img, obs)
        if self.uses_sense and self.world > 1:
            self._allreduce_tensor(g)
        return g

    def reduce_ksp(self, v: torch.Tensor) -> torch.Tensor:
        """Sum k-space-domain scalars over ranks (k-space is always sharded by coil)."""
        if self.world == 1:
            return v
        v = v.clone()
        self._allreduce_tensor(v.reshape(-1))
        return v

    def reduce_img(self, v: torch.Tensor) -> torch.Tensor:
        """Sum image-domain scalars over ranks: a no-op for the replicated SENSE image."""
        if self.uses_sense or self.world == 1:
            return v
        v = v.clone()
        self._allreduce_tensor(v.reshape(-1))
        return v

    def get_lipschitz_cst(self, max_iter=10):
        """Rank 0's power-method estimate, broadcast: the single-coil operator is the same on every
        rank but the power method starts from an unseeded random image (base.py:1194), and the CG
        step size must be identical everywhere or the replicated iterates drift apart."""
        lip = float(self.local.get_lipschitz_cst(max_iter))
        if self.world > 1:
            t = torch.tensor([lip], dtype=torch.float64, device=getattr(self.local, "device", "cpu"))
            src = 0 if self.group is None else dist.get_global_rank(self.group, 0)
            dist.broadcast(t, src=src, group=self.group)
            lip = float(t.item())
        return lip

    def pinv_solver(self, ksp_local, optim="lsqr", **kwargs):
        """``pinv_solver`` on this rank's k-space slice (base.py:667-690): device-resident solvers with
        the inner products all-reduced.  SENSE: the image is identical on every rank; calibrationless:
        every rank gets the images of its own coils."""
        from .solvers import SOLVERS

        if optim not in SOLVERS:
            raise ValueError(f"coil-sharded pinv_solver supports {sorted(SOLVERS)}, got {optim!r}")
        if optim == "cg":
            kwargs.setdefault("lipschitz_cst", self.get_lipschitz_cst())
            return SOLVERS[optim](self, ksp_local, reduce_fn=self.reduce_img, reduce_ksp=self.reduce_ksp, **kwargs)
        return SOLVERS[optim](self, ksp_local, reduce_ksp=self.reduce_ksp, reduce_img=self.reduce_img, **kwargs)

    def gather_kspace(self, ksp_local):
        """All-gather the coil-sharded k-space (only if the caller wants it on every rank)."""
        t = ksp_local if isinstance(ksp_local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ksp_local))
        parts = [None] * self.world
        sizes = [coil_slice(self.n_coils, r, self.world) for r in range(self.world)]
        axis = t.ndim - 2
        bufs = []
        for r, (lo, hi) in enumerate(sizes):
            shp = list(t.shape)
            shp[axis] = hi - lo
            bufs.append(torch.empty(shp, dtype=t.dtype, device=t.device))
        if t.is_complex():
            dist.all_gather([torch.view_as_real(b) for b in bufs], torch.view_as_real(t.contiguous()), group=self.group)
        else:
            dist.all_gather(bufs, t.contiguous(), group=self.group)
        out = torch.cat(bufs, dim=axis)
        _ = parts
        return out if isinstance(ksp_local, torch.Tensor) else out.numpy()

    def __getattr__(self, name):
        return getattr(self.local, name)
