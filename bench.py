#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200 NUFFT backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (``config.workload``, BASELINE.json configs[2], the configuration the metric is quoted on):
3-D 256^3 image, 32 coils with sensitivity maps, M = 2^23 samples of a phyllotaxis radial
trajectory (``initialize_3D_phyllotaxis_radial(16384, 512)``), complex64, eps = 1e-6.
A step is one ``op`` (type 2, all coils) + one ``adj_op`` (type 1, all coils, SENSE combine).
Multi-GPU (torchrun, one rank per GPU): STRONG scaling of that workload through
``mrinufft_b200.dist.CoilShardedOperator`` -- the 32 coils are sharded, 32 / N per rank, the SENSE
adjoint image is summed over ranks with one NCCL all-reduce per step.  (The weak-scaling figure of
round 1, 32 coils per GPU, is kept under ``extras``.)

Prints ONE JSON line on rank 0 (see the driver contract in the task statement).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "baseline" / "_ref"):
    if _p.exists() and str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

METRIC = "3D 32-coil NUFFT op+adj_op throughput"
# DRAM bytes (read + write) of one launch of the row kernels at cfg-C, `ncu --set full`:
# profiles/r02_rows_class32_full.txt
FP32_PEAK_TFMA = 36.8  # measured: profiles/r01_ffma2_rate.jsonl (fp32_fma_per_s of FFMA2)
NCU_TRAFFIC_GB = {"spread": 40.0, "interp": 42.5}  # profiles/r02_rows_class32_full.txt
UNIT = "k-samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=256, help="image size per axis")
    ap.add_argument("--coils", type=int, default=32, help="coils of the acquisition IN TOTAL (sharded over the GPUs)")
    ap.add_argument("--spokes", type=int, default=16384)
    ap.add_argument("--ns", type=int, default=512, help="samples per spoke")
    ap.add_argument("--traj", default="radial", choices=["radial", "random"])
    ap.add_argument("--cpu-sample-coils", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE's secondary configurations (extras.configs)")
    ap.add_argument("--spread-method", type=int, default=0)
    ap.add_argument("--interp-method", type=int, default=0)
    ap.add_argument("--fft-method", type=int, default=0, help="0 auto, 1 cuFFT + pad/crop kernels, 2 fused zero-padding-aware passes")
    ap.add_argument("--rows-dbg", type=int, default=0, help="debug switches of the row kernels (timing experiments)")
    ap.add_argument("--fft-lookahead", type=int, default=-1, help="L2 prefetch look-ahead of the FFT passes in CTAs (-1: library default)")
    args = ap.parse_args()
    args.total_coils = args.coils
    return args


def make_trajectory(args):
    """(M, 3) float32 in [-0.5, 0.5]; deterministic."""
    M = args.spokes * args.ns
    if args.traj == "radial":
        from mrinufft.trajectories import initialize_3D_phyllotaxis_radial

        t = initialize_3D_phyllotaxis_radial(args.spokes, args.ns)
        return np.ascontiguousarray(t.reshape(-1, 3).astype(np.float32))
    from scipy.stats import truncnorm

    t = truncnorm(-3, 3, loc=0, scale=0.16).rvs(size=(M, 3), random_state=0)
    return np.ascontiguousarray(t.astype(np.float32))


def algorithmic_bytes(N, Nf, M, d, C, s=1):
    """bytes_1 = 8 N (1+s) + 32 N_f + M (8 + 4 d) per coil per transform."""
    b1 = 8 * N * (1 + s) + 32 * Nf + M * (8 + 4 * d)
    return {
        "per_transform_per_coil": b1,
        "pair_all_coils": 2 * C * b1,
        # per kernel, per launch over C coils (DESIGN.md "kernels")
        "spread": C * (8 * M + 8 * Nf) + 4 * d * M,
        "interp": C * (8 * M + 8 * Nf) + 4 * d * M,
        "fft": C * 16 * Nf,
        "pad": C * (8 * N * s + 8 * Nf) + 8 * N,
        "crop": C * (8 * N * s + 8 * Nf) + 8 * N,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU leg
def make_cpu_operator(traj_unit, shape, n_coils, smaps):
    """The reference's CPU path for this workload: its own ``get_operator("finufft")`` when finufft is
    importable; otherwise the reference's ``FourierOperatorCPU`` coil loop (base.py:980-1010, 1142-1152)
    around the finufft-algorithm C/OpenMP port as its ``raw_op`` (finufft is not installable offline).
    Every host thread this process may run on is used, set explicitly (torchrun exports
    OMP_NUM_THREADS=1).  Returns (operator, kind, cores, seconds spent in setpts)."""
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import mrinufft
    from mrinufft.operators.base import FourierOperatorCPU, check_backend

    samples = traj_unit.astype(np.float32)
    t0 = time.perf_counter()
    if check_backend("finufft"):
        op = mrinufft.get_operator("finufft")(samples, shape, n_coils=n_coils, smaps=smaps, squeeze_dims=False,
                                              eps=1e-6, nthreads=cores)
        return op, "reference", cores, time.perf_counter() - t0
    from oracle.c_oracle import PortRawOp, set_threads

    set_threads(cores)

    class PortCPU(FourierOperatorCPU):
        backend = "finufft-port"
        available = True

    raw = PortRawOp(samples * np.float32(2 * np.pi), shape, eps=1e-6, precision="f32", workers=cores)
    op = PortCPU(samples, shape, density=False, n_coils=n_coils, smaps=smaps, raw_op=raw, squeeze_dims=False)
    return op, "port", cores, time.perf_counter() - t0


def cpu_pair_time(op, img, ksp):
    """One op + adj_op pair through the reference's public API, seconds."""
    t0 = time.perf_counter()
    y = op.op(img)
    x = op.adj_op(ksp)
    dt = time.perf_counter() - t0
    assert y.shape[-1] == ksp.shape[-1] and np.isfinite(x.ravel()[0])
    return dt


REFERENCE_BUDGET_S = 200.0  # the whole `--impl reference` run ends within a few minutes


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (see make_cpu_operator) on
    the SAME workload -- all coils with sensitivity maps, full size, one op + one adj_op per step.  A
    step takes about a minute on 16 cores, so the K requested steps are cut short once the time budget
    is spent (at least one full step is always timed; ``steps`` reports how many were), and the warm-up
    is a 2-coil operator on the same trajectory rather than W full steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = (args.n,) * 3
    C = args.total_coils
    traj = make_trajectory(args)
    M = traj.shape[0]
    rng = np.random.default_rng(1)
    smaps = np.empty((C, *shape), np.complex64)
    for c in range(C):
        smaps[c] = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    smaps /= np.sqrt(np.sum(np.abs(smaps) ** 2, axis=0, keepdims=True))
    img = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
    ksp = np.empty((1, C, M), np.complex64)
    for c in range(C):
        ksp[0, c] = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    t_start = time.perf_counter()
    op, kind, cores, t_setpts = make_cpu_operator(traj, shape, C, smaps)
    warm = 0
    if args.warmup > 0:  # page in the libraries and the FFT plans on a 2-coil slice of the same workload
        wop, _, _, _ = make_cpu_operator(traj, shape, 2, smaps[:2])
        cpu_pair_time(wop, img, ksp[:, :2])
        del wop
        warm = 1
    times = []
    for _ in range(max(1, args.steps)):
        times.append(cpu_pair_time(op, img, ksp))
        if time.perf_counter() - t_start + times[-1] > REFERENCE_BUDGET_S:
            break
    dt = float(np.mean(times))
    value = M * C / dt / 1e3
    sample = (f"all {C} coils with smaps, full {args.n}^3 / M={M}, {len(times)} op+adj_op step(s) of {dt:.1f} s "
              f"through {'get_operator(finufft)' if kind == 'reference' else 'FourierOperatorCPU + finufft-algorithm C/OpenMP port'}"
              f", {cores} threads; warm-up = one 2-coil step; setpts {t_setpts:.1f} s")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "complex64", "data": "synthetic", "impl": "reference",
        "reference_impl": ("mrinufft get_operator('finufft')" if kind == "reference" else
                           "mrinufft FourierOperatorCPU coil loop around a C/OpenMP port of finufft's algorithm "
                           "(finufft itself is not installable offline)"),
        "config": workload_config(args, M, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, M, n_gpus):
    C = args.total_coils
    return {
        "workload": f"3D {args.n}^3, {C} coils with smaps, M={M} "
                    f"({f'phyllotaxis radial {args.spokes}x{args.ns}' if args.traj == 'radial' else 'truncnorm random'}), "
                    "complex64, eps=1e-6, sigma=2, one op + one adj_op per step",
        "coils_total": C, "coils_per_gpu": C // max(n_gpus, 1), "n_samples": int(M), "image": [args.n] * 3,
        "parallelism": (f"coil-sharded x{n_gpus} through mrinufft_b200.dist.CoilShardedOperator (strong: {C} coils in "
                        f"total, {C // max(n_gpus, 1)} per GPU; NCCL all-reduce of the SENSE adjoint image)"),
        "value_path": "CoilShardedOperator._op_device / ._adj_device on device-resident tensors (the public op / "
                      "adj_op minus the array-type shim); e2e = CoilShardedOperator.op / .adj_op on host numpy arrays",
        "l2": (f"per-GPU inputs (k-space {8e-9 * M * C / max(n_gpus, 1):.2g} GB, smaps {8e-9 * args.n ** 3 * C / max(n_gpus, 1):.2g} GB, "
               f"grids {64e-9 * args.n ** 3 * C / max(n_gpus, 1):.2g} GB) against the 126 MB L2"),
    }


# ------------------------------------------------------------------------------------------ GPU leg
def timed_steps(torch, dist, world, dev, fn, steps, warmup):
    """W warm-up calls, then exactly K calls between barrier + synchronize on both sides, CUDA events on the
    stream the library launches on (torch's current stream); returns the max over ranks of ms per call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def wall_steps(torch, dist, world, dev, fn, steps, warmup=1):
    """Host-clock version for calls that end with a device -> host copy (max over ranks, seconds per call)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mrinufft  # noqa: F401
    import mrinufft_b200
    from mrinufft_b200 import _lib
    from mrinufft_b200.dist import CoilShardedOperator, bind_to_gpu_numa_node, coil_slice

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner (and any NCCL_DEBUG output) to stdout by default: keep stdout
        # for the one JSON line of the contract
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU: keep it (and the page-locked buffers it touches first) on the GPU's socket
        numa = bind_to_gpu_numa_node(local_rank)
        print(f"[rank {rank}] NUMA binding: {numa}", file=sys.stderr, flush=True)
    if not mrinufft_b200.MRIB200NUFFT.available:
        raise RuntimeError("b200 backend unavailable (libb200nufft.so missing or no GPU): no fallback")

    shape = (args.n,) * 3
    C = args.total_coils
    if C % world:
        raise SystemExit(f"{C} coils do not split evenly over {world} GPUs")
    lo, hi = coil_slice(C, rank, world)
    Cl = hi - lo
    traj = make_trajectory(args)
    M = traj.shape[0]
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def crandn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, device=dev, generator=g, dtype=torch.float32))

    def make_operator(n_local):
        smaps = crandn(n_local, *shape)
        smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True) * np.sqrt(world)
        sop = CoilShardedOperator(traj, shape, n_coils=n_local * world, smaps=smaps, squeeze_dims=False,
                                  coil_chunk=n_local)
        pl = sop.local.raw_op.plan
        pl.set_option(0, args.spread_method)
        pl.set_option(1, args.interp_method)
        pl.set_option(2, args.fft_method)
        pl.set_option(3, args.rows_dbg)
        if args.fft_lookahead >= 0:
            pl.set_option(5, args.fft_lookahead)
        return sop, smaps

    # the product path: this rank's shard of the 32-coil operator (strong scaling: C / N coils per GPU)
    sop, smaps = make_operator(Cl)
    op, plan = sop.local, sop.local.raw_op.plan
    img_d = crandn(1, 1, *shape)
    if world > 1:
        dist.broadcast(torch.view_as_real(img_d), src=0)  # the SENSE image is replicated
    ksp_d = crandn(1, Cl, M)

    def step():
        y = sop._op_device(img_d)     # type 2, this rank's coils: no communication
        x = sop._adj_device(ksp_d)    # type 1 + SENSE combine + NCCL all-reduce of the image over ranks
        return y, x

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_max = timed_steps(torch, dist, world, dev, step, args.steps, 0)
    kernels, ffts = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = M * C / (ms_max * 1e-3) / 1e3

    # per-kernel timings: CUDA events recorded by the library on the stream it launches on (torch's
    # current stream), averaged over a few extra iterations
    plan.enable_timing(True)
    kt = {"spread_ms": [], "interp_ms": [], "fft_ms": [], "grid_ms": [], "spread_rows_ms": [],
          "interp_rows_ms": []}
    for _ in range(min(args.steps, 3)):
        op._op_device(img_d)
        t2 = plan.last_timings()
        op._adj_device(ksp_d)
        t1 = plan.last_timings()
        kt["interp_ms"].append(t2["interp_ms"])
        kt["interp_rows_ms"].append(t2["rows_ms"])
        kt["spread_ms"].append(t1["spread_ms"])
        kt["spread_rows_ms"].append(t1["rows_ms"])
        kt["fft_ms"].append(t2["fft_ms"] + t1["fft_ms"])
        kt["grid_ms"].append(t2["grid_ms"] + t1["grid_ms"])
    plan.enable_timing(False)
    kt = {k: float(np.mean(v)) for k, v in kt.items()}
    cls = plan.rows_class(Cl)
    plan_w, op_nf, ws_gb = plan.w, tuple(plan.nf), plan.workspace_bytes / 1e9

    extras = {}
    if world == 1:
        # secondary: Toeplitz Gram operator (A^H A through two zero-padding-aware FFTs per coil) against the
        # op + adj_op pair it replaces inside CG-type solvers; device resident, not part of `value`
        try:
            op.compute_toeplitz_kernel()
            extras["toeplitz_gram_ms"] = timed_steps(torch, dist, 1, dev, lambda: op._gram_device(img_d), 3, 2)
            op._toeplitz_kernel = None
        except Exception as exc:  # noqa: BLE001
            extras["toeplitz_gram_error"] = str(exc)[:200]
        # secondary: cost of new sample locations (fold + sort + visit-stream rebuild), i.e. what
        # `update_samples` adds to the first transform after it (trajectory-learning loops pay it per step)
        try:
            pts_d = op.raw_op._pts

            def with_setpts():
                op.raw_op._set_pts(pts_d)
                op._op_device(img_d)

            t_with = wall_steps(torch, dist, 1, dev, with_setpts, 3)
            t_without = wall_steps(torch, dist, 1, dev, lambda: op._op_device(img_d), 3)
            extras["setpts_and_stream_rebuild_ms"] = (t_with - t_without) * 1e3
        except Exception as exc:  # noqa: BLE001
            extras["setpts_error"] = str(exc)[:200]

    # end-to-end through the public API with HOST buffers, copies inside the timed region.  Host arrays run as
    # chunks of 8 coils whose PCIe copies overlap the neighbouring chunk's transform (operator.py, `_chunks`).
    e2e = None
    if not args.no_e2e:
        img_h = torch.empty((1, 1, *shape), dtype=torch.complex64, pin_memory=True)
        ksp_h = torch.empty((1, Cl, M), dtype=torch.complex64, pin_memory=True)
        img_h.copy_(img_d)
        ksp_h.copy_(ksp_d)
        img_np, ksp_np = img_h.numpy(), ksp_h.numpy()

        def e2e_step(i_np=img_np, k_np=ksp_np):
            y = sop.op(i_np)          # H2D image, D2H this rank's k-space (numpy out, page-locked)
            x = sop.adj_op(k_np)      # H2D k-space, all-reduce on the device, D2H image
            return y, x

        n_e2e = max(2, min(args.steps, 3))
        dt = wall_steps(torch, dist, world, dev, e2e_step, n_e2e)
        nbytes = int(img_np.nbytes + ksp_np.nbytes)
        e2e = {"value": M * C / dt / 1e3, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "host_memory": "page-locked (numpy views of pinned tensors)",
               "host_chunks": [c1 - c0 for c0, c1 in op._chunks(host=True)]}
        # the same with plain (pageable) numpy arrays: the library page-locks the caller's buffers in place on
        # first sight (cudaHostRegister, cached by address) -- the warm-up call pays for that
        try:
            img_pg, ksp_pg = np.array(img_np), np.array(ksp_np)
            t0 = time.perf_counter()
            e2e_step(img_pg, ksp_pg)
            torch.cuda.synchronize()
            first = time.perf_counter() - t0
            dtp = wall_steps(torch, dist, world, dev, lambda: e2e_step(img_pg, ksp_pg), n_e2e, warmup=0)
            e2e["pageable"] = {"value": M * C / dtp / 1e3, "ms_per_step": dtp * 1e3,
                               "first_call_ms_incl_cudaHostRegister": first * 1e3}
            del img_pg, ksp_pg
        except Exception as exc:  # noqa: BLE001
            e2e["pageable"] = {"error": str(exc)[:200]}
        # where the end-to-end time goes: this rank's PCIe rates with every rank copying at the same time
        try:
            def rate(fn, nb):
                return nb / (wall_steps(torch, dist, world, dev, fn, 2) + 1e-12) / 1e9

            kd = torch.empty_like(ksp_d)
            h2d = rate(lambda: kd.copy_(ksp_h, non_blocking=True), ksp_h.numel() * 8)
            d2h = rate(lambda: ksp_h.copy_(kd, non_blocking=True), ksp_h.numel() * 8)
            del kd
            e2e["pcie_gbs_per_rank_all_ranks_busy"] = {"h2d": h2d, "d2h": d2h}
            e2e["copy_floor_ms"] = nbytes / 1e9 / h2d * 1e3 / 2 + nbytes / 1e9 / d2h * 1e3 / 2
        except Exception as exc:  # noqa: BLE001
            e2e["pcie_error"] = str(exc)[:200]
        del img_h, ksp_h

    # N > 1: the weak-scaling number of round 1 (32 coils PER GPU, a 32 N-coil acquisition) next to the strong one
    if world > 1 and not args.no_weak:
        try:
            del sop, op, plan, smaps, ksp_d
            torch.cuda.empty_cache()
            wop, wsm = make_operator(C)
            wk = crandn(1, C, M)

            def wstep():
                wop._op_device(img_d)
                wop._adj_device(wk)

            wms = timed_steps(torch, dist, world, dev, wstep, max(2, min(args.steps, 3)), 2)
            extras["weak_scaling_32_coils_per_gpu"] = {"ms_per_step": wms, "value": M * C * world / (wms * 1e-3) / 1e3,
                                                        "unit": UNIT, "coils_total": C * world}
            del wop, wsm, wk
        except Exception as exc:  # noqa: BLE001
            extras["weak_scaling_error"] = str(exc)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # secondary configurations of BASELINE.json (A, B, D, E) on this GPU, a few seconds in total
    if world == 1 and not args.no_configs:
        try:
            torch.cuda.empty_cache()
            sys.path.insert(0, str(ROOT / "tools"))
            import bench_configs

            extras["configs"] = bench_configs.run_all(quick=args.n < 256)
        except Exception as exc:  # noqa: BLE001
            extras["configs_error"] = repr(exc)[:300]

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    Nf = int(np.prod(op_nf))
    ab = algorithmic_bytes(int(np.prod(shape)), Nf, M, 3, Cl)
    # dominant own kernel: the row kernel of the spreader or of the interpolator (k_rows<.., SPREAD, .., class>)
    dom_name, dom_ms = max(("spread", kt["spread_rows_ms"]), ("interp", kt["interp_rows_ms"]),
                           key=lambda kv: kv[1])
    achieved = ab[dom_name] / (dom_ms * 1e-3) / 1e9
    is_cfg_c = (args.n == 256 and Cl == 32 and M == 1 << 23 and args.traj == "radial")
    roofline = {
        "bound": "hbm",
        "kernel": f"k_rows<3,{plan_w},{'true' if dom_name == 'spread' else 'false'},{'true' if cls['class'] == 32 else 'false'},"
                  f"{cls['class']}> ({dom_name}, coil class {cls['class']})",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` (profiles/)
        "traffic": (NCU_TRAFFIC_GB[dom_name] * 1e9 if is_cfg_c else None),
        "peak_source": peak_src, "kernel_ms": kt, "algorithmic_bytes_per_launch": ab[dom_name],
        "step_algorithmic_bytes": ab["pair_all_coils"] * world,
        "step_frac": ab["pair_all_coils"] * world / (ms_max * 1e-3) / 1e9 / (peak * world),
        "rows_class": cls,
        "note": "the row kernels are bound by the FP32 FMA pipe and instruction issue, not by HBM; frac is quoted "
                "against the HBM floor as the contract asks, `fp32` is the bound that applies",
    }
    # the bound that does apply to the row kernels: w^3 complex accumulations per sample and coil on the
    # FP32 pipe; peak = packed FFMA2 issue rate measured on this pool (tools/ffma2_rate.cu ->
    # profiles/r01_ffma2_rate.jsonl: 0.495 warp instructions / clk / SM sub-partition)
    fma = 2.0 * M * float(plan_w) ** 3 * Cl
    roofline["fp32"] = {
        "algorithmic_fma_per_launch": fma, "achieved_tfma_s": fma / (dom_ms * 1e-3) / 1e12,
        "peak_tfma_s": FP32_PEAK_TFMA, "frac": fma / (dom_ms * 1e-3) / 1e12 / FP32_PEAK_TFMA,
        "peak_source": "tools/ffma2_rate.cu, profiles/r01_ffma2_rate.jsonl",
    }

    cpu_baseline = None
    if not args.no_cpu and world == 1:
        cs = args.cpu_sample_coils
        rng = np.random.default_rng(1)
        sm = (rng.standard_normal((cs, *shape)) + 1j * rng.standard_normal((cs, *shape))).astype(np.complex64)
        sm /= np.sqrt(np.sum(np.abs(sm) ** 2, axis=0, keepdims=True))
        cop, kind, cores, t_setpts = make_cpu_operator(traj, shape, cs, sm)
        im = (rng.standard_normal((1, 1, *shape)) + 1j * rng.standard_normal((1, 1, *shape))).astype(np.complex64)
        ks = (rng.standard_normal((1, cs, M)) + 1j * rng.standard_normal((1, cs, M))).astype(np.complex64)
        dt = cpu_pair_time(cop, im, ks)
        cpu_baseline = {"value": M * cs / dt / 1e3, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{cs} of {C} coils of the same workload (full {args.n}^3, M={M}) with smaps, one op+adj_op "
                                  f"pair through the reference's FourierOperatorCPU coil loop, float32; {dt:.1f} s, "
                                  f"setpts {t_setpts:.1f} s", "seconds": dt}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": workload_config(args, M, world),
        "pairs_per_s": 1e3 / ms_max,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(kernels + ffts),
        "gpu_launches_detail": {"own_kernels": int(kernels), "cufft_execs": int(ffts)},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "kernel_width": plan_w, "fine_grid": list(op_nf), "workspace_gb": ws_gb,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
