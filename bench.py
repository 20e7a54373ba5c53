#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200 NUFFT backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (``config.workload``, BASELINE.json configs[2], the configuration the metric is quoted on):
3-D 256^3 image, 32 coils with sensitivity maps, M = 2^23 samples of a phyllotaxis radial
trajectory (``initialize_3D_phyllotaxis_radial(16384, 512)``), complex64, eps = 1e-6.
A step is one ``op`` (type 2, all coils) + one ``adj_op`` (type 1, all coils, SENSE combine).
Multi-GPU (torchrun, one rank per GPU): weak scaling, every rank owns 32 coils of a 32*N-coil
acquisition; the SENSE adjoint image is summed over ranks with one NCCL all-reduce per step.

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
# profiles/r01_k_rows_stream_full.txt
FP32_PEAK_TFMA = 36.8  # measured: profiles/r01_ffma2_rate.jsonl (fp32_fma_per_s of FFMA2)
NCU_TRAFFIC_GB = {"spread": 39.7, "interp": 42.5}
UNIT = "k-samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=256, help="image size per axis")
    ap.add_argument("--coils", type=int, default=32, help="coils per GPU")
    ap.add_argument("--spokes", type=int, default=16384)
    ap.add_argument("--ns", type=int, default=512, help="samples per spoke")
    ap.add_argument("--traj", default="radial", choices=["radial", "random"])
    ap.add_argument("--cpu-sample-coils", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--spread-method", type=int, default=0)
    ap.add_argument("--interp-method", type=int, default=0)
    ap.add_argument("--fft-method", type=int, default=0, help="0 auto, 1 cuFFT + pad/crop kernels, 2 fused zero-padding-aware passes")
    ap.add_argument("--rows-dbg", type=int, default=0, help="debug switches of the row kernels (timing experiments)")
    return ap.parse_args()


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
    return {
        "workload": f"3D {args.n}^3, {args.coils} coils/GPU with smaps, M={M} "
                    f"({f'phyllotaxis radial {args.spokes}x{args.ns}' if args.traj == 'radial' else 'truncnorm random'}), "
                    "complex64, eps=1e-6, sigma=2, one op + one adj_op per step",
        "coils_per_gpu": args.coils, "n_samples": int(M), "image": [args.n] * 3,
        "parallelism": f"coil-sharded x{n_gpus} (weak: {args.coils} coils per GPU, all-reduce of the SENSE adjoint image)",
        "l2": (f"inputs (k-space {8e-9 * M * args.coils:.2g} GB, smaps {8e-9 * args.n ** 3 * args.coils:.2g} GB, "
               f"grids {64e-9 * args.n ** 3 * args.coils:.2g} GB) against the 126 MB L2"),
    }


# ------------------------------------------------------------------------------------------ GPU leg
def run_b200(args):
    import torch
    import torch.distributed as dist

    import mrinufft
    import mrinufft_b200
    from mrinufft_b200 import _lib

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
        from mrinufft_b200.dist import bind_to_gpu_numa_node

        numa = bind_to_gpu_numa_node(local_rank)
        print(f"[rank {rank}] NUMA binding: {numa}", file=sys.stderr, flush=True)
    if not mrinufft_b200.MRIB200NUFFT.available:
        raise RuntimeError("b200 backend unavailable (libb200nufft.so missing or no GPU): no fallback")

    shape = (args.n,) * 3
    C = args.coils
    traj = make_trajectory(args)
    M = traj.shape[0]
    rng = np.random.default_rng(100 + rank)
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def crandn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, device=dev, generator=g, dtype=torch.float32))

    smaps = crandn(C, *shape)
    smaps /= torch.linalg.norm(smaps, dim=0, keepdim=True)
    op = mrinufft.get_operator("b200")(traj, shape, n_coils=C, smaps=smaps, squeeze_dims=False,
                                       coil_chunk=C)
    plan = op.raw_op.plan
    plan.set_option(0, args.spread_method)
    plan.set_option(1, args.interp_method)
    plan.set_option(2, args.fft_method)
    plan.set_option(3, args.rows_dbg)
    img_d = crandn(1, 1, *shape)
    ksp_d = crandn(1, C, M)

    def step():
        y = op._op_device(img_d)
        x = op._adj_device(ksp_d)
        if world > 1:
            dist.all_reduce(torch.view_as_real(x))
        return y, x

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    kernels, ffts = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = M * C * world / (ms_max * 1e-3) / 1e3

    # per-kernel timings: CUDA events recorded by the library on the stream it launches on (torch's
    # current stream), averaged over the timed steps' worth of extra iterations
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

    # secondary: Toeplitz Gram operator (A^H A through two zero-padding-aware FFTs per coil) against the
    # op + adj_op pair it replaces inside CG-type solvers; device resident, not part of `value`
    extras = {}
    try:
        op.compute_toeplitz_kernel()
        for _ in range(2):
            op._gram_device(img_d)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            op._gram_device(img_d)
        g1.record()
        torch.cuda.synchronize()
        extras["toeplitz_gram_ms"] = g0.elapsed_time(g1) / 3
        op._toeplitz_kernel = None
    except Exception as exc:  # noqa: BLE001
        extras["toeplitz_gram_error"] = str(exc)[:200]

    # secondary: cost of new sample locations (fold + sort + visit-stream rebuild), i.e. what
    # `update_samples` adds to the first transform after it (trajectory-learning loops pay it per step)
    try:
        pts_d = op.raw_op._pts
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            op.raw_op._set_pts(pts_d)
            op._op_device(img_d)
        torch.cuda.synchronize()
        t_with = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for _ in range(3):
            op._op_device(img_d)
        torch.cuda.synchronize()
        t_without = (time.perf_counter() - t0) / 3
        extras["setpts_and_stream_rebuild_ms"] = (t_with - t_without) * 1e3
    except Exception as exc:  # noqa: BLE001
        extras["setpts_error"] = str(exc)[:200]

    # end-to-end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        img_h = torch.empty((1, 1, *shape), dtype=torch.complex64, pin_memory=True)
        ksp_h = torch.empty((1, C, M), dtype=torch.complex64, pin_memory=True)
        img_h.copy_(img_d)
        ksp_h.copy_(ksp_d)
        img_np, ksp_np = img_h.numpy(), ksp_h.numpy()

        def e2e_step():
            y = op.op(img_np)          # H2D image, D2H k-space (numpy out, page-locked)
            x = op.adj_op(ksp_np)      # H2D k-space, D2H image
            if world > 1:
                xt = torch.from_numpy(x).to(dev)
                dist.all_reduce(torch.view_as_real(xt))
                x = xt.cpu().numpy()
            return y, x

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        dt_t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t.item())
        e2e = {"value": M * C * world / dt / 1e3, "unit": UNIT,
               "h2d_bytes_per_step": int(img_np.nbytes + ksp_np.nbytes),
               "d2h_bytes_per_step": int(img_np.nbytes + ksp_np.nbytes),
               "ms_per_step": dt * 1e3}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    Nf = int(np.prod(plan.nf))
    ab = algorithmic_bytes(int(np.prod(shape)), Nf, M, 3, C)
    # dominant own kernel: the row kernel of the spreader or of the interpolator (k_rows<.., SPREAD>)
    dom_name, dom_ms = max(("spread", kt["spread_rows_ms"]), ("interp", kt["interp_rows_ms"]),
                           key=lambda kv: kv[1])
    achieved = ab[dom_name] / (dom_ms * 1e-3) / 1e9
    is_cfg_c = (args.n == 256 and C == 32 and M == 1 << 23 and args.traj == "radial")
    roofline = {
        "bound": "hbm", "kernel": f"k_rows<3,{plan.w},{'true' if dom_name == 'spread' else 'false'},true> ({dom_name})",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` (profiles/)
        "traffic": (NCU_TRAFFIC_GB[dom_name] * 1e9 if is_cfg_c else None),
        "peak_source": peak_src, "kernel_ms": kt, "algorithmic_bytes_per_launch": ab[dom_name],
        "step_algorithmic_bytes": ab["pair_all_coils"],
        "step_frac": ab["pair_all_coils"] / (ms_max * 1e-3) / 1e9 / peak,
        "note": "the row kernels are bound by the FP32 FMA pipe and instruction issue, not by HBM "
                "(ncu: issue 54-56 %, dram 26-31 %); frac is quoted against the HBM floor as the "
                "contract asks, `fp32` is the bound that applies",
    }
    # the bound that does apply to the row kernels: w^3 complex accumulations per sample and coil on the
    # FP32 pipe; peak = packed FFMA2 issue rate measured on this pool (tools/ffma2_rate.cu ->
    # profiles/r01_ffma2_rate.jsonl: 0.495 warp instructions / clk / SM sub-partition)
    fma = 2.0 * M * float(plan.w) ** 3 * C
    roofline["fp32"] = {
        "algorithmic_fma_per_launch": fma, "achieved_tfma_s": fma / (dom_ms * 1e-3) / 1e12,
        "peak_tfma_s": FP32_PEAK_TFMA, "frac": fma / (dom_ms * 1e-3) / 1e12 / FP32_PEAK_TFMA,
        "peak_source": "tools/ffma2_rate.cu, profiles/r01_ffma2_rate.jsonl",
    }

    cpu_baseline = None
    if not args.no_cpu and world == 1:
        cs = args.cpu_sample_coils
        dt, t_setpts, cores = cpu_pair_time(traj, shape, smaps[:cs].cpu().numpy(), cs)
        cpu_baseline = {"value": M * cs / dt / 1e3, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{cs} of {C} coils of the same workload (full {args.n}^3, M={M}), one op+adj_op pair, "
                                  f"float32 finufft-algorithm oracle; {dt:.1f} s, setpts {t_setpts:.1f} s",
                        "seconds": dt}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": workload_config(args, M, world),
        "pairs_per_s": 1e3 / ms_max * world,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(kernels + ffts),
        "gpu_launches_detail": {"own_kernels": int(kernels), "cufft_execs": int(ffts)},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "kernel_width": plan.w, "fine_grid": list(plan.nf), "workspace_gb": plan.workspace_bytes / 1e9,
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
