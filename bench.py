#!/usr/bin/env python3
"""bench.py -- H.x throughput of the B200-native hot path (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--dtype c128|f64]
    python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores

A "step" is one matrix-vector product y <- H x over the whole basis of the workload.
  value        basis states / s, inputs resident in HBM, CUDA events on the launching stream, L2 flushed
               between timed iterations, max over ranks
  e2e          same metric through the public host-buffer call (pinned host x -> C ABI -> host y), H2D and
               D2H inside the timed region
  roofline     algorithmic bytes (N (8 + 2E) + nnz (8 + 2E)) / duration of the dominant
               kernel (k_generate), against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the oracle (OpenMP restatement of the reference algorithm; Chapel toolchain unavailable)
               timed on this box's host cores
Under torchrun (N > 1) the basis is hash-partitioned over the ranks, the (beta, coeff) records are
exchanged with NCCL, scaling is "strong" (same workload for every N).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DEFAULT_WORKLOAD = "heisenberg_chain_24"   # BASELINE.json configs[1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DMV_WORKLOAD", DEFAULT_WORKLOAD))
    ap.add_argument("--dtype", default=os.environ.get("DMV_DTYPE", "c128"), choices=["c128", "f64"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 2 ms; the nvidia-smi
    query of B200_PROFILING.md polls too slowly for a region of a few tens of milliseconds)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if visible:
                try:
                    idx = int(visible.split(",")[self.index])
                except ValueError:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            while not self._stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                self._stop.wait(0.002)
        except Exception as e:  # NVML missing: fall back to one nvidia-smi query
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                a, b = [float(v) for v in out.stdout.strip().split(",")]
                self.sm.append(a)
                self.sm_max = b
            except Exception:
                self.reasons.add(f"unsampled ({type(e).__name__})")

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        time.sleep(0.01)
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons) or ["unsampled"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def measured_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_reference_run(matrix, reps, x, seconds: float, max_iters: int = 5):
    """Time the oracle (OpenMP, all host threads) on the full workload; returns (states/s, iters, threads, y)."""
    from oracle import pyoracle as po
    po.set_num_threads(host_threads())   # torchrun exports OMP_NUM_THREADS=1: use all the host threads anyway
    threads = po.num_threads()
    t_first = time.perf_counter()
    y = po.matvec_blocks(matrix, [reps], [x], num_tasks=threads)[0]
    t_first = time.perf_counter() - t_first
    times = [t_first]
    while sum(times) < seconds and len(times) < max_iters:
        t = time.perf_counter()
        po.matvec_blocks(matrix, [reps], [x], num_tasks=threads)
        times.append(time.perf_counter() - t)
    best = min(times)
    return reps.shape[0] / best, len(times), threads, y, best


def run_reference(args):
    """`--impl reference`: the reference's algorithm on the host cores.  The reference itself (Chapel +
    liblattice_symmetries_haskell) cannot be built in this image, so this is the OpenMP restatement in
    oracle/oracle.c (cpu_baseline.kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401  (only to build the basis on the GPU when one is present)
    from distributed_matvec_b200 import load_config_from_yaml
    from oracle import pyoracle as po
    basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", args.workload + ".yaml"))
    reps = build_representatives_for_cpu(basis, matrix)
    cplx = args.dtype == "c128"
    rng = np.random.default_rng(42)
    x = rng.random(reps.shape[0]) - 0.5
    if cplx:
        x = x + 1j * (rng.random(reps.shape[0]) - 0.5)
    po.set_num_threads(host_threads())
    threads = po.num_threads()
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        po.matvec_blocks(matrix, [reps], [x], num_tasks=threads)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 120:
            break
    ms = 1e3 * sum(times) / len(times)
    value = reps.shape[0] / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "H.x basis states/s", "value": value, "unit": "states/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": args.workload, "basis_states": int(reps.shape[0]), "x": "uniform(-0.5,0.5) seed 42"},
        "cpu_baseline": {"value": value, "unit": "states/s", "cores": threads, "kind": "port",
                         "sample": f"full workload, {len(times)} products"},
        "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_representatives_for_cpu(basis, matrix):
    """Representatives for the CPU legs: from the GPU enumeration when a device is present (fast), else
    from the oracle's own enumeration."""
    import torch
    if torch.cuda.is_available():
        from distributed_matvec_b200 import Operator
        op = Operator(matrix, device=0)
        op.basis.build()
        reps = op.basis.representatives()
        op.close()
        return reps
    from oracle import pyoracle as po
    return po.enumerate_states(basis)[0]


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from distributed_matvec_b200 import DistributedOperator, Operator, load_config_from_yaml
    from distributed_matvec_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", args.workload + ".yaml"))
    cplx = args.dtype == "c128"
    E = 16 if cplx else 8
    t0 = time.perf_counter()
    if world > 1:
        dop = DistributedOperator(matrix, device=local_rank)
        op = dop.op
    else:
        op = Operator(matrix, device=local_rank)
    if os.environ.get("DMV_EXCHANGE"):
        op.set_option("exchange", int(os.environ["DMV_EXCHANGE"]))
    op.basis.build()
    n_local = op.basis.numberStates()
    op.use_torch_stream()
    send_counts = op.plan()
    nnz_local = op.numberTerms()
    build_s = time.perf_counter() - t0

    # synthetic input, recipe of input_for_matvec.py:8,31: uniform(-0.5, 0.5), seed 42 (+ rank)
    rng = np.random.default_rng(42 + rank)
    x_host = rng.random(n_local) - 0.5
    if cplx:
        x_host = x_host + 1j * (rng.random(n_local) - 0.5)
    x_pinned = torch.from_numpy(x_host).pin_memory()
    y_pinned = torch.zeros_like(x_pinned).pin_memory()
    x_dev = x_pinned.cuda(non_blocking=True)
    y_dev = torch.zeros_like(x_dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    totals = torch.tensor([n_local, nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(totals)
    n_total, nnz_total = int(totals[0]), int(totals[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def product_device():
        op.matvec(x_dev, y_dev)

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        product_device()
    barrier()
    op.synchronize()   # surfaces device-side errors of the warm-up

    # ---- timed: K steps, per-step CUDA events on the launching stream, L2 flushed between steps
    launches0 = nat.lib().dmv_launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    gen_ms = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            if world > 1:
                dist.barrier()
            starts[k].record()
            product_device()
            ends[k].record()
        barrier()
    launches = nat.lib().dmv_launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    ms_per_step = float(np.mean(step_ms))
    t = torch.tensor([ms_per_step, float(np.min(step_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, ms_best = float(t[0]), float(t[1])
    op.synchronize()

    # ---- dominant kernel duration: the generate stage of the library's own event timeline
    kern_ms = []
    for k in range(min(args.steps, 5)):
        flush.fill_(k)
        torch.cuda.synchronize()
        op.matvec(x_dev, y_dev)
        torch.cuda.synchronize()
        kern_ms.append(op.timings()["generate(diag+offdiag+local accumulate)"])
    kernel_ms = float(np.mean(kern_ms))

    # ---- e2e: pinned host x -> public call -> host y; wall clock around the blocking call
    for _ in range(2):
        op.matvec(x_pinned.numpy(), y_pinned.numpy())
    barrier()
    e2e_times = []
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        barrier()
        t1 = time.perf_counter()
        op.matvec(x_pinned.numpy(), y_pinned.numpy())
        e2e_times.append(time.perf_counter() - t1)
    e2e_ms = 1e3 * float(np.mean(e2e_times))
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    stage = op.timings()

    peak, peak_kind = measured_peaks()
    bytes_alg_local = n_local * (8 + 2 * E) + nnz_local * (8 + 2 * E)
    achieved = bytes_alg_local / (kernel_ms * 1e-3) / 1e9
    gather = bool(op.info("gather"))
    kernel_name = "k_gather" if gather else "k_generate"
    # compulsory HBM traffic of the row traversal: sigma, x and y once each (the per-term gathers of x are
    # served by L1/L2: neighbouring rows share neighbours)
    bytes_compulsory = n_local * (8 + 2 * E)
    if gather:
        note = ("row traversal without atomics: the per-term 8+2E bytes of the SURVEY 8d model never reach HBM "
                "(x gathers hit L1/L2), so the algorithmic rate exceeds the HBM peak; the kernel is bound by L1/L2 "
                "gather throughput, compulsory HBM traffic is N(8+2E)")
    else:
        note = "scatter form: bound by L2 atomics / random access, not HBM"

    line = {
        "metric": "H.x basis states/s", "value": n_total / (ms_per_step * 1e-3), "unit": "states/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "basis_states": n_total, "off_diag_terms": nnz_total,
                   "terms_per_s": nnz_total / (ms_per_step * 1e-3), "partition": f"hash{world}", "exchange": ("replicated x: NCCL all-gather + row gather" if op.info("replicated") else
                                "peer-direct NVLink stores" if op.info("peer_direct") else
                                ("nccl send/recv" if world > 1 else "none")),
                   "x": "uniform(-0.5,0.5) seed 42", "l2": "flushed between timed iterations (256 MB write)",
                   "ms_best_step": ms_best, "basis_build_s": build_s},
        "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "states/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n_local * E), "d2h_bytes_per_step": int(n_local * E),
                "stages_ms": stage},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(f"{kernel_name}:{args.workload}:{args.dtype}") if world == 1 else None,
                     "peak_kind": peak_kind, "kernel": kernel_name, "kernel_ms": kernel_ms,
                     "algorithmic_bytes": int(bytes_alg_local), "compulsory_hbm_bytes": int(bytes_compulsory),
                     "frac_compulsory": bytes_compulsory / (kernel_ms * 1e-3) / 1e9 / peak, "note": note},
        "clocks": clocks.summary(),
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        reps = op.basis.representatives()
        v, iters, threads, y_cpu, best = cpu_reference_run(matrix, reps, x_host, args.cpu_seconds)
        torch.cuda.synchronize()
        op.matvec(x_dev, y_dev)
        torch.cuda.synchronize()
        err = float(np.abs(y_dev.cpu().numpy() - y_cpu).max() / max(np.abs(y_cpu).max(), 1e-300))
        line["cpu_baseline"] = {"value": v, "unit": "states/s", "cores": threads, "kind": "port",
                                "sample": f"full workload, best of {iters} products ({best * 1e3:.1f} ms)",
                                "max_rel_err_gpu_vs_cpu": err}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
