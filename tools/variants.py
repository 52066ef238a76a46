#!/usr/bin/env python3
"""Time the product kernel for every (mode, index, dtype) variant of a workload (device-resident vectors,
L2 flushed between iterations, CUDA events).  Usage: python tools/variants.py [workload ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_matvec_b200 import Operator, load_config_from_yaml  # noqa: E402


def main():
    workloads = sys.argv[1:] or ["heisenberg_chain_24"]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name in workloads:
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        op = Operator(matrix)
        if os.environ.get("DMV_CANON"):
            op.set_option("canon", int(os.environ["DMV_CANON"]))
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); op.basis.build(); t1.record(); torch.cuda.synchronize()
        n = op.basis.numberStates()
        op.use_torch_stream()
        if os.environ.get("DMV_ROWS_CTAS"):
            op.set_option("rows_ctas", int(os.environ["DMV_ROWS_CTAS"]))
        if os.environ.get("DMV_ROWS_INDEX"):
            op.set_option("rows_index", int(os.environ["DMV_ROWS_INDEX"]))
        if os.environ.get("DMV_GATHER_WALK"):
            op.set_option("gather_walk", int(os.environ["DMV_GATHER_WALK"]))
        op.set_option("mode", 0)
        nnz = int(op.plan().sum())
        print(f"== {name}: N={n} nnz={nnz} build {t0.elapsed_time(t1):.1f} ms "
              f"orbit(q,t,stages)=({op.info('orbit_n_q')},{op.info('orbit_n_t')},{op.info('orbit_n_stages')}) "
              f"canon_mode={op.info('canon_mode')}", flush=True)
        rng = np.random.default_rng(42)
        for cplx in (True, False):
            x = rng.random(n) - 0.5
            if cplx:
                x = x + 1j * (rng.random(n) - 0.5)
            xd = torch.from_numpy(x).cuda()
            yd = torch.zeros_like(xd)
            ref = None
            sym = bool(op.info("rows_ok"))
            for mode, gather in ((0, -1), (1, -1), (1, 0)):
                for index in (0, -1):
                    op.set_option("mode", mode)
                    op.set_option("gather", gather)
                    if sym:   # symmetric bases: (1, -1) = pipelined k_rows, (1, 0) = queued k_pull
                        op.set_option("rows", gather)
                    op.set_option("index", index)
                    if index == -1 and op.info("index_mode") == 0:
                        continue
                    if mode == 1 and gather == -1 and not (op.info("gather") or op.info("rows")):
                        continue
                    for _ in range(3):
                        op.matvec(xd, yd)
                    torch.cuda.synchronize()
                    times = []
                    for k in range(10):
                        flush.fill_(k)
                        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
                        s.record(); op.matvec(xd, yd); e.record(); torch.cuda.synchronize()
                        times.append(s.elapsed_time(e))
                    op.synchronize()
                    if ref is None:
                        ref = yd.clone()
                    err = float((yd - ref).abs().max() / ref.abs().max())
                    E = 16 if cplx else 8
                    ms = float(np.median(times))
                    gbs = (n * (8 + 2 * E) + nnz * (8 + 2 * E)) / (ms * 1e-3) / 1e9
                    print(f"  {'c128' if cplx else 'f64 '} mode={('gather' if op.info('gather') else ('rows' if op.info('rows') else 'pull')) if mode else 'push'} index_mode={op.info('index_mode')}"
                          f"  median {ms:.4f} ms  min {min(times):.4f} ms  {n / ms / 1e6:.2f} Gstates/s "
                          f"{nnz / ms / 1e6:.1f} Gterms/s  alg {gbs:.0f} GB/s  diff_vs_first {err:.1e}", flush=True)
        if op.info("gather") or True:
            op.set_option("mode", -1); op.set_option("gather", -1); op.set_option("index", -1)
            if op.info("gather"):
                for cplx in (True, False):
                    K = 8
                    X = torch.from_numpy(np.stack([rng.random(n) - 0.5 + (1j * (rng.random(n) - 0.5) if cplx else 0) for _ in range(K)])).cuda()
                    Y = torch.zeros_like(X)
                    for label, fn in (("8 single products", lambda: [op.matvec(X[j], Y[j]) for j in range(K)]),
                                      ("one batch of 8 (2 x 4 vectors)", lambda: op.matvec_batch(X, Y))):
                        for _ in range(2):
                            fn()
                        torch.cuda.synchronize()
                        times = []
                        for k in range(6):
                            flush.fill_(k)
                            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
                            s.record(); fn(); e.record(); torch.cuda.synchronize()
                            times.append(s.elapsed_time(e))
                        ms = float(np.median(times))
                        print(f"  {'c128' if cplx else 'f64 '} {label}: median {ms:.4f} ms  = {ms / K:.4f} ms per vector  "
                              f"{K * n / ms / 1e6:.2f} Gstates/s", flush=True)
        op.close()


if __name__ == "__main__":
    main()
