#!/usr/bin/env python3
"""Time dmv_matvec_batch on bases with permutation symmetries: k_rows_batch (up to six doubles per state share one orbit
minimum and one 64-byte look-up per term) against the same vectors one by one through k_rows.  Device-resident vectors,
L2 flushed between calls, CUDA events.  Usage: python tools/batch_timing.py [workload ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_matvec_b200 import Operator, load_config_from_yaml  # noqa: E402


def timed(fn, flush, reps=5):
    t = []
    for k in range(reps + 2):
        flush.fill_(k)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if k >= 2:
            t.append(a.elapsed_time(b))
    return float(np.mean(t)), float(np.min(t))


def main():
    workloads = sys.argv[1:] or ["heisenberg_square_6x6"]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name in workloads:
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        op = Operator(matrix)
        op.basis.build()
        op.use_torch_stream()
        if os.environ.get("DMV_ROWS_CTAS"):
            op.set_option("rows_ctas", int(os.environ["DMV_ROWS_CTAS"]))
        op.set_option("rows_batch_min", 2)
        n = op.basis.numberStates()
        print(f"== {name}: N={n} rows={op.info('rows')}", flush=True)
        rng = np.random.default_rng(42)
        for cplx, ks in ((True, (2, 3)), (False, (4, 6))):
            for k in ks:
                x = rng.random((k, n)) - 0.5
                if cplx:
                    x = x + 1j * (rng.random((k, n)) - 0.5)
                xd = torch.from_numpy(x).cuda()
                yd = torch.zeros_like(xd)
                out = {}
                for batch in (0, -1):
                    op.set_option("rows_batch", batch)
                    mean, best = timed(lambda: op.matvec_batch(xd, yd), flush)
                    out[batch] = (mean, best, yd.clone())
                err = float((out[0][2] - out[-1][2]).abs().max() / out[0][2].abs().max())
                print(f"{'c128' if cplx else 'f64 '} k={k:2d}: one by one {out[0][0]:8.2f} ms, batched {out[-1][0]:8.2f} ms "
                      f"(best {out[-1][1]:.2f}) = {out[0][0] / out[-1][0]:.2f}x, {out[-1][0] / k:.2f} ms per vector, "
                      f"max rel diff {err:.1e}", flush=True)
                del xd, yd, out
        op.close()


if __name__ == "__main__":
    main()
