#!/usr/bin/env python3
"""The outside pins of tests/test_zz_literature_gpu.py in a few seconds of GPU time, without torch or pytest: dmv_lanczos
through the C ABI on the bench workload and the symmetric chains, against the literature / Bethe ansatz / the oracle's
at-size eigenvalue.  Every result is appended to gpurun_out/quick_pins.log at once (a run cut short keeps what it has).

    CUDA_VISIBLE_DEVICES=0 python tools/quick_gpu_pins.py
"""
import os
import sys
import time

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "0")
os.environ.setdefault("DMV_NO_REBUILD", "1")
T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "quick_pins.log"), "a")


def say(text):
    line = f"[{time.time() - T0:6.1f} s] {text}"
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


say("start")
import bethe  # noqa: E402
from distributed_matvec_b200 import Operator, load_config_from_yaml  # noqa: E402

say("imports done")
ORACLE_E0_6X6 = -97.757589597          # profiles/r02_oracle_6x6_literature.log
CASES = [  # model, complex vectors, outside value (file units), tolerance, source
    ("heisenberg_square_6x6", True, 4 * 36 * -0.678872, 1.5e-4, "literature -0.678872 J per site"),
    ("heisenberg_chain_32_symm", False, 4 * bethe.heisenberg_ring_e0(32), 6e-7, "Bethe ansatz"),
    ("heisenberg_chain_36_symm", True, 4 * bethe.heisenberg_ring_e0(36), 7e-7, "Bethe ansatz"),
    ("heisenberg_square_4x4", False, 4 * 16 * -0.7017802, 8e-6, "literature -0.7017802 J per site"),
    ("heisenberg_chain_24_symm", False, 4 * bethe.heisenberg_ring_e0(24), 5e-7, "Bethe ansatz"),
    ("heisenberg_chain_24", False, 4 * bethe.heisenberg_ring_e0(24), 5e-7, "Bethe ansatz"),
]
failed = 0
for name, cplx, want, tol, source in CASES:
    try:
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        op = Operator(matrix)
        t = time.time()
        op.basis.build()
        n = op.basis.numberStates()
        t_build = time.time() - t
        kernel = "k_gather" if op.info("gather") else ("k_rows" if op.info("rows") else "other")
        t = time.time()
        value, _, iters, res = op.lanczos(max_iters=400, tol=1e-11, complex_vectors=cplx, eigenvector=False)
        ok = abs(value - want) < tol
        extra = ""
        if name == "heisenberg_square_6x6":
            ok = ok and abs(value - ORACLE_E0_6X6) < 1e-6
            extra = f"; oracle's matrix at full size {ORACLE_E0_6X6:.9f} (difference {value - ORACLE_E0_6X6:+.1e})"
        failed += 0 if ok else 1
        say(f"{'OK  ' if ok else 'FAIL'} {name}: {n} states, {kernel}, {'c128' if cplx else 'f64'}, build {t_build:.1f} s, "
            f"{iters} Lanczos iterations in {time.time() - t:.1f} s, residual {res:.1e}: E0 = {value:.9f}; {source}: "
            f"{want:.9f} (difference {value - want:+.1e}){extra}")
        op.close()
    except Exception as e:  # keep going: the later cases are independent
        failed += 1
        say(f"FAIL {name}: {type(e).__name__}: {e}")
say(f"done, {failed} failed")
sys.exit(1 if failed else 0)
