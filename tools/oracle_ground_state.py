#!/usr/bin/env python3
"""At-size pin of the CPU oracle (no GPU): Lanczos on the oracle's whole product of a symmetric chain model, compared with
the Bethe-ansatz ground-state energy of the ring (tests/bethe.py).  heisenberg_chain_32_symm: 4 707 969 representatives,
|G| = 128, about 17 s per product on 8 host threads (oracle_matvec_rows with the group as Benes networks).

Usage:  python tools/oracle_ground_state.py heisenberg_chain_32_symm 32 [threads] [max seconds]  >  profiles/r02_oracle_chain32_bethe.log

The Ritz value of Lanczos is variational (theta_k >= E0 at every step), so a run cut short by `max seconds` still says
something: theta_k - E0(Bethe) stays positive and falls geometrically.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import bethe  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    name, n_sites = sys.argv[1], int(sys.argv[2])
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    max_seconds = float(sys.argv[4]) if len(sys.argv) > 4 else float("inf")
    t_start = time.time()
    basis, matrix = omodel.load_model(os.path.join(ROOT, "data", name + ".yaml"))
    po.set_num_threads(threads)
    t = time.time()
    reps, _ = po.enumerate_states_parallel(basis, networks=True)
    N = reps.shape[0]
    print(f"{name}: {N} representatives, enumerated in {time.time() - t:.1f} s, {threads} threads", flush=True)
    model = po.Model(matrix, networks=True)
    want = 4.0 * bethe.heisenberg_ring_e0(n_sites)
    print(f"Bethe ansatz: E0 = {want:.12f} (sigma units)", flush=True)
    v = np.random.default_rng(7).random(N) - 0.5
    v /= np.linalg.norm(v)
    u = np.zeros(N)
    alphas, betas, beta, last, theta = [], [], 0.0, None, 0.0
    for j in range(300):
        w = np.zeros(N)
        t = time.time()
        po.matvec_rows(model, reps, v, w, 0, N, num_tasks=threads)
        dt = time.time() - t
        a = float(v @ w)
        w -= a * v + beta * u
        alphas.append(a)
        theta = float(np.linalg.eigvalsh(np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1))[0])
        beta = float(np.linalg.norm(w))
        print(f"{j:4d}  theta {theta:.12f}  theta - bethe {theta - want:+.3e}  beta {beta:.3e}  {dt:.1f} s", flush=True)
        if last is not None and abs(theta - last) < 2e-12 * abs(theta):
            break
        if time.time() - t_start + dt > max_seconds:
            print(f"stopped after {j + 1} iterations ({max_seconds:.0f} s allowed): not converged", flush=True)
            break
        last = theta
        betas.append(beta)
        u, v = v, w / beta
    print(f"E0 (oracle, Lanczos) = {theta:.12f}; Bethe = {want:.12f}; relative difference {abs(theta - want) / abs(want):.2e}")


if __name__ == "__main__":
    main()
