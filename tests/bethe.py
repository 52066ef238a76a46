"""Bethe-ansatz ground-state energy of the spin-1/2 Heisenberg ring -- an EXACT, independent algorithm (a root search on
N / 2 real numbers) that shares nothing with the product, the oracle or the reference.  Test infrastructure.

H = J sum_i S_i.S_{i+1} on N (even) sites, periodic.  The ground state is the N / 2-magnon state with real rapidities
lambda_j solving (Bethe 1931; Hulthen 1938)

    N arctan(2 lambda_j) = pi I_j + sum_k arctan(lambda_j - lambda_k),   I_j = -(M - 1) / 2, ..., (M - 1) / 2,  M = N / 2

and E0 = J (N / 4 - sum_j 2 / (4 lambda_j^2 + 1)).  N = 12: -5.387390917445 J, N = 16: -7.142296360617 J (the values the
exact-diagonalisation literature quotes).  The sigma-form model files (H = sum sigma.sigma) carry 4 E0.
"""
import numpy as np


def heisenberg_ring_e0(n_sites: int) -> float:
    if n_sites < 2 or n_sites % 2:
        raise ValueError("even number of sites")
    N, M = n_sites, n_sites // 2
    quantum = np.arange(M) - (M - 1) / 2.0
    lam = 0.5 * np.tan(np.pi * quantum / N)
    for _ in range(100000):          # damped fixed-point iteration; converges linearly for every N used here
        new = 0.5 * np.tan((np.pi * quantum + np.arctan(lam[:, None] - lam[None, :]).sum(axis=1)) / N)
        done = np.abs(new - lam).max() < 1e-15
        lam = 0.5 * (new + lam)
        if done:
            break
    residual = N * np.arctan(2 * lam) - np.pi * quantum - np.arctan(lam[:, None] - lam[None, :]).sum(axis=1)
    if np.abs(residual).max() > 1e-12:
        raise RuntimeError("Bethe equations did not converge")
    return float(N / 4.0 - np.sum(2.0 / (4.0 * lam * lam + 1.0)))
