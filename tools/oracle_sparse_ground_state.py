#!/usr/bin/env python3
"""At-size pin of the CPU oracle on a model whose whole product is too slow to iterate (no GPU): the oracle's
computeOffDiag (reference src/BatchedOperator.chpl:82-213, group as Benes networks) is run ONCE over every basis state,
the entries are stored as a sparse matrix of the symmetry-adapted basis, its Hermiticity is checked with random vectors,
and its lowest eigenvalue is compared with a value from outside the repository.

heisenberg_square_6x6 (the bench workload: 15 804 956 representatives, |G| = 576, 585 262 534 entries): the
exact-diagonalisation literature gives E0 / N = -0.678872 J (Schulz, Ziman & Poilblanc 1996); the file is in sigma-form
(H = 4 J sum S.S), so the lowest eigenvalue must be 4 x 36 x -0.678872 = -97.757568.

heisenberg_chain_36_symm (63 068 876 representatives, |G| = 144, 1.17e9 entries, about 30 GB of host memory): the
Bethe-ansatz energy of the 36-site ring (tests/bethe.py), -63.904943288257 in sigma units.

Usage:  python tools/oracle_sparse_ground_state.py heisenberg_square_6x6 -97.757568 [threads] > profiles/r02_oracle_6x6_literature.log
        python tools/oracle_sparse_ground_state.py heisenberg_chain_36_symm -63.904943288257 > profiles/r02_oracle_chain36_bethe.log
"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
from scipy.sparse.linalg import LinearOperator, eigsh  # noqa: E402

from oracle import model as omodel  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    name, want = sys.argv[1], float(sys.argv[2])
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    basis, matrix = omodel.load_model(os.path.join(ROOT, "data", name + ".yaml"))
    po.set_num_threads(threads)
    t = time.time()
    reps, _ = po.enumerate_states_parallel(basis, networks=True)
    N = int(reps.shape[0])
    print(f"{name}: {N} representatives, |G| = {len(basis.group)}, enumerated in {time.time() - t:.0f} s", flush=True)
    model = po.Model(matrix, networks=True)
    T = max(1, len(matrix.off_diag))
    L = po.lib()
    chunk = 1 << 15
    starts = list(range(0, N, chunk))

    def column_block(lo):
        """Columns lo .. hi of H: (beta, c) = computeOffDiag(alpha_i, xs = 1) means H[index(beta), i] = c."""
        hi = min(N, lo + chunk)
        n = hi - lo
        alphas = np.ascontiguousarray(reps[lo:hi])
        betas = np.zeros(n * T, dtype=np.uint64)
        coeffs = np.zeros(n * T, dtype=np.complex128)
        keys = np.zeros(n * T, dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.int64)
        total = L.oracle_compute_off_diag(C.byref(model.c), 1, n, alphas, None, 1, betas, coeffs, keys, offsets)
        betas, coeffs = betas[:total], coeffs[:total]
        assert np.abs(coeffs.imag).max(initial=0.0) == 0.0        # real operator, trivial characters
        idx = np.searchsorted(reps, betas)
        assert np.array_equal(reps[np.minimum(idx, N - 1)], betas), "a generated state is not in the basis (DMV:115-118)"
        return idx.astype(np.int32), np.ascontiguousarray(coeffs.real), np.diff(offsets)

    t = time.time()
    with ThreadPoolExecutor(threads) as pool:
        blocks = []
        for k, blk in enumerate(pool.map(column_block, starts)):
            blocks.append(blk)
            if k % 50 == 0:
                print(f"  columns {starts[k]:>9d} ...  {time.time() - t:6.0f} s", flush=True)
    indices = np.concatenate([b[0] for b in blocks])
    data = np.concatenate([b[1] for b in blocks])
    counts = np.concatenate([b[2] for b in blocks])
    del blocks
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    print(f"computeOffDiag over the whole basis: {nnz} entries in {time.time() - t:.0f} s ({threads} threads)", flush=True)
    # rows of this CSR matrix are the COLUMNS of H (transposed storage); H real symmetric <=> A == A^T
    offdiag = sp.csr_matrix((data, indices, indptr), shape=(N, N))
    diag = po.apply_diag(matrix, reps, np.ones(N))
    A = LinearOperator((N, N), dtype=np.float64, matvec=lambda x: offdiag @ x.ravel() + diag * x.ravel())
    rng = np.random.default_rng(3)
    u, v = rng.random(N) - 0.5, rng.random(N) - 0.5
    lhs, rhs = float(u @ A.matvec(v)), float(v @ A.matvec(u))
    print(f"Hermiticity at size: u.(A v) = {lhs:.12f}, v.(A u) = {rhs:.12f}, relative difference "
          f"{abs(lhs - rhs) / max(abs(lhs), 1e-300):.2e}", flush=True)
    t = time.time()
    vals, _ = eigsh(A, k=1, which="SA", tol=1e-10, ncv=24, maxiter=5000)
    e0 = float(vals[0])
    print(f"lowest eigenvalue (eigsh, {time.time() - t:.0f} s) = {e0:.9f};  outside value = {want:.9f};  difference {e0 - want:+.2e} "
          f"({(e0 - want) / basis.number_sites / 4:+.2e} J per site)")


if __name__ == "__main__":
    main()
