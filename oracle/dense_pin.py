"""Independent pin for the oracle: explicit (sparse) matrices built by Kronecker products.

TEST INFRASTRUCTURE ONLY.  Shares with the product nothing but the expression tokenizer
(`parse_expression`) and the YAML reader: the Hamiltonian is assembled on the full 2^n space as
sum over bonds of kron(I, ..., A_i, ..., B_j, ..., I), the symmetry-adapted basis vectors are built
explicitly from the projector  P = 1/|G| sum_g conj(chi(g)) U_g, and the projected matrix is
B^dagger H B.  Used for n <= 16 densely and up to n = 20 as a sparse matrix (substitute pin 1).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from distributed_matvec_b200.expr import parse_expression

_I2 = sp.identity(2, format="csr", dtype=np.complex128)


def _site_op(n: int, site: int, mat2: np.ndarray) -> sp.csr_matrix:
    """Operator acting with mat2 on `site` (bit `site` of the state index) of an n-site system.
    kron order: the LEFTMOST factor is the most significant bit, i.e. site n-1."""
    out = sp.identity(1, format="csr", dtype=np.complex128)
    for s in range(n - 1, -1, -1):
        out = sp.kron(out, sp.csr_matrix(mat2) if s == site else _I2, format="csr")
    return out


def full_hamiltonian(term_specs: list[dict], n: int) -> sp.csr_matrix:
    dim = 1 << n
    H = sp.csr_matrix((dim, dim), dtype=np.complex128)
    cache: dict = {}

    def site_op(site, key, mat):
        k = (site, key)
        if k not in cache:
            cache[k] = _site_op(n, site, mat)
        return cache[k]

    for spec in term_specs:
        if "expression" in spec:
            products = parse_expression(spec["expression"])
            for sites in spec["sites"]:
                for p in products:
                    term = sp.identity(dim, format="csr", dtype=np.complex128) * p.coeff
                    for f in p.factors:
                        term = term @ site_op(int(sites[f.site]), (f.kind, f.comp), f.matrix())
                    H = H + term
        else:
            mat = np.array(spec["matrix"], dtype=np.complex128)
            k = int(mat.shape[0]).bit_length() - 1
            for sites in spec["sites"]:
                # matrix index = (bit_site0 << (k-1)) | ... | bit_site_{k-1}
                rows, cols, vals = [], [], []
                rest_mask = ((1 << n) - 1)
                for st in sites:
                    rest_mask &= ~(1 << int(st))
                for state in range(dim):
                    loc_in = 0
                    for pos, st in enumerate(sites):
                        loc_in |= ((state >> int(st)) & 1) << (k - 1 - pos)
                    for loc_out in range(1 << k):
                        v = mat[loc_out, loc_in]
                        if v != 0:
                            out = state & rest_mask
                            for pos, st in enumerate(sites):
                                out |= ((loc_out >> (k - 1 - pos)) & 1) << int(st)
                            rows.append(out); cols.append(state); vals.append(v)
                H = H + sp.csr_matrix((vals, (rows, cols)), shape=(dim, dim), dtype=np.complex128)
    return H


def _apply_element(perm, flip, n, states: np.ndarray) -> np.ndarray:
    out = np.zeros_like(states)
    for i, j in enumerate(perm):
        out |= ((states >> np.uint64(int(j))) & np.uint64(1)) << np.uint64(i)
    if flip:
        out ^= np.uint64((1 << n) - 1)
    return out


def symmetry_adapted_basis(basis):
    """Returns (representatives ascending uint64[N], norms[N], B sparse [2^n, N]) where column k of B
    is P|r_k> / ||P|r_k>||.  Sector restriction (Hamming weight) is applied to candidates."""
    n = basis.number_sites
    dim = 1 << n
    states = np.arange(dim, dtype=np.uint64)
    if basis.hamming_weight is not None:
        pop = np.array([bin(int(s)).count("1") for s in states])
        states = states[pop == basis.hamming_weight]
    if not basis.requires_projection():
        N = states.shape[0]
        B = sp.csr_matrix((np.ones(N), (states.astype(np.int64), np.arange(N))), shape=(dim, N),
                          dtype=np.complex128)
        return states, np.ones(N), B
    g = basis.group
    G = len(g)
    # P|s> = 1/|G| sum_g conj(chi(g)) |g s>
    images = np.stack([_apply_element(g.perms[e], g.flips[e], n, states) for e in range(G)])  # [G, S]
    orbit_min = images.min(axis=0)
    is_rep = orbit_min == states
    reps, norms, cols_r, cols_c, cols_v = [], [], [], [], []
    for idx in np.nonzero(is_rep)[0]:
        vec: dict[int, complex] = {}
        for e in range(G):
            t = int(images[e, idx])
            vec[t] = vec.get(t, 0j) + np.conj(g.characters[e]) / G
        nrm2 = sum(abs(v) ** 2 for v in vec.values())
        if nrm2 < 1e-20:
            continue
        k = len(reps)
        reps.append(int(states[idx]))
        nrm = np.sqrt(nrm2)
        norms.append(nrm)
        for t, v in vec.items():
            if v != 0:
                cols_r.append(t); cols_c.append(k); cols_v.append(v / nrm)
    N = len(reps)
    B = sp.csr_matrix((cols_v, (cols_r, cols_c)), shape=(dim, N), dtype=np.complex128)
    return np.array(reps, dtype=np.uint64), np.array(norms), B


def projected_hamiltonian(term_specs: list[dict], basis, dense: bool = True):
    """(representatives, norms, H_proj [N, N]) with H_proj = B^dagger H B; dense = False keeps it sparse (CSR), which
    carries the construction to 16 - 20 sites (heisenberg_kagome_16 of BASELINE.json at full size)."""
    H = full_hamiltonian(term_specs, basis.number_sites)
    reps, norms, B = symmetry_adapted_basis(basis)
    Hp = B.conj().T @ (H @ B)
    return reps, norms, (Hp.toarray() if dense else Hp.tocsr())
