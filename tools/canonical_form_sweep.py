#!/usr/bin/env python3
"""Full-size sweep of the orbit minima (no GPU): EVERY state that one product canonicalises -- alpha ^ x_t for every basis
state alpha and every emitting term t, and every alpha itself (BO:163-212: `ls_hs_state_info` on totalCount + count states)
-- goes through

  * the device functions the CUDA kernels run (csrc/dmv_device.cuh: the square-torus / dihedral canonical form, the
    block-rotation form and the general walk), compiled for the host and evaluated by dmv_debug_compile_group, which also
    cross-checks the three forms against one another state by state, and
  * the oracle's state_info (the group as Benes networks, pinned against the bit-by-bit permutation in tests/),

and the representatives and stabiliser sizes (norms) must agree bit for bit.  The GPU tests compare samples; rare states
(large stabilisers) are what a sample can miss.

Usage:  python tools/canonical_form_sweep.py heisenberg_square_6x6 [threads] [first row] [rows]  > profiles/r02_canonical_form_sweep_6x6.log
"""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from distributed_matvec_b200 import _native as nat  # noqa: E402  (the host-side self-check entry only: no device is used)
from oracle import model as omodel  # noqa: E402
from oracle import networks as nw  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    name = sys.argv[1]
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    basis, matrix = omodel.load_model(os.path.join(ROOT, "data", name + ".yaml"))
    po.set_num_threads(threads)
    t = time.time()
    reps, _ = po.enumerate_states_parallel(basis, networks=True)
    N = int(reps.shape[0])
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else N - first
    g = basis.group
    G = len(g)
    print(f"{name}: {N} representatives, |G| = {G}, enumerated in {time.time() - t:.0f} s; rows [{first}, {first + rows})",
          flush=True)
    perms, flips, chars = (np.ascontiguousarray(g.perms), np.ascontiguousarray(g.flips), np.ascontiguousarray(g.characters))
    masks = np.ascontiguousarray(nw.group_networks(g))
    deltas = np.array(nw.DELTAS, dtype=np.int32)
    bd = nat.BasisDesc()
    bd.number_sites, bd.hamming_weight, bd.spin_inversion, bd.has_permutations = (
        basis.number_sites, -1 if basis.hamming_weight is None else basis.hamming_weight, basis.spin_inversion, 1)
    bd.group_order, bd.perms, bd.flips, bd.characters = G, perms.ctypes.data, flips.ctypes.data, chars.ctypes.data
    L, D = po.lib(), nat.lib()
    ext = np.zeros(16, dtype=np.int64)
    nat.check(D.dmv_debug_compile_group(C.byref(bd), ext.ctypes.data, -2, None, None, None))
    print(f"device orbit program: canon_mode {ext[6]}, k {ext[7]}, R {ext[8]}, torus_mode {ext[12]}, dihedral {ext[15]}",
          flush=True)
    chunk = 1 << 13
    starts = list(range(first, first + rows, chunk))

    def sweep(lo):
        hi = min(first + rows, lo + chunk)
        alphas = np.ascontiguousarray(reps[lo:hi])
        raw, _, _ = po.apply_off_diag(matrix, alphas)             # the term kernel: alpha ^ x_t for every emitting term
        states = np.ascontiguousarray(np.concatenate([raw, alphas]))
        n = states.shape[0]
        o_reps = np.zeros(n, dtype=np.uint64)
        o_chars = np.zeros(n, dtype=np.complex128)
        o_norms = np.zeros(n, dtype=np.float64)
        L.oracle_state_info_networks(basis.number_sites, G, len(nw.DELTAS), deltas, masks, flips, chars, n, states,
                                     o_reps, o_chars, o_norms)
        d_reps = np.zeros(n, dtype=np.uint64)
        d_stab = np.zeros(n, dtype=np.int32)
        nat.check(D.dmv_debug_compile_group(C.byref(bd), None, n, states.ctypes.data, d_reps.ctypes.data,
                                            d_stab.ctypes.data))
        bad = int(np.count_nonzero(d_reps != o_reps))
        bad_norm = int(np.count_nonzero(np.abs(np.sqrt(d_stab / G) - o_norms) > 1e-15))
        return n, bad, bad_norm, int(d_stab.max()), int(np.count_nonzero(d_stab > 1))

    t = time.time()
    total = bad = bad_norm = max_stab = nontrivial = 0
    with ThreadPoolExecutor(threads) as pool:
        for k, (n, b, bn, ms, nt) in enumerate(pool.map(sweep, starts)):
            total += n; bad += b; bad_norm += bn; max_stab = max(max_stab, ms); nontrivial += nt
            if k % 100 == 0:
                print(f"  rows {starts[k]:>9d} ...  {total:>11d} states, {bad} / {bad_norm} mismatches, {time.time() - t:6.0f} s",
                      flush=True)
    print(f"{total} states canonicalised by one product ({rows} rows): representative mismatches {bad}, norm mismatches "
          f"{bad_norm}; states with a non-trivial stabiliser {nontrivial}, largest stabiliser {max_stab}; "
          f"{time.time() - t:.0f} s on {threads} threads")
    sys.exit(1 if bad or bad_norm else 0)


if __name__ == "__main__":
    main()
