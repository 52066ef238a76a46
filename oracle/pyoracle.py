"""ctypes front-end of oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.

The high-level helpers mirror the reference's test driver
(reference: test/TestMatrixVectorProduct.chpl:25-60): enumerate states -> hash-partition ->
matrixVectorProduct on P logical locales -> back to global (file) order.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

c128_p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")
u64_p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64_p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32_p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8_p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
f64_p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class OracleModel(C.Structure):
    _fields_ = [
        ("T_off", C.c_int64), ("off_v", C.c_void_p), ("off_m", C.c_void_p), ("off_r", C.c_void_p),
        ("off_x", C.c_void_p), ("off_s", C.c_void_p),
        ("T_diag", C.c_int64), ("diag_v", C.c_void_p), ("diag_m", C.c_void_p), ("diag_r", C.c_void_p),
        ("diag_s", C.c_void_p),
        ("max_off_diag", C.c_int64),
        ("n", C.c_int), ("spin_inversion", C.c_int), ("has_permutations", C.c_int),
        ("state_index_is_identity", C.c_int),
        ("G", C.c_int64), ("perms", C.c_void_p), ("flips", C.c_void_p), ("chars", C.c_void_p),
        ("n_stages", C.c_int64), ("net_delta", C.c_void_p), ("net_masks", C.c_void_p),
    ]


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.oracle_hash64_01.restype = C.c_uint64
        L.oracle_hash64_01.argtypes = [C.c_uint64]
        L.oracle_locale_idx_of.restype = C.c_int
        L.oracle_locale_idx_of.argtypes = [C.c_uint64, C.c_int]
        L.oracle_locale_idx_of_many.restype = None
        L.oracle_locale_idx_of_many.argtypes = [C.c_int64, u64_p, C.c_int, u8_p]
        L.oracle_apply_diag_x1.restype = None
        L.oracle_apply_diag_x1.argtypes = [C.c_int64, c128_p, u64_p, u64_p, u64_p, C.c_int64, u64_p,
                                           f64_p, C.c_void_p, C.c_int]
        L.oracle_apply_off_diag_x1.restype = C.c_int64
        L.oracle_apply_off_diag_x1.argtypes = [C.c_int64, c128_p, u64_p, u64_p, u64_p, u64_p, C.c_int64,
                                               u64_p, u64_p, c128_p, i64_p, C.c_void_p, C.c_int]
        L.oracle_state_info.restype = None
        L.oracle_state_info.argtypes = [C.c_int, C.c_int64, i32_p, u8_p, c128_p, C.c_int64, u64_p, u64_p,
                                        c128_p, f64_p]
        L.oracle_is_representative.restype = None
        L.oracle_is_representative.argtypes = [C.c_int, C.c_int64, i32_p, u8_p, c128_p, C.c_int64, u64_p,
                                               u8_p, f64_p]
        L.oracle_enumerate_states.restype = C.c_int64
        L.oracle_enumerate_states.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_state_index.restype = None
        L.oracle_state_index.argtypes = [C.c_int64, u64_p, C.c_int64, u64_p, i64_p]
        L.oracle_matvec.restype = C.c_int64
        L.oracle_matvec.argtypes = [C.POINTER(OracleModel), C.c_int, i64_p, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int64,
                                    C.c_int]
        L.oracle_compute_off_diag.restype = C.c_int64
        L.oracle_compute_off_diag.argtypes = [C.POINTER(OracleModel), C.c_int, C.c_int64, u64_p,
                                              C.c_void_p, C.c_int, u64_p, c128_p, u8_p, i64_p]
        L.oracle_state_info_networks.restype = None
        L.oracle_state_info_networks.argtypes = [C.c_int, C.c_int64, C.c_int64, i32_p, u64_p, u8_p, c128_p, C.c_int64,
                                                 u64_p, u64_p, c128_p, f64_p]
        L.oracle_enumerate_states_parallel.restype = C.c_int64
        L.oracle_enumerate_states_parallel.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int64,
                                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_matvec_rows.restype = C.c_int64
        L.oracle_matvec_rows.argtypes = [C.POINTER(OracleModel), C.c_int64, u64_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int64, C.c_int, C.c_int64, C.c_int64]
        L.oracle_expected_rows.restype = C.c_int64
        L.oracle_expected_rows.argtypes = [C.POINTER(OracleModel), C.c_int64, u64_p, C.c_void_p, C.c_int, C.c_int64,
                                           i64_p, C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Model:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, op, networks: bool = False):
        """networks = True: also hand the group over as Benes networks (oracle/networks.py) -- the TIMED CPU arm;
        the checker keeps the bit-by-bit group."""
        b = op.basis
        off, diag = op.off_diag, op.diag
        self._keep = [np.ascontiguousarray(a) for a in
                      (off.v, off.m, off.r, off.x, off.s, diag.v, diag.m, diag.r, diag.s)]
        k = self._keep
        m = OracleModel()
        m.T_off, m.off_v, m.off_m, m.off_r, m.off_x, m.off_s = len(off), _ptr(k[0]), _ptr(k[1]), _ptr(k[2]), _ptr(k[3]), _ptr(k[4])
        m.T_diag, m.diag_v, m.diag_m, m.diag_r, m.diag_s = len(diag), _ptr(k[5]), _ptr(k[6]), _ptr(k[7]), _ptr(k[8])
        m.max_off_diag = op.number_off_diag_terms()
        m.n = b.number_sites
        m.spin_inversion = b.spin_inversion
        m.has_permutations = int(b.has_permutation_symmetries())
        m.state_index_is_identity = int(b.is_state_index_identity())
        if b.requires_projection():
            g = b.group
            self.perms = np.ascontiguousarray(g.perms)
            self.flips = np.ascontiguousarray(g.flips)
            self.chars = np.ascontiguousarray(g.characters)
            m.G, m.perms, m.flips, m.chars = len(g), _ptr(self.perms), _ptr(self.flips), _ptr(self.chars)
        else:
            m.G, m.perms, m.flips, m.chars = 0, None, None, None
        m.n_stages, m.net_delta, m.net_masks = 0, None, None
        if networks and b.has_permutation_symmetries():
            from . import networks as nw
            self.net_masks = np.ascontiguousarray(nw.group_networks(b.group))
            self.net_delta = np.array(nw.DELTAS, dtype=np.int32)
            m.n_stages, m.net_delta, m.net_masks = len(nw.DELTAS), _ptr(self.net_delta), _ptr(self.net_masks)
        self.c = m
        self.op = op


def hash64_01(x: int) -> int:
    return int(lib().oracle_hash64_01(C.c_uint64(int(x))))


def locale_idx_of(states: np.ndarray, num_locales: int) -> np.ndarray:
    states = np.ascontiguousarray(states, dtype=np.uint64)
    keys = np.zeros(states.shape[0], dtype=np.uint8)
    lib().oracle_locale_idx_of_many(states.shape[0], states, num_locales, keys)
    return keys


def state_info(basis, alphas: np.ndarray):
    g = basis.group
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    betas = np.zeros(n, dtype=np.uint64)
    chars = np.zeros(n, dtype=np.complex128)
    norms = np.zeros(n, dtype=np.float64)
    lib().oracle_state_info(basis.number_sites, len(g), np.ascontiguousarray(g.perms),
                            np.ascontiguousarray(g.flips), np.ascontiguousarray(g.characters), n, alphas,
                            betas, chars, norms)
    return betas, chars, norms


def is_representative(basis, alphas: np.ndarray):
    g = basis.group
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    flags = np.zeros(n, dtype=np.uint8)
    norms = np.zeros(n, dtype=np.float64)
    lib().oracle_is_representative(basis.number_sites, len(g), np.ascontiguousarray(g.perms),
                                   np.ascontiguousarray(g.flips), np.ascontiguousarray(g.characters), n,
                                   alphas, flags, norms)
    return flags, norms


def enumerate_states(basis):
    """Sequential restatement of ``enumerateStates`` for ONE locale: ascending representatives
    (reference: src/StatesEnumeration.chpl:516-603) and their norms."""
    lo, hi = basis.min_state_estimate(), basis.max_state_estimate()
    fixed = int(basis.is_hamming_weight_fixed())
    if basis.requires_projection():
        g = basis.group
        perms, flips, chars = (np.ascontiguousarray(g.perms), np.ascontiguousarray(g.flips),
                               np.ascontiguousarray(g.characters))
        args = (len(g), _ptr(perms), _ptr(flips), _ptr(chars))
    else:
        args = (0, None, None, None)
    L = lib()
    count = L.oracle_enumerate_states(lo, hi, fixed, basis.number_sites, *args, None, None)
    out = np.zeros(count, dtype=np.uint64)
    norms = np.zeros(count, dtype=np.float64)
    L.oracle_enumerate_states(lo, hi, fixed, basis.number_sites, *args, _ptr(out), _ptr(norms))
    return out, norms


def state_info_networks(basis, alphas: np.ndarray):
    """state_info with the group applied as Benes networks (what the timed CPU arm runs)."""
    from . import networks as nw
    g = basis.group
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    betas = np.zeros(n, dtype=np.uint64)
    chars = np.zeros(n, dtype=np.complex128)
    norms = np.zeros(n, dtype=np.float64)
    lib().oracle_state_info_networks(basis.number_sites, len(g), len(nw.DELTAS), np.array(nw.DELTAS, dtype=np.int32),
                                     np.ascontiguousarray(nw.group_networks(g)), np.ascontiguousarray(g.flips),
                                     np.ascontiguousarray(g.characters), n, alphas, betas, chars, norms)
    return betas, chars, norms


def enumerate_states_parallel(basis, networks: bool = True):
    """``enumerateStates`` for one locale with OpenMP over candidate chunks (same result as enumerate_states);
    networks = True applies the group as Benes networks.  Prepares the input of the CPU arm of bench.py."""
    lo, hi = basis.min_state_estimate(), basis.max_state_estimate()
    fixed = int(basis.is_hamming_weight_fixed())
    keep = []
    if basis.requires_projection():
        g = basis.group
        perms, flips, chars = (np.ascontiguousarray(g.perms), np.ascontiguousarray(g.flips),
                               np.ascontiguousarray(g.characters))
        keep += [perms, flips, chars]
        args = [len(g), _ptr(perms), _ptr(flips), _ptr(chars)]
        if networks:
            from . import networks as nw
            masks = np.ascontiguousarray(nw.group_networks(g))
            delta = np.array(nw.DELTAS, dtype=np.int32)
            keep += [masks, delta]
            args += [len(nw.DELTAS), _ptr(delta), _ptr(masks)]
        else:
            args += [0, None, None]
    else:
        args = [0, None, None, None, 0, None, None]
    L = lib()
    count = L.oracle_enumerate_states_parallel(lo, hi, fixed, basis.number_sites, *args, None, None)
    out = np.zeros(count, dtype=np.uint64)
    norms = np.zeros(count, dtype=np.float64)
    L.oracle_enumerate_states_parallel(lo, hi, fixed, basis.number_sites, *args, _ptr(out), _ptr(norms))
    return out, norms


def state_index(representatives: np.ndarray, spins: np.ndarray) -> np.ndarray:
    spins = np.ascontiguousarray(spins, dtype=np.uint64)
    idx = np.zeros(spins.shape[0], dtype=np.int64)
    lib().oracle_state_index(representatives.shape[0], np.ascontiguousarray(representatives), spins.shape[0],
                             spins, idx)
    return idx


def apply_off_diag(op, alphas: np.ndarray, xs: np.ndarray | None = None):
    off = op.off_diag
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    cap = max(1, n * max(1, len(off)))
    betas = np.zeros(cap, dtype=np.uint64)
    coeffs = np.zeros(cap, dtype=np.complex128)
    offsets = np.zeros(n + 1, dtype=np.int64)
    elt = 1
    if xs is not None:
        elt = 2 if np.iscomplexobj(xs) else 1
        xs = np.ascontiguousarray(xs, dtype=np.complex128 if elt == 2 else np.float64)
    total = lib().oracle_apply_off_diag_x1(len(off), off.v, off.m, off.r, off.x, off.s, n, alphas, betas,
                                           coeffs, offsets, _ptr(xs), elt)
    return betas[:total], coeffs[:total], offsets


def apply_diag(op, alphas: np.ndarray, xs: np.ndarray | None = None) -> np.ndarray:
    d = op.diag
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    elt = 1
    if xs is not None:
        elt = 2 if np.iscomplexobj(xs) else 1
        xs = np.ascontiguousarray(xs, dtype=np.complex128 if elt == 2 else np.float64)
    ys = np.zeros(n * elt, dtype=np.float64)
    lib().oracle_apply_diag_x1(len(d), d.v, d.m, d.r, d.s, n, alphas, ys, _ptr(xs), elt)
    return ys.view(np.complex128) if elt == 2 else ys


def compute_off_diag(op, num_locales: int, alphas: np.ndarray, xs: np.ndarray):
    """One ``BatchedOperator.computeOffDiag`` call (reference: src/BatchedOperator.chpl:82-213)."""
    model = Model(op)
    alphas = np.ascontiguousarray(alphas, dtype=np.uint64)
    n = alphas.shape[0]
    elt = 2 if np.iscomplexobj(xs) else 1
    xs = np.ascontiguousarray(xs, dtype=np.complex128 if elt == 2 else np.float64)
    cap = max(1, n * max(1, len(op.off_diag)))
    betas = np.zeros(cap, dtype=np.uint64)
    coeffs = np.zeros(cap, dtype=np.complex128)
    keys = np.zeros(cap, dtype=np.uint8)
    offsets = np.zeros(n + 1, dtype=np.int64)
    total = lib().oracle_compute_off_diag(C.byref(model.c), num_locales, n, alphas, _ptr(xs), elt, betas,
                                          coeffs, keys, offsets)
    return betas[:total], coeffs[:total], keys[:total], offsets


def partition_by_hash(states: np.ndarray, num_locales: int):
    """masks[i] = owner of the i-th state in global sorted order; blocks stay ascending
    (reference: src/StatesEnumeration.chpl:138-156, 379-395)."""
    masks = locale_idx_of(states, num_locales)
    return masks, [np.ascontiguousarray(states[masks == p]) for p in range(num_locales)]


def block_to_hashed(arr: np.ndarray, masks: np.ndarray, num_locales: int):
    """``arrFromBlockToHashed`` (reference: src/BlockToHashed.chpl:87-208): stable split by owner."""
    return [np.ascontiguousarray(arr[masks == p]) for p in range(num_locales)]


def hashed_to_block(blocks, masks: np.ndarray) -> np.ndarray:
    """``arrFromHashedToBlock`` (reference: src/HashedToBlock.chpl:67-153)."""
    out = np.zeros(masks.shape[0], dtype=blocks[0].dtype)
    for p, blk in enumerate(blocks):
        out[masks == p] = blk
    return out


def matvec_blocks(op, reps_blocks, x_blocks, remote_buffer_size: int = 150000, num_tasks: int = 1,
                  y_blocks=None):
    """``matrixVectorProduct`` on P logical locales (reference: DMV:1072-1093)."""
    model = Model(op)
    P = len(reps_blocks)
    elt = 2 if np.iscomplexobj(x_blocks[0]) else 1
    dt = np.complex128 if elt == 2 else np.float64
    xs = [np.ascontiguousarray(x, dtype=dt) for x in x_blocks]
    if y_blocks is None:
        ys = [np.zeros(r.shape[0], dtype=dt) for r in reps_blocks]   # `similar(x)`: test :37
    else:
        ys = [np.ascontiguousarray(y, dtype=dt) for y in y_blocks]
    reps = [np.ascontiguousarray(r, dtype=np.uint64) for r in reps_blocks]
    sizes = np.array([r.shape[0] for r in reps], dtype=np.int64)
    arr_t = C.c_void_p * P
    rp = arr_t(*[r.ctypes.data for r in reps])
    xp = arr_t(*[x.ctypes.data for x in xs])
    yp = arr_t(*[y.ctypes.data for y in ys])
    st = lib().oracle_matvec(C.byref(model.c), P, sizes, rp, xp, yp, elt, remote_buffer_size, num_tasks)
    if st != 0:
        raise RuntimeError("invalid index: a generated state is not in the basis (DMV:115-118)")
    return ys


def matvec_rows(model: "Model", representatives: np.ndarray, x: np.ndarray, y: np.ndarray, row_lo: int, row_hi: int,
                remote_buffer_size: int = 150000, num_tasks: int = 1):
    """Contributions of the source rows [row_lo, row_hi) of a one-locale product to y (accumulated in place): the
    bounded sample the CPU arm of bench.py times."""
    elt = 2 if np.iscomplexobj(x) else 1
    st = lib().oracle_matvec_rows(C.byref(model.c), representatives.shape[0], representatives, _ptr(x), _ptr(y), elt,
                                  remote_buffer_size, num_tasks, row_lo, row_hi)
    if st != 0:
        raise RuntimeError("invalid index: a generated state is not in the basis (DMV:115-118)")


def expected_rows(op, representatives: np.ndarray, x: np.ndarray, rows: np.ndarray, model: "Model | None" = None):
    """y[rows] of y = H x for a HERMITIAN operator, computed column by column with computeOffDiag on the sampled rows
    only (oracle_expected_rows): the at-size parity check for bases too large for a whole CPU product."""
    model = model or Model(op)
    reps = np.ascontiguousarray(representatives, dtype=np.uint64)
    elt = 2 if np.iscomplexobj(x) else 1
    x = np.ascontiguousarray(x, dtype=np.complex128 if elt == 2 else np.float64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.zeros(rows.shape[0], dtype=x.dtype)
    st = lib().oracle_expected_rows(C.byref(model.c), reps.shape[0], reps, _ptr(x), elt, rows.shape[0], rows, _ptr(out))
    if st != 0:
        raise RuntimeError("invalid index: a generated state is not in the basis (DMV:115-118)")
    return out


def matvec_global(op, representatives: np.ndarray, x: np.ndarray, num_locales: int = 1, **kw) -> np.ndarray:
    """Full test pipeline of test/TestMatrixVectorProduct.chpl:25-60 in global sorted order."""
    masks, reps_blocks = partition_by_hash(representatives, num_locales)
    xb = block_to_hashed(x, masks, num_locales)
    yb = matvec_blocks(op, reps_blocks, xb, **kw)
    return hashed_to_block(yb, masks)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))
