"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle -- the parity tests proper.

Tolerances: bit-exact for representatives, indices, hashes and orbit representatives; y within
1e-10 relative (BASELINE.json north_star), checked with the reference's own criterion shape
|a-b| <= max(atol, rtol*max(|a|,|b|)) (test/TestMatrixVectorProduct.chpl:15-20) at rtol = 1e-12 scaled
by the vector norm to allow for the different summation order of atomics.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from distributed_matvec_b200 import (BatchedOperator, ChapelKernels, EmulatedCluster, Operator, block_to_hashed,  # noqa: E402
                                     hashed_to_block, load_config_from_yaml, locale_idx_of)
from oracle import pyoracle as po  # noqa: E402

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

SMALL = ["heisenberg_chain_4", "heisenberg_chain_6", "heisenberg_chain_8", "heisenberg_chain_10",
         "heisenberg_chain_12", "heisenberg_chain_16", "heisenberg_kagome_12", "heisenberg_kagome_12_symm",
         "heisenberg_kagome_16", "heisenberg_square_4x4", "issue_01"]
MEDIUM = ["heisenberg_chain_20", "heisenberg_chain_24_symm"]


def _load(name):
    return load_config_from_yaml(os.path.join(DATA, name + ".yaml"))


def _x(n, cplx, seed=42):
    rng = np.random.default_rng(seed)   # recipe of input_for_matvec.py:8,31: uniform(-0.5, 0.5)
    x = rng.random(n) - 0.5
    if cplx:
        x = x + 1j * (rng.random(n) - 0.5)
    return x


def _close(a, b, rtol=1e-12, atol=1e-14):
    """The reference's per-element criterion |a - b| <= max(atol, rtol max(|a|, |b|)) (test/TestMatrixVectorProduct.chpl:
    15-20), with the absolute floor in units of the largest element: on symmetric bases the orbit-norm ratios of BO:200
    scale individual terms by up to sqrt(|G|), and so the rounding of a sum that cancels."""
    a, b = np.asarray(a), np.asarray(b)
    floor = atol * max(1.0, float(np.abs(b).max(initial=0.0)))
    return bool(np.all(np.abs(a - b) <= np.maximum(floor, rtol * np.maximum(np.abs(a), np.abs(b)))))


def _recipe_x(n, cplx):
    """x of the reference's generator (input_for_matvec.py:8,31): RandomState(42), rand(N) - 0.5, global sorted order."""
    rs = np.random.RandomState(42)
    x = rs.rand(n) - 0.5
    if cplx:
        x = x + 1j * (rs.rand(n) - 0.5)
    return x


@pytest.fixture(scope="module")
def need_cuda():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")


@pytest.mark.parametrize("name", SMALL + MEDIUM)
def test_enumeration_matches_oracle(need_cuda, name):
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    o_reps, o_norms = po.enumerate_states(basis)
    assert reps.dtype == np.uint64
    assert np.array_equal(reps, o_reps)           # bit-exact, ascending
    if basis.has_permutation_symmetries():
        assert np.allclose(op.basis.norms(), o_norms, rtol=0, atol=1e-15)
    op.close()


@pytest.mark.parametrize("num_ranks", [2, 3, 4])
@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_kagome_12_symm", "heisenberg_chain_16",
                                  "heisenberg_square_4x4"])
def test_enumeration_hash_partition(need_cuda, name, num_ranks):
    basis, matrix = _load(name)
    o_reps, _ = po.enumerate_states(basis)
    masks, blocks = po.partition_by_hash(o_reps, num_ranks)
    cl = EmulatedCluster(matrix, num_ranks).build()
    for r, blk in enumerate(cl.representatives()):
        assert np.array_equal(blk, blocks[r])
    cl.close()


def test_hash_and_locale_index_bit_exact(need_cuda):
    basis, matrix = _load("heisenberg_chain_10")
    op = Operator(matrix)
    rng = np.random.default_rng(7)
    states = rng.integers(0, 2**63, size=100000, dtype=np.uint64)
    states[:4] = [0, 1, 2**64 - 1, 0x8000000000000000]
    for P in (1, 2, 3, 4, 5, 7, 8, 16, 255, 256):
        assert np.array_equal(locale_idx_of(op, states, P), po.locale_idx_of(states, P)), P
    op.close()


@pytest.mark.parametrize("name", ["heisenberg_chain_16", "heisenberg_kagome_16", "heisenberg_chain_24_symm",
                                  "heisenberg_chain_12"])
def test_state_index_bit_exact(need_cuda, name):
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    rng = np.random.default_rng(3)
    probe = np.concatenate([reps, reps ^ np.uint64(1), rng.integers(0, 2**basis.number_sites, 5000, dtype=np.uint64),
                            np.array([0, 2**64 - 1], dtype=np.uint64)])
    got = op.basis.stateIndex(probe)
    want = po.state_index(reps, probe)
    assert np.array_equal(got, want)
    op.close()


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_kagome_12_symm", "issue_01",
                                  "heisenberg_square_4x4", "heisenberg_chain_24_symm", "heisenberg_chain_32_symm",
                                  "heisenberg_square_6x6"])
def test_state_info_matches_oracle(need_cuda, name):
    basis, matrix = _load(name)
    op = Operator(matrix)
    rng = np.random.default_rng(11)
    alphas = rng.integers(0, 2**basis.number_sites, 3000, dtype=np.uint64)
    b, c, n = op.basis.stateInfo(alphas)
    ob, oc, on = po.state_info(basis, alphas)
    assert np.array_equal(b, ob)
    ok = on > 0        # characters are only meaningful for states with non-zero norm
    assert np.allclose(c[ok], oc[ok], atol=1e-15)
    assert np.allclose(n, on, atol=1e-15)
    op.close()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("mode", ["auto", "push", "pull", "pull_queued"])  # pull = k_gather / k_rows where they apply
@pytest.mark.parametrize("name", SMALL + MEDIUM)
def test_local_matvec_matches_oracle(need_cuda, name, cplx, mode):
    """test/TestMatrixVectorProduct.chpl on one locale: host vectors through the C ABI.
    push = the reference's traversal (scatter with atomics), pull = by rows (k_gather where it applies, else
    the queued k_pull), pull_queued = k_pull everywhere; all index kernels."""
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.set_option("mode", {"auto": -1, "push": 0, "pull": 1, "pull_queued": 1}[mode])
    if mode == "pull_queued":
        op.set_option("gather", 0)
        op.set_option("rows", 0)
    op.basis.build()
    reps = op.basis.representatives()
    x = _x(reps.shape[0], cplx)
    y_ref = po.matvec_global(matrix, reps, x, 1)
    for index in (-1, 2, 0):       # auto (Lin tables / identity where they apply), combinadic rank, directory search
        op.set_option("index", index)
        y = op.matvec(x)
        assert _close(y, y_ref), (index, op.info("index_mode"), np.abs(y - y_ref).max())
        # device-resident vectors give the same answer
        xd = torch.from_numpy(x).cuda()
        yd = op.matvec(xd)
        torch.cuda.synchronize()
        assert _close(yd.cpu().numpy(), y_ref)
    if mode == "pull" and op.info("rows"):        # k_rows with the dense index (two-level perfect hash) instead of the table
        assert op.info("rows_dense") == 0
        op.set_option("rows_index", 1)
        assert _close(op.matvec(x), y_ref)
        assert 0.8 * reps.shape[0] <= op.info("rows_dense") <= reps.shape[0]
        op.set_option("rows_index", -1)
        assert _close(op.matvec(x), y_ref)
    gather_applies = not basis.has_permutation_symmetries()    # two-body operators: bit-parallel emit test
    # k_rows: permutation symmetries with trivial characters, real two-body operator
    rows_applies = basis.has_permutation_symmetries() and basis.group.all_characters_trivial
    assert op.info("rows_ok") == (1 if rows_applies else 0)
    if mode == "auto":
        assert op.info("pull") == (1 if (gather_applies or rows_applies) else 0)
        assert op.info("gather") == (1 if gather_applies else 0) and op.info("rows") == (1 if rows_applies else 0)
    else:
        assert op.info("pull") == (0 if mode == "push" else 1)
        assert op.info("gather") == (1 if (mode == "pull" and gather_applies) else 0)
        assert op.info("rows") == (1 if (mode == "pull" and rows_applies) else 0)
    op.close()


def test_rank_index_is_selected_and_bit_exact(need_cuda):
    """The combinadic-rank index kernel must agree bit for bit with the sorted-array search."""
    for option, expected in ((2, {"heisenberg_chain_16": 2, "heisenberg_chain_10": 2, "heisenberg_chain_12": 1,
                                  "heisenberg_kagome_12_symm": 0}),
                             (-1, {"heisenberg_chain_16": 3, "heisenberg_chain_10": 3, "heisenberg_chain_12": 1,
                                   "heisenberg_kagome_12_symm": 0, "heisenberg_kagome_16": 3})):
        for name, expect in expected.items():
            basis, matrix = _load(name)
            op = Operator(matrix)
            op.set_option("index", option)
            op.basis.build()
            assert op.info("index_mode") == expect, (name, option)
            reps = op.basis.representatives()
            rng = np.random.default_rng(5)
            probe = np.concatenate([reps, reps ^ np.uint64(3), rng.integers(0, 2**(basis.number_sites + 1), 4000,
                                                                            dtype=np.uint64)])
            got = op.basis.stateIndex(probe)
            assert np.array_equal(got, po.state_index(reps, probe)), (name, option)
            op.close()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("num_ranks", [2, 3, 4, 8])
@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_chain_12", "heisenberg_kagome_16",
                                  "heisenberg_kagome_12_symm", "heisenberg_square_4x4", "issue_01",
                                  "heisenberg_chain_24_symm"])
def test_emulated_ranks_match_oracle(need_cuda, name, num_ranks, cplx):
    """P logical ranks on one GPU (GASNet-smp analogue): bucketing + accumulate of remote records."""
    basis, matrix = _load(name)
    o_reps, _ = po.enumerate_states(basis)
    masks, blocks = po.partition_by_hash(o_reps, num_ranks)
    x = _x(o_reps.shape[0], cplx)
    y_ref = po.matvec_global(matrix, o_reps, x, num_ranks)
    y_one = po.matvec_global(matrix, o_reps, x, 1)
    assert _close(y_ref, y_one)        # P-invariance of the oracle itself
    cl = EmulatedCluster(matrix, num_ranks).build()
    xb = [torch.from_numpy(b).cuda() for b in block_to_hashed(x, masks, num_ranks)]
    yb = cl.matvec(xb)
    torch.cuda.synchronize()
    y = hashed_to_block([t.cpu().numpy() for t in yb], masks)
    assert _close(y, y_ref), np.abs(y - y_ref).max()
    # the plan is exact: what each rank sends is what the oracle's bucketing produces
    counts = np.array([op.plan() for op in cl.ops])
    for r in range(num_ranks):
        xs = np.ones(blocks[r].shape[0])
        _, _, keys, _ = po.compute_off_diag(matrix, num_ranks, blocks[r], xs)
        assert np.array_equal(counts[r], np.bincount(keys, minlength=num_ranks)), r
    cl.close()


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_kagome_16", "heisenberg_kagome_12_symm",
                                  "heisenberg_square_4x4", "issue_01"])
def test_compute_off_diag_matches_oracle(need_cuda, name):
    """BatchedOperator.computeOffDiag, all three branches, as multisets of (beta, coeff, key)."""
    basis, matrix = _load(name)
    o_reps, _ = po.enumerate_states(basis)
    op = Operator(matrix)
    bo = BatchedOperator(op, o_reps.shape[0])
    for cplx in (False, True):
        xs = _x(o_reps.shape[0], cplx, seed=5)
        n, betas, coeffs, keys = bo.computeOffDiag(o_reps.shape[0], o_reps, xs)
        ob, oc, ok, _ = po.compute_off_diag(matrix, 3 if False else 1, o_reps, xs)
        assert n == ob.shape[0]
        order = np.lexsort((coeffs.imag, coeffs.real, betas))
        oorder = np.lexsort((oc.imag, oc.real, ob))
        assert np.array_equal(betas[order], ob[oorder])
        assert np.allclose(coeffs[order], oc[oorder], rtol=1e-13, atol=1e-15)
    op.close()


def _rows_merged(betas, coeffs, offsets):
    """CSR rows as sorted (beta, coefficient) lists with equal betas merged and zero sums dropped."""
    rows = []
    for i in range(offsets.shape[0] - 1):
        acc = {}
        for b, c in zip(betas[offsets[i]:offsets[i + 1]], coeffs[offsets[i]:offsets[i + 1]]):
            acc[int(b)] = acc.get(int(b), 0.0) + complex(c)
        rows.append(sorted((b, c) for b, c in acc.items() if abs(c) > 1e-15))
    return rows


@pytest.mark.parametrize("name", ["heisenberg_chain_12", "heisenberg_kagome_16", "heisenberg_chain_16",
                                  "three_site", "seven_site_complex", "complex_hopping", "wide_two_magnon"])
def test_chapel_plugin_kernels_match_oracle(need_cuda, name):
    """The reference's plugin table (src/FFI.chpl:233-239): ls_chpl_operator_apply_diag / _apply_off_diag
    (BO:217-275), ls_chpl_enumerate_representatives (SE:588-603), ls_chpl_matrix_vector_product (DMV:1095-1110)
    against the oracle's restatement of the term kernels, enumeration and product."""
    basis, matrix = GENERAL_MODELS[name]() if name in GENERAL_MODELS else _load(name)
    reps, _ = po.enumerate_states(basis)
    op = Operator(matrix)
    k = ChapelKernels(op)
    assert np.array_equal(k.enumerate_representatives(), reps)
    rng = np.random.default_rng(13)
    alphas = np.concatenate([reps[:4000], rng.integers(0, 2**basis.number_sites, 500, dtype=np.uint64)])
    # diagonal: real part of the diagonal matrix element (real(64) output, BO:229-230)
    d = k.operator_apply_diag(alphas)
    d_ref = po.apply_diag(matrix, alphas, np.ones(alphas.shape[0], dtype=np.complex128)).real
    assert np.allclose(d, d_ref, rtol=1e-14, atol=1e-14)
    # off-diagonal: CSR by row; same row pointer semantics, rows compared as merged multisets
    betas, coeffs, offsets = k.operator_apply_off_diag(alphas)
    ob, oc, oo = po.apply_off_diag(matrix, alphas)
    assert offsets.shape == (alphas.shape[0] + 1,) and offsets[0] == 0 and np.all(np.diff(offsets) >= 0)
    assert betas.shape[0] == alphas.shape[0] * op.numberOffDiagTerms()        # full-capacity arrays (BO:250-252)
    got, want = _rows_merged(betas, coeffs, offsets), _rows_merged(ob, oc, oo)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert [b for b, _ in g] == [b for b, _ in w]
        assert np.allclose([c for _, c in g], [c for _, c in w], rtol=1e-14, atol=1e-15)
    # the product entry point (real(64), one vector) when the operator is real
    if not np.iscomplexobj(po.matvec_global(matrix, reps, np.ones(reps.shape[0]), 1)):
        x = _x(reps.shape[0], False, 19)
        assert _close(k.matrix_vector_product(x), po.matvec_global(matrix, reps, x, 1))
    k.close()
    op.close()


@pytest.mark.parametrize("name", ["heisenberg_chain_16", "heisenberg_kagome_12_symm", "heisenberg_chain_10"])
def test_primme_matvec_callback(need_cuda, name):
    """ls_chpl_primme_matvec (src/Diagonalize.chpl:134-162): blockSize columns with leading dimensions >= nLocal,
    contiguous (batched on the GPU) and padded (column by column), against the oracle's product per column."""
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    n = reps.shape[0]
    ck = ChapelKernels(op)
    rng = np.random.default_rng(12)
    for block, pad in ((1, 0), (5, 0), (3, 7)):
        X = np.zeros((block, n + pad))
        X[:, :n] = rng.random((block, n)) - 0.5
        Y = ck.primme_matvec(X, ldy=n + pad)
        for k in range(block):
            y_ref = po.matvec_global(matrix, reps, np.ascontiguousarray(X[k, :n]), 1)
            assert _close(Y[k, :n], y_ref), (block, pad, k)
            assert not np.any(Y[k, n:])
    ck.close()
    op.close()


def test_chapel_plugin_kernels_refuse_projected_bases(need_cuda):
    """BO:224-227, 245-248: bases that require projection are not supported by the apply kernels."""
    basis, matrix = _load("heisenberg_chain_10")
    op = Operator(matrix)
    from distributed_matvec_b200 import _native as nat
    a = np.array([3], dtype=np.uint64)
    out = np.zeros(1)
    assert nat.lib().dmv_apply_diag(op._ctx, 1, a.ctypes.data, out.ctypes.data) != 0
    assert b"projection" in nat.lib().dmv_last_error()
    op.close()


@pytest.mark.parametrize("num_ranks", [1, 2, 3, 8])
def test_block_hashed_redistribution_kernels(need_cuda, num_ranks):
    """arrFromBlockToHashed / arrFromHashedToBlock (src/BlockToHashed.chpl:87, src/HashedToBlock.chpl:67): the
    GPU position + permutation kernels, with the all-to-all played by the host, against the numpy statement."""
    basis, matrix = _load("heisenberg_chain_16")
    op = Operator(matrix)
    reps, _ = po.enumerate_states(basis)
    masks = po.locale_idx_of(reps, num_ranks)
    for dtype in (np.float64, np.complex128, np.uint64):
        arr = reps.copy() if dtype == np.uint64 else _x(reps.shape[0], dtype == np.complex128, 5).astype(dtype)
        want = block_to_hashed(arr, masks, num_ranks)
        bounds = np.linspace(0, reps.shape[0], num_ranks + 1).astype(int)      # contiguous chunks, one per rank
        grouped, counts = [], []
        for r in range(num_ranks):
            m = masks[bounds[r]:bounds[r + 1]]
            c, pos = op.hashed_positions(m, num_ranks)
            assert np.array_equal(c, np.bincount(m, minlength=num_ranks))
            grouped.append(op.permute(arr[bounds[r]:bounds[r + 1]], pos, gather=False))
            counts.append(c)
        got = []
        for q in range(num_ranks):      # the "all-to-all": owner q concatenates its share of every chunk, in chunk order
            parts = [grouped[r][counts[r][:q].sum():counts[r][:q + 1].sum()] for r in range(num_ranks)]
            got.append(np.concatenate(parts))
        for q in range(num_ranks):
            assert np.array_equal(got[q], want[q]), (dtype, q)
        # and back: chunk r takes, from every owner, the elements that fall into it, then un-groups them
        back = np.zeros_like(arr)
        taken = [0] * num_ranks
        for r in range(num_ranks):
            m = masks[bounds[r]:bounds[r + 1]]
            c, pos = op.hashed_positions(m, num_ranks)
            parts = []
            for q in range(num_ranks):
                parts.append(got[q][taken[q]:taken[q] + c[q]])
                taken[q] += c[q]
            back[bounds[r]:bounds[r + 1]] = op.permute(np.concatenate(parts), pos, gather=True)
        assert np.array_equal(back, arr)
    if num_ranks == 1:      # the collective entry points degenerate to the permutation alone
        op.basis.build()
        x = _x(reps.shape[0], True, 7)
        assert np.array_equal(op.block_to_hashed(x, masks), x)
        assert np.array_equal(op.hashed_to_block(x, masks), x)
        xd = torch.from_numpy(x).cuda()
        assert torch.equal(op.hashed_to_block(op.block_to_hashed(xd, masks), masks), xd)
    op.close()


def test_missing_state_is_an_error(need_cuda):
    """DMV:115-118: a generated state that is not in the basis halts."""
    basis, matrix = _load("heisenberg_chain_10")
    reps, _ = po.enumerate_states(basis)
    for mode in (0, 1):
        op = Operator(matrix)
        op.set_option("mode", mode)
        op.basis.uncheckedSetRepresentatives(reps[:-7])     # drop a few states
        x = np.ones(reps.shape[0] - 7)
        with pytest.raises(Exception, match="invalid index"):
            op.matvec(x)
        op.close()


def test_no_diagonal_accumulates_into_y(need_cuda):
    """DMV:1062-1069: without diagonal terms y is not cleared."""
    from distributed_matvec_b200.config import basis_from_dict, operator_from_dict
    n = 8
    bonds = [[i, (i + 1) % n] for i in range(n)]
    basis = basis_from_dict({"number_spins": n, "hamming_weight": 4})
    matrix = operator_from_dict({"terms": [{"expression": "σ⁺₀ σ⁻₁", "sites": bonds},
                                           {"expression": "σ⁻₀ σ⁺₁", "sites": bonds}]}, basis)
    for mode in (0, 1):
        op = Operator(matrix)
        op.set_option("mode", mode)
        op.basis.build()
        reps = op.basis.representatives()
        x = _x(reps.shape[0], False)
        y0 = _x(reps.shape[0], False, seed=9)
        y = op.matvec(x, y0.copy())
        y_ref = po.matvec_blocks(matrix, [reps], [x], y_blocks=[y0.copy()])[0]
        assert _close(y, y_ref)
        op.close()


def test_hermiticity_and_linearity_at_size(need_cuda):
    """Size-independent properties on a basis the oracle would be slow on: <u,Hv> = <Hu,v>, linearity."""
    basis, matrix = _load("heisenberg_chain_24")
    op = Operator(matrix)
    op.basis.build()
    n = op.basis.numberStates()
    assert n == 2704156
    u = torch.from_numpy(_x(n, True, 1)).cuda()
    v = torch.from_numpy(_x(n, True, 2)).cuda()
    results = []
    for mode in (0, 1):
        op.set_option("mode", mode)
        Hu, Hv = op.matvec(u), op.matvec(v)
        lhs, rhs = torch.vdot(u, Hv), torch.vdot(Hu, v)
        assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
        w = op.matvec(2.0 * u - 0.5j * v)
        assert torch.allclose(w, 2.0 * Hu - 0.5j * Hv, rtol=1e-11, atol=1e-11)
        results.append(Hu)
    assert torch.allclose(results[0], results[1], rtol=1e-12, atol=1e-12)   # push == pull
    op.close()


@pytest.mark.parametrize("name", ["heisenberg_chain_16", "heisenberg_chain_10", "heisenberg_kagome_16",
                                  "anisotropic_bonds", "complex_hopping", "heisenberg_chain_24_symm", "no_diagonal",
                                  "heisenberg_square_4x4", "heisenberg_kagome_12_symm", "no_diagonal_symm"])
def test_matvec_batch(need_cuda, name):
    """numVectors > 1 (dmv_matvec_batch): every column equals the single-vector product -- bit for bit on the k_gather
    path (same order of operations per column), within rounding elsewhere -- for 1 .. 9 columns, real and complex,
    device and host pointers; operators without a diagonal accumulate into Y (DMV:1062-1069).  Bases with permutation
    symmetries take device batches through k_rows_batch (up to six doubles per state share one look-up per term)."""
    if name == "no_diagonal_symm":     # translations + parity + spin inversion: the row kernels accumulate into Y
        bonds = [[i, (i + 1) % 12] for i in range(12)]
        basis, matrix = _custom(12, 6, [{"expression": "σ⁺₀ σ⁻₁", "sites": bonds}, {"expression": "σ⁻₀ σ⁺₁", "sites": bonds}],
                                symmetries=[{"permutation": [(i + 1) % 12 for i in range(12)], "sector": 0},
                                            {"permutation": [11 - i for i in range(12)], "sector": 0}], spin_inversion=1)
    elif name == "no_diagonal":
        bonds = [[i, (i + 1) % 8] for i in range(8)]
        basis, matrix = _custom(8, 4, [{"expression": "σ⁺₀ σ⁻₁", "sites": bonds}, {"expression": "σ⁻₀ σ⁺₁", "sites": bonds}])
    else:
        basis, matrix = GENERAL_MODELS[name]() if name in GENERAL_MODELS else _load(name)
    op = Operator(matrix)
    op.basis.build()
    n = op.basis.numberStates()
    gather = bool(op.info("gather"))
    for cplx in (False, True):
        for k in (1, 3, 4, 9):
            X = np.stack([_x(n, cplx, 100 + j) for j in range(k)])
            Y0 = np.stack([_x(n, cplx, 200 + j) for j in range(k)]) if name.startswith("no_diagonal") else np.zeros_like(X)
            singles = np.stack([op.matvec(torch.from_numpy(X[j]).cuda(), torch.from_numpy(Y0[j].copy()).cuda()).cpu().numpy()
                                for j in range(k)])
            Yd = op.matvec_batch(torch.from_numpy(X).cuda(), torch.from_numpy(Y0.copy()).cuda()).cpu().numpy()
            if gather:
                assert np.array_equal(Yd, singles), (name, cplx, k)
            else:
                assert _close(Yd, singles), (name, cplx, k)
            Yh = op.matvec_batch(X, Y0.copy())
            assert _close(Yh, singles), (name, cplx, k)
    if op.info("rows"):               # the same batches vector by vector (option rows_batch = 0)
        op.set_option("rows_batch", 0)
        X = np.stack([_x(n, True, 300 + j) for j in range(5)])
        Y0 = np.stack([_x(n, True, 400 + j) for j in range(5)]) if name.startswith("no_diagonal") else np.zeros_like(X)
        a = op.matvec_batch(torch.from_numpy(X).cuda(), torch.from_numpy(Y0.copy()).cuda()).cpu().numpy()
        op.set_option("rows_batch", -1)
        b = op.matvec_batch(torch.from_numpy(X).cuda(), torch.from_numpy(Y0.copy()).cuda()).cpu().numpy()
        assert _close(a, b), name
    op.close()


def test_rows_batch_at_size_sampled_rows(need_cuda):
    """k_rows_batch at size (heisenberg_chain_32_symm, 4.7 M states; heisenberg_square_6x6, 15.8 M): three complex / six
    real vectors in one call, sampled rows of EVERY vector against the oracle's column-by-column recomputation."""
    for name, cplx, k in (("heisenberg_chain_32_symm", False, 6), ("heisenberg_square_6x6", True, 3)):
        basis, matrix = _load(name)
        po.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        op = Operator(matrix)
        op.basis.build()
        reps = op.basis.representatives()
        n = reps.shape[0]
        rows = np.sort(np.random.default_rng(11).choice(n, size=1024, replace=False))
        X = np.stack([_x(n, cplx, 500 + j) for j in range(k)])
        Y = op.matvec_batch(torch.from_numpy(X).cuda())
        torch.cuda.synchronize()
        got = Y[:, torch.from_numpy(rows).cuda()].cpu().numpy()
        for j in range(k):
            expect = po.expected_rows(matrix, reps, X[j], rows)
            assert _close(got[j], expect), (name, j, np.abs(got[j] - expect).max())
        del Y
        op.close()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matvec_golden.npz")


@pytest.mark.parametrize("name", ["heisenberg_chain_4", "heisenberg_chain_6", "heisenberg_chain_8",
                                  "heisenberg_chain_10", "heisenberg_chain_12", "heisenberg_chain_16",
                                  "heisenberg_chain_24_symm", "heisenberg_kagome_12", "heisenberg_kagome_12_symm",
                                  "heisenberg_kagome_16", "heisenberg_square_4x4"])
def test_golden_vectors(need_cuda, name):
    """The committed fixtures in the layout of the reference's data/matvec/*.h5 (/representatives, /x, /y): /x follows
    the reference generator's RandomState stream, /y is the oracle's (tests/golden/make_golden.py).  Checked with the
    reference's own criterion (test/TestMatrixVectorProduct.chpl:15-20: atol 1e-14, rtol 1e-12)."""
    g = np.load(GOLDEN)
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    assert np.array_equal(op.basis.representatives(), g[name + "/representatives"])      # bit-exact
    want = g[name + "/y"]
    for mode in (-1, 0):
        op.set_option("mode", mode)
        y = op.matvec(g[name + "/x"])
        assert np.all(np.abs(y - want) <= np.maximum(1e-14, 1e-12 * np.maximum(np.abs(y), np.abs(want))) + 1e-13)
    op.close()


def test_golden_digest_chain_20(need_cuda):
    import hashlib
    sys_path = os.path.dirname(GOLDEN)
    import sys
    sys.path.insert(0, sys_path)
    import make_golden as mg
    g = np.load(GOLDEN)
    for name, reps, x in mg.replay():
        if name == "heisenberg_chain_20":
            break
    assert np.array_equal(np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8), g[name + "/x_sha256"])
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    y = op.matvec(x)
    d = g[name + "/digest"]
    assert abs(x.sum() - d[0]) < 1e-9 and abs(y.sum() - d[1]) <= 1e-9 * max(1.0, abs(d[1]))
    assert abs(np.abs(y).max() - d[2]) <= 1e-12 * d[2] and abs(y[0] - d[3]) < 1e-12 and abs(y[-1] - d[4]) < 1e-12
    op.close()


def test_symmetric_properties_at_size(need_cuda):
    """heisenberg_chain_32_symm at full size (4 707 969 states, |G| = 128; the oracle would need minutes): exact
    dimension, Hermiticity in the symmetry-adapted basis, linearity, and the canonical form against the chain walk."""
    basis, matrix = _load("heisenberg_chain_32_symm")
    op = Operator(matrix)
    op.basis.build()
    n = op.basis.numberStates()
    assert n == 4707969
    reps = op.basis.representatives()
    assert np.all(np.diff(reps.astype(np.int64)) > 0)     # ascending, unique
    u = torch.from_numpy(_x(n, True, 1)).cuda()
    v = torch.from_numpy(_x(n, True, 2)).cuda()
    Hu, Hv = op.matvec(u), op.matvec(v)
    lhs, rhs = torch.vdot(u, Hv), torch.vdot(Hu, v)
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    w = op.matvec(1.5 * u + 0.25j * v)
    assert torch.allclose(w, 1.5 * Hu + 0.25j * Hv, rtol=1e-11, atol=1e-11)
    op.set_option("canon", 0)
    assert torch.allclose(op.matvec(u), Hu, rtol=1e-12, atol=1e-12)
    op.set_option("canon", -1)
    for mode, rows in ((0, -1), (1, 0)):                  # scatter with atomics; queued row traversal (k_pull)
        op.set_option("mode", mode)
        op.set_option("rows", rows)
        assert torch.allclose(op.matvec(u), Hu, rtol=1e-12, atol=1e-12)
    op.close()


def test_chain_24_full_vector_at_size(need_cuda):
    """BASELINE configs[1] at full size (2 704 156 states), every element against the oracle's product, x by the
    reference's recipe, f64 (the reference's element type) and c128, row and scatter traversals."""
    basis, matrix = _load("heisenberg_chain_24")
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    assert reps.shape[0] == 2704156
    po.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    for cplx in (False, True):
        x = _recipe_x(reps.shape[0], cplx)
        y_ref = po.matvec_blocks(matrix, [reps], [x], num_tasks=po.num_threads())[0]
        for mode in (-1, 0):
            op.set_option("mode", mode)
            y = op.matvec(torch.from_numpy(x).cuda()).cpu().numpy()
            assert _close(y, y_ref), (cplx, mode, np.abs(y - y_ref).max())
    op.close()


@pytest.mark.parametrize("name,states", [("heisenberg_chain_32_symm", 4707969), ("heisenberg_square_6x6", 15804956),
                                         ("heisenberg_chain_36_symm", 63068876)])
def test_symmetric_products_at_size_sampled_rows(need_cuda, name, states):
    """The symmetric BASELINE configs at full size: 4096 sampled rows of y against the oracle, which recomputes them
    column by column (oracle_expected_rows: computeOffDiag on the sampled sources with the bit-by-bit group), for the
    three forms of the product: rows (k_rows), scatter with atomics (k_generate) and the replicated-x form on two
    logical ranks (every rank holds the whole basis; hash partition of x and y)."""
    basis, matrix = _load(name)
    po.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    assert reps.shape[0] == states
    x = _recipe_x(states, True)
    rows = np.sort(np.random.default_rng(5).choice(states, size=4096, replace=False))
    expect = po.expected_rows(matrix, reps, x, rows)
    xd = torch.from_numpy(x).cuda()
    # k_rows (open-addressing table), k_rows with the dense index (perfect hash), scatter, queued rows
    for mode, rows_opt, rows_index in ((-1, -1, -1), (-1, -1, 1), (0, -1, -1), (1, 0, -1)):
        if name == "heisenberg_chain_36_symm" and mode == 1:
            continue                                  # the queued kernel adds nothing new at this size
        op.set_option("mode", mode)
        op.set_option("rows", rows_opt)
        op.set_option("rows_index", rows_index)
        y = op.matvec(xd)
        torch.cuda.synchronize()
        got = y[torch.from_numpy(rows).cuda()].cpu().numpy()
        assert _close(got, expect), (name, mode, np.abs(got - expect).max())
        del y
    assert op.info("rows_ok") == 1
    op.close()
    del xd
    if name == "heisenberg_chain_36_symm":
        return                                        # two more whole bases of 63 M states: covered by the two above
    P = 2
    masks = po.locale_idx_of(reps, P)
    cl = EmulatedCluster(matrix, P).build()
    xb = [torch.from_numpy(np.ascontiguousarray(x[masks == r])).cuda() for r in range(P)]
    yb = cl.matvec_replicated(xb)
    y = np.zeros(states, dtype=np.complex128)
    for r in range(P):
        y[masks == r] = yb[r].cpu().numpy()
    assert _close(y[rows], expect), np.abs(y[rows] - expect).max()
    cl.close()


def test_host_exchange_views_with_real_records(need_cuda):
    """HostExchangedProduct's tensor views on a basis with a real -1 character (spin_inversion = -1): float64 vectors
    give one double per record, and the views handed to the host exchange must have that layout (Operator.record_width
    asks the library).  The host all-to-all is played in-process between two logical ranks on one GPU."""
    basis, matrix = _load("heisenberg_chain_10")
    o_reps, _ = po.enumerate_states(basis)
    P = 2
    masks, blocks = po.partition_by_hash(o_reps, P)
    for cplx in (False, True):
        x = _x(o_reps.shape[0], cplx, seed=8)
        y_ref = po.matvec_global(matrix, o_reps, x, P)
        cl = EmulatedCluster(matrix, P).build()
        xb = [torch.from_numpy(np.ascontiguousarray(x[masks == r])).cuda() for r in range(P)]
        ys = [torch.zeros_like(t) for t in xb]
        counts = [op.plan() for op in cl.ops]
        outs = []
        for r, op in enumerate(cl.ops):
            width = op.record_width(xb[r])
            assert width == (2 if cplx else 1)
            op.generate(xb[r], ys[r])
            op.synchronize()
            betas, coeffs = op.outgoing_tensors(width)
            total = int(sum(counts[r][q] for q in range(P) if q != r))
            assert betas.numel() == total and coeffs.numel() == total * width
            outs.append((betas.clone(), coeffs.clone(), width))
        for r, (betas, coeffs, width) in enumerate(outs):      # two ranks: everything rank r emits goes to the other one
            dst = 1 - r
            cl.ops[dst].accumulate_tensors(xb[dst], betas, coeffs, ys[dst])
            cl.ops[dst].synchronize()
        y = np.zeros_like(y_ref)
        for r in range(P):
            y[masks == r] = ys[r].cpu().numpy()
        assert _close(y, y_ref), np.abs(y - y_ref).max()
        cl.close()


def test_rank_invariance_at_size(need_cuda):
    """P-invariance (SURVEY.md section 8c pin 4) on heisenberg_chain_20 (184 756 states): one rank, 4 logical ranks with
    the record exchange and 4 logical ranks with the replicated-x form give the same vector after un-hashing."""
    basis, matrix = _load("heisenberg_chain_20")
    one = Operator(matrix)
    one.basis.build()
    reps = one.basis.representatives()
    x = _x(reps.shape[0], True, 11)
    y_one = one.matvec(x)
    one.close()
    P = 4
    masks = po.locale_idx_of(reps, P)
    cl = EmulatedCluster(matrix, P).build()
    assert sum(op.basis.numberStates() for op in cl.ops) == reps.shape[0]
    xb = [torch.from_numpy(b).cuda() for b in block_to_hashed(x, masks, P)]
    y_rec = hashed_to_block([t.cpu().numpy() for t in cl.matvec(xb)], masks)
    y_rep = hashed_to_block([t.cpu().numpy() for t in cl.matvec_replicated(xb)], masks)
    assert _close(y_rec, y_one) and _close(y_rep, y_one)
    cl.close()


def _custom(n, hw, terms, **basis_kw):
    from distributed_matvec_b200.config import basis_from_dict, operator_from_dict
    basis = basis_from_dict({"number_spins": n, "hamming_weight": hw, **basis_kw})
    return basis, operator_from_dict({"terms": terms}, basis)


GENERAL_MODELS = {
    # 3-site terms: support of 3 bits -> general LUT path (not bit-parallel)
    "three_site": lambda: _custom(10, None, [
        {"expression": "σˣ₀ σˣ₁ σᶻ₂", "sites": [[i, (i + 1) % 10, (i + 2) % 10] for i in range(10)]},
        {"expression": "0.7 × σᶻ₀ σᶻ₁", "sites": [[i, (i + 1) % 10] for i in range(10)]}]),
    # 7-site string: support of 7 bits -> term-by-term (generic) path; sigma^y makes coefficients complex
    "seven_site_complex": lambda: _custom(9, None, [
        {"expression": "σʸ₀ σᶻ₁ σᶻ₂ σᶻ₃ σᶻ₄ σᶻ₅ σˣ₆", "sites": [[(i + k) % 9 for k in range(7)] for i in range(9)]},
        {"expression": "σˣ₀", "sites": [[i] for i in range(9)]}]),
    # single-site field + hopping, fixed magnetisation, complex hopping amplitude
    "complex_hopping": lambda: _custom(10, 5, [
        {"expression": "σ⁺₀ σ⁻₁", "sites": [[i, (i + 1) % 10] for i in range(10)]},
        {"expression": "σ⁻₀ σ⁺₁", "sites": [[i, (i + 1) % 10] for i in range(10)]},
        {"expression": "0.3j × σ⁺₀ σ⁻₁", "sites": [[i, (i + 2) % 10] for i in range(10)]},
        {"expression": "-0.3j × σ⁻₀ σ⁺₁", "sites": [[i, (i + 2) % 10] for i in range(10)]},
        {"expression": "σᶻ₀", "sites": [[0], [3]]}]),
    # two-body, bond-dependent couplings: bit-parallel emit test with a non-uniform coefficient table
    "anisotropic_bonds": lambda: _custom(12, 6, [
        {"expression": "σ⁺₀ σ⁻₁", "sites": [[i, (i + 1) % 12] for i in range(12)]},
        {"expression": "σ⁻₀ σ⁺₁", "sites": [[i, (i + 1) % 12] for i in range(12)]},
        {"expression": "0.37 × σ⁺₀ σ⁻₁", "sites": [[i, (i + 3) % 12] for i in range(0, 12, 2)]},
        {"expression": "0.37 × σ⁻₀ σ⁺₁", "sites": [[i, (i + 3) % 12] for i in range(0, 12, 2)]},
        {"expression": "1.3 × σᶻ₀ σᶻ₁", "sites": [[i, (i + 1) % 12] for i in range(12)]}]),
    # more than 32 sites / groups: the 64-bit row path of k_gather (two magnons on a 34-site ring)
    "wide_two_magnon": lambda: _custom(34, 2, [
        {"expression": "σˣ₀ σˣ₁", "sites": [[i, (i + 1) % 34] for i in range(34)]},
        {"expression": "σʸ₀ σʸ₁", "sites": [[i, (i + 1) % 34] for i in range(34)]},
        {"expression": "σᶻ₀ σᶻ₁", "sites": [[i, (i + 1) % 34] for i in range(34)]}]),
    # spin inversion (BatchedOperator branch b) with a next-nearest-neighbour zz coupling, odd sector
    "inversion_two_body": lambda: _custom(12, 6, [
        {"expression": "σˣ₀ σˣ₁", "sites": [[i, (i + 1) % 12] for i in range(12)]},
        {"expression": "σʸ₀ σʸ₁", "sites": [[i, (i + 1) % 12] for i in range(12)]},
        {"expression": "0.5 × σᶻ₀ σᶻ₁", "sites": [[i, (i + 2) % 12] for i in range(12)]}], spin_inversion=-1),
    # translation symmetry with a complex character (momentum sector 1)
    "momentum_sector": lambda: _custom(10, 5, [
        {"expression": "σˣ₀ σˣ₁", "sites": [[i, (i + 1) % 10] for i in range(10)]},
        {"expression": "σʸ₀ σʸ₁", "sites": [[i, (i + 1) % 10] for i in range(10)]},
        {"expression": "σᶻ₀ σᶻ₁", "sites": [[i, (i + 1) % 10] for i in range(10)]}],
        symmetries=[{"permutation": [(i + 1) % 10 for i in range(10)], "sector": 1}]),
}


@pytest.mark.parametrize("model", sorted(GENERAL_MODELS))
def test_general_operators_all_paths(need_cuda, model):
    """Operators beyond two-body real Heisenberg: every generation path (bit-parallel, LUT walk, term by
    term), push and pull, real and complex vectors, 1 and 3 ranks."""
    basis, matrix = GENERAL_MODELS[model]()
    reps, _ = po.enumerate_states(basis)
    for cplx in (False, True):
        x = _x(reps.shape[0], cplx, seed=17)
        y_ref = po.matvec_global(matrix, reps, x, 1)
        for mode in (0, 1):
            for bitparallel in (1, 0):
                op = Operator(matrix)
                op.set_option("mode", mode)
                op.set_option("bitparallel", bitparallel)
                op.basis.build()
                assert np.array_equal(op.basis.representatives(), reps)
                y = op.matvec(x)
                assert _close(y, y_ref), (model, cplx, mode, bitparallel, np.abs(y - y_ref).max())
                op.close()
        masks, blocks = po.partition_by_hash(reps, 3)
        cl = EmulatedCluster(matrix, 3).build()
        yb = cl.matvec([torch.from_numpy(b).cuda() for b in block_to_hashed(x, masks, 3)])
        torch.cuda.synchronize()
        assert _close(hashed_to_block([t.cpu().numpy() for t in yb], masks), y_ref)
        cl.close()


def test_gather_kernel_variants(need_cuda):
    """Which k_gather specialisation each operator gets (32-bit rows, LUT-free uniform coefficient), and that
    every one of them reproduces the oracle for real and complex vectors, host and device pointers."""
    expect = {   # name: (gather applies, narrow, uniform coefficient)
        "anisotropic_bonds": (1, 1, 0), "wide_two_magnon": (1, 0, 1), "inversion_two_body": (1, 1, 1),
        "complex_hopping": (1, 1, 0), "three_site": (0, 1, 0)}
    for model, (applies, narrow, uniform) in expect.items():
        basis, matrix = GENERAL_MODELS[model]()
        reps, _ = po.enumerate_states(basis)
        op = Operator(matrix)
        op.basis.build()
        assert (op.info("gather"), op.info("gather_narrow")) == (applies, narrow), model
        if applies:
            assert op.info("gather_uniform") == uniform, model
        for cplx in (False, True):
            x = _x(reps.shape[0], cplx, seed=23)
            y_ref = po.matvec_global(matrix, reps, x, 1)
            y = op.matvec(x)
            assert _close(y, y_ref), (model, cplx, np.abs(y - y_ref).max())
            yd = op.matvec(torch.from_numpy(x).cuda()).cpu().numpy()
            if applies:
                assert np.array_equal(yd, y), model    # no atomics: bit-reproducible
            else:
                assert _close(yd, y_ref), model
        op.close()


def test_gather_is_bit_reproducible_and_chunked_d2h(need_cuda):
    """k_gather writes every y element once: repeated products are bit-identical, and the host-pointer call
    (row chunks with overlapped D2H copies) returns exactly the device result."""
    basis, matrix = _load("heisenberg_chain_20")
    op = Operator(matrix)
    op.basis.build()
    n = op.basis.numberStates()
    assert n >= 1 << 16 and op.info("gather") == 1
    x = _x(n, True, 3)
    xd = torch.from_numpy(x).cuda()
    y1 = op.matvec(xd).cpu().numpy()
    y2 = op.matvec(xd).cpu().numpy()
    y_host = op.matvec(x)
    assert np.array_equal(y1, y2) and np.array_equal(y1, y_host)
    op.close()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("num_ranks", [2, 3, 8])
@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_chain_12", "heisenberg_kagome_16",
                                  "heisenberg_chain_16", "anisotropic_bonds", "wide_two_magnon",
                                  "complex_hopping", "heisenberg_kagome_12_symm", "heisenberg_square_4x4",
                                  "issue_01", "heisenberg_chain_24_symm", "three_site", "momentum_sector"])
def test_replicated_x_product_matches_oracle(need_cuda, name, num_ranks, cplx):
    """The replicated-x form of the distributed product (all-gather of x + row traversal: k_gather, or the queued
    k_pull for bases with permutation symmetries / general operators) on P logical ranks: same hash partition,
    same result as the oracle's P-rank product."""
    basis, matrix = GENERAL_MODELS[name]() if name in GENERAL_MODELS else _load(name)
    o_reps, _ = po.enumerate_states(basis)
    masks, blocks = po.partition_by_hash(o_reps, num_ranks)
    x = _x(o_reps.shape[0], cplx, seed=31)
    y_ref = po.matvec_global(matrix, o_reps, x, num_ranks)
    cl = EmulatedCluster(matrix, num_ranks).build()
    xb = [torch.from_numpy(b).cuda() for b in block_to_hashed(x, masks, num_ranks)]
    yb = cl.matvec_replicated(xb)
    y = hashed_to_block([t.cpu().numpy() for t in yb], masks)
    assert _close(y, y_ref), np.abs(y - y_ref).max()
    assert cl.ops[0].info("global_states") == o_reps.shape[0]
    # same answer as the record-exchange form on the same ranks
    y_push = hashed_to_block([t.cpu().numpy() for t in cl.matvec(xb)], masks)
    assert _close(y, y_push)
    cl.close()


@pytest.mark.parametrize("name,mode", [("heisenberg_chain_24_symm", 2), ("heisenberg_square_4x4", 1),
                                       ("heisenberg_kagome_12_symm", 0), ("heisenberg_chain_32_symm", 2),
                                       ("heisenberg_square_6x6", 1)])
def test_block_rotation_canonical_form(need_cuda, name, mode):
    """Orbit minima through the canonical form of the translation subgroup (no walk over its elements) are bit-identical
    to the chain walk: state_info against the oracle, and the product with the canonical form on and off."""
    basis, matrix = _load(name)
    op = Operator(matrix)
    assert op.info("canon_mode") == mode
    rng = np.random.default_rng(29)
    alphas = rng.integers(0, 2**basis.number_sites, 4000, dtype=np.uint64)
    alphas[:4] = [0, 2**basis.number_sites - 1, 0x5555555555555555 & (2**basis.number_sites - 1), 1]
    want = po.state_info(basis, alphas)[0]
    if basis.number_sites <= 24:
        op.basis.build()
        x = _x(op.basis.numberStates(), True, 3)
        y_on = op.matvec(x)
        op.set_option("canon", 0)
        assert op.info("canon_mode") == 0
        y_off = op.matvec(x)
        assert _close(y_on, y_off)
        y_ref = po.matvec_global(matrix, op.basis.representatives(), x, 1)
        assert _close(y_on, y_ref)
        op.set_option("canon", -1)
    # computeOffDiag goes through orbit_representative for every emitted state: compare the projected states
    bo = BatchedOperator(op, 512)
    n, betas, coeffs, keys = bo.computeOffDiag(512, po.state_info(basis, alphas[:512])[0], np.ones(512))
    ob, oc, ok, _ = po.compute_off_diag(matrix, 1, po.state_info(basis, alphas[:512])[0], np.ones(512))
    assert np.array_equal(np.sort(betas), np.sort(ob))
    assert np.array_equal(op.basis.stateInfo(alphas)[0], want)
    op.close()


def _dense_from_oracle(matrix, reps, cplx):
    n = reps.shape[0]
    H = np.zeros((n, n), dtype=np.complex128 if cplx else np.float64)
    for j in range(n):
        e = np.zeros(n, dtype=H.dtype)
        e[j] = 1.0
        H[:, j] = po.matvec_global(matrix, reps, e, 1)
    return H


@pytest.mark.parametrize("name,known", [("heisenberg_chain_4", -8.0), ("heisenberg_chain_6", -11.2111),
                                        ("heisenberg_chain_8", -14.6044), ("heisenberg_chain_10", -18.0618),
                                        ("heisenberg_square_4x4", None), ("heisenberg_kagome_12_symm", None),
                                        ("complex_hopping", None), ("momentum_sector", None)])
def test_lanczos_ground_state(need_cuda, name, known):
    """dmv_lanczos (device-resident three-term recurrence on top of the product) against dense diagonalisation of the
    oracle's matrix, and the Heisenberg-ring ground-state energies in sigma units (SURVEY.md section 8c pin 3)."""
    basis, matrix = GENERAL_MODELS[name]() if name in GENERAL_MODELS else _load(name)
    reps, _ = po.enumerate_states(basis)
    cplx = name in ("complex_hopping", "momentum_sector")
    H = _dense_from_oracle(matrix, reps, cplx)
    assert np.allclose(H, H.conj().T, atol=1e-12)
    w = np.linalg.eigvalsh(H)
    op = Operator(matrix)
    op.basis.build()
    e0, vec, iters, res = op.lanczos(max_iters=400, tol=1e-12, complex_vectors=cplx)
    assert abs(e0 - w[0]) <= 1e-9 * max(1.0, abs(w[0])), (e0, w[0], iters, res)
    if known is not None:
        assert abs(e0 - known) < 5e-4
    assert iters <= reps.shape[0] and abs(np.linalg.norm(vec) - 1.0) < 1e-8
    assert np.linalg.norm(H @ vec - e0 * vec) <= 1e-6 * max(1.0, abs(w[0])), res
    op.close()


@pytest.mark.parametrize("name,cplx", [("heisenberg_chain_10", False), ("heisenberg_kagome_12_symm", False),
                                       ("heisenberg_square_4x4", True), ("heisenberg_chain_12", False)])
def test_block_eigensolver_lobpcg(need_cuda, name, cplx):
    """The block eigensolver on dmv_matvec_batch (what PRIMME with blockSize > 1 is to the reference,
    src/Diagonalize.chpl:134-225): three lowest eigenvalues against dense diagonalisation of the oracle's matrix."""
    from distributed_matvec_b200 import lobpcg
    basis, matrix = _load(name)
    op = Operator(matrix)
    op.basis.build()
    reps = op.basis.representatives()
    H = _dense_from_oracle(matrix, reps, cplx)
    want = np.linalg.eigvalsh((H + H.conj().T) / 2)[:3]
    lam, X, iters, res = lobpcg(op, k=3, max_iters=400, tol=1e-9, complex_vectors=cplx)
    assert np.allclose(lam, want, rtol=0, atol=1e-8 * max(1.0, np.abs(want).max())), (lam, want, iters, res)
    # the vectors are eigenvectors: H x = lambda x through the single-vector product
    for j in range(3):
        y = op.matvec(X[j].contiguous())
        assert torch.linalg.norm(y - lam[j] * X[j]) <= 1e-6 * max(1.0, abs(lam[j]))
    op.close()


def test_lanczos_sector_spectrum_at_size(need_cuda):
    """The fully symmetric sector contains the ground state of the chain: the lowest eigenvalue of chain_24 (2.7 M
    states, k_gather) equals that of chain_24_symm (28 968 states, orbit scans) -- SURVEY.md section 8c pin 3."""
    energies = []
    for name in ("heisenberg_chain_24", "heisenberg_chain_24_symm"):
        basis, matrix = _load(name)
        op = Operator(matrix)
        op.basis.build()
        e0, _, iters, res = op.lanczos(max_iters=300, tol=1e-11, eigenvector=False)
        energies.append(e0)
        op.close()
    assert abs(energies[0] - energies[1]) <= 1e-8 * abs(energies[0]), energies
    assert abs(energies[0] / 24 - (-1.7738)) < 0.02      # approaches 4 (1/4 - ln 2) = -1.7726 per site from below


def test_bitparallel_matches_group_walk(need_cuda):
    basis, matrix = _load("heisenberg_kagome_16")
    op = Operator(matrix)
    op.basis.build()
    assert op.info("bp_words") == 1
    x = _x(op.basis.numberStates(), True)
    ys = []
    for mode in (0, 1):
        for bp in (1, 0):
            op.set_option("mode", mode)
            op.set_option("bitparallel", bp)
            ys.append(op.matvec(x))
    for y in ys[1:]:
        assert _close(y, ys[0])
    op.close()
