"""world_size-2 (and 3) CPU tests of the N > 1 host path: hash partition, plan exchange and the
all_to_all of (beta, coeff) records through torch.distributed (gloo), with the per-rank device work
replaced by a stand-in built on the oracle (TEST ONLY: the product path never runs on the CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleRank:
    """Stand-in for one rank's `Operator` with the methods HostExchangedProduct drives."""

    def __init__(self, matrix, blocks, rank, world):
        from oracle import pyoracle as po
        self.po, self.matrix, self.blocks, self.rank, self.num_ranks = po, matrix, blocks, rank, world
        self._out = None

    def plan(self):
        xs = np.ones(self.blocks[self.rank].shape[0])
        _, _, keys, _ = self.po.compute_off_diag(self.matrix, self.num_ranks, self.blocks[self.rank], xs)
        return np.bincount(keys, minlength=self.num_ranks).astype(np.int64)

    def record_width(self, x):
        # the library's rule (Operator.record_width / complex_values in csrc/dmv_api.cu): two doubles per coefficient for
        # complex vectors or when a coefficient / character has an imaginary part; real -1 characters stay width 1
        b = self.matrix.basis
        cplx = bool(np.any(self.matrix.off_diag.v.imag != 0) or np.any(self.matrix.diag.v.imag != 0))
        if b.has_permutation_symmetries():
            cplx |= bool(np.any(b.group.characters.imag != 0))
        return 2 if (x.is_complex() or cplx) else 1

    def generate(self, x, y):
        po, mine = self.po, self.blocks[self.rank]
        xs = x.numpy()
        if len(self.matrix.diag):
            y.copy_(torch.from_numpy(np.asarray(po.apply_diag(self.matrix, mine, xs))))
        betas, coeffs, keys, _ = po.compute_off_diag(self.matrix, self.num_ranks, mine, xs)
        order = np.argsort(keys, kind="stable")
        betas, coeffs, keys = betas[order], coeffs[order], keys[order]
        own = keys == self.rank
        self._accumulate(betas[own], coeffs[own], y)
        self._out = (betas[~own], coeffs[~own])

    def outgoing_tensors(self, width):
        b, c = self._out
        flat = np.ascontiguousarray(c).view(np.float64).copy() if width == 2 else np.ascontiguousarray(c.real)
        return torch.from_numpy(b.view(np.int64).copy()), torch.from_numpy(flat)

    def _accumulate(self, betas, coeffs, y):
        idx = self.po.state_index(self.blocks[self.rank], betas)
        assert np.all(idx >= 0)
        yn = y.numpy()
        np.add.at(yn, idx, coeffs if np.iscomplexobj(yn) else coeffs.real)

    def accumulate_tensors(self, x, betas, coeffs, y):
        c = coeffs.numpy()
        c = c.view(np.complex128) if self.record_width(x) == 2 else c.astype(np.complex128)
        self._accumulate(betas.numpy().view(np.uint64), c, y)


def _worker(rank, world, port, name, cplx, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from distributed_matvec_b200 import load_config_from_yaml
        from distributed_matvec_b200.distributed import HostExchangedProduct, block_to_hashed
        from oracle import pyoracle as po
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        reps, _ = po.enumerate_states(basis)
        masks, blocks = po.partition_by_hash(reps, world)
        rng = np.random.default_rng(42)
        x = rng.random(reps.shape[0]) - 0.5
        if cplx:
            x = x + 1j * (rng.random(reps.shape[0]) - 0.5)
        y_ref = po.matvec_global(matrix, reps, x, 1)
        mine = torch.from_numpy(block_to_hashed(x, masks, world)[rank])
        prod = HostExchangedProduct(OracleRank(matrix, blocks, rank, world))
        y = prod.matvec(mine, torch.zeros_like(mine))
        err = float(np.abs(y.numpy() - y_ref[masks == rank]).max() / np.abs(y_ref).max())
        # every record this rank received was announced by the plan exchange
        ok = err < 1e-12 and sum(prod.recv_counts) >= 0 and prod.send_counts[rank] == 0
        queue.put((rank, ok, err))
    finally:
        dist.destroy_process_group()


class OracleReplicatedRank:
    """Stand-in for one rank's `Operator` with the methods HostReplicatedProduct drives: the rows of this rank from the
    gathered x, computed by the oracle on the whole basis."""

    def __init__(self, matrix, reps, masks, rank, world):
        from oracle import pyoracle as po
        self.po, self.matrix, self.reps, self.masks, self.rank, self.world = po, matrix, reps, masks, rank, world
        self.block = 0

    def replicated_setup(self):
        largest = max(int((self.masks == r).sum()) for r in range(self.world))
        self.block = (largest + 1) // 2 * 2            # same rule as dmv_replicated_setup
        return self.block

    def replicated_rows(self, x_cat, y):
        xc = x_cat.numpy()
        assert xc.shape[0] == self.block * self.world
        x = np.zeros(self.reps.shape[0], dtype=xc.dtype)
        for r in range(self.world):                    # slot r holds rank r's block, zero padded
            n_r = int((self.masks == r).sum())
            x[self.masks == r] = xc[r * self.block:r * self.block + n_r]
            assert not np.any(xc[r * self.block + n_r:(r + 1) * self.block])
        y_all = self.po.matvec_global(self.matrix, self.reps, x, 1)
        y.copy_(torch.from_numpy(np.ascontiguousarray(y_all[self.masks == self.rank])))
        return y


def _worker_replicated(rank, world, port, name, cplx, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from distributed_matvec_b200 import load_config_from_yaml
        from distributed_matvec_b200.distributed import HostReplicatedProduct, block_to_hashed
        from oracle import pyoracle as po
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        reps, _ = po.enumerate_states(basis)
        masks, blocks = po.partition_by_hash(reps, world)
        rng = np.random.default_rng(42)
        x = rng.random(reps.shape[0]) - 0.5
        if cplx:
            x = x + 1j * (rng.random(reps.shape[0]) - 0.5)
        y_ref = po.matvec_global(matrix, reps, x, world)
        mine = torch.from_numpy(block_to_hashed(x, masks, world)[rank])
        prod = HostReplicatedProduct(OracleReplicatedRank(matrix, reps, masks, rank, world))
        y = prod.matvec(mine, torch.zeros_like(mine))
        err = float(np.abs(y.numpy() - y_ref[masks == rank]).max() / np.abs(y_ref).max())
        queue.put((rank, err < 1e-12 and prod.block >= mine.shape[0], err))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name,cplx", [("heisenberg_chain_10", False), ("heisenberg_kagome_12_symm", True),
                                       ("heisenberg_chain_12", False)])
def test_host_exchanged_product_gloo(name, cplx, world):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, cplx, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, err in results:
        assert ok, (rank, err)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name,cplx", [("heisenberg_chain_10", True), ("heisenberg_kagome_12_symm", False)])
def test_host_replicated_product_gloo(name, cplx, world):
    """The replicated-x form with the all-gather owned by the host: equal slots (largest block, zero padded), every rank
    computes its own rows -- the N > 1 host logic on CPU (gloo)."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_replicated, args=(r, world, port, name, cplx, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, err in results:
        assert ok, (rank, err)
