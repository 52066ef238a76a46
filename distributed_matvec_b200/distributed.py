"""Hash-partitioned (multi-GPU) form of the product: ``matrixVectorProduct`` (DMV:1072-1093).

Ways to run P ranks:

* ``DistributedOperator``: one process per GPU under torch.distributed (the bench / production shape).
  The exchange -- all-gather of x (replicated-x form) or the all-to-all of (beta, coeff) records -- is done inside
  libdmv_b200 (``dmv_matvec``, NCCL / NVLink peer stores); torch is only used to share the NCCL unique id and for
  barriers.
* ``HostExchangedProduct`` / ``HostReplicatedProduct``: the same two forms with the collective owned by the host
  through torch.distributed (any backend; covered on CPU with gloo).
* ``EmulatedCluster``: P logical ranks (P contexts) on ONE GPU, the analogue of the reference's
  GASNet-smp oversubscription (reference env/setup-env.sh:5,13).  The exchange is replaced by handing
  each destination context the sender's outgoing bucket (same device, so the pointers are valid).
  Used by the single-GPU tests to cover the bucketing / accumulate path for P in {2, 3, 4, ...}.

The block <-> hashed conversions of vectors (arrFromBlockToHashed / arrFromHashedToBlock,
src/BlockToHashed.chpl:87, src/HashedToBlock.chpl:67) run on the GPU through ``Operator.block_to_hashed`` /
``Operator.hashed_to_block`` (dmv_block_to_hashed / dmv_hashed_to_block); the numpy helpers below state the same
permutation for the tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from .config import OperatorSpec
from .operator import Operator, _elt_of, _ptr, locale_idx_of


def masks_of(op: Operator, states: np.ndarray, num_ranks: int) -> np.ndarray:
    """masks[i] = owner of the i-th state in global sorted order (SE:138-156)."""
    if num_ranks == 1:
        return np.zeros(states.shape[0], dtype=np.uint8)
    return locale_idx_of(op, states, num_ranks)


def block_to_hashed(arr: np.ndarray, masks: np.ndarray, num_ranks: int) -> list[np.ndarray]:
    return [np.ascontiguousarray(arr[masks == p]) for p in range(num_ranks)]


def hashed_to_block(blocks: list[np.ndarray], masks: np.ndarray) -> np.ndarray:
    out = np.zeros(masks.shape[0], dtype=blocks[0].dtype)
    for p, blk in enumerate(blocks):
        out[masks == p] = blk
    return out


class EmulatedCluster:
    """P logical ranks on one GPU."""

    def __init__(self, spec: OperatorSpec, num_ranks: int, device: int = 0):
        self.num_ranks = num_ranks
        self.ops = [Operator(spec, device=device, rank=r, num_ranks=num_ranks) for r in range(num_ranks)]

    def build(self):
        for op in self.ops:
            op.basis.build()
        return self

    def set_representatives(self, blocks, norms=None):
        for r, op in enumerate(self.ops):
            op.basis.uncheckedSetRepresentatives(blocks[r], None if norms is None else norms[r])
        return self

    def representatives(self) -> list[np.ndarray]:
        return [op.basis.representatives() for op in self.ops]

    def matvec(self, x_blocks):
        """x_blocks / result: list of torch CUDA tensors, one per logical rank."""
        import torch
        ys = [torch.zeros_like(x) for x in x_blocks]
        elt = _elt_of(x_blocks[0])
        for op in self.ops:
            op.plan()
        # generate everywhere (each rank also accumulates the records it owns itself) ...
        for r, op in enumerate(self.ops):
            op.generate(x_blocks[r], ys[r])
        # ... then the "exchange": destination q consumes the bucket rank r made for it
        for r, op in enumerate(self.ops):
            op.synchronize()
        for r, src in enumerate(self.ops):
            for q, dst in enumerate(self.ops):
                if q == r:
                    continue
                betas, coeffs, n = src.outgoing(q)
                if n > 0:
                    dst.accumulate(elt, n, betas, coeffs, ys[q])
        for op in self.ops:
            op.synchronize()
        return ys

    def matvec_replicated(self, x_blocks):
        """The replicated-x form (dmv_replicated_*): the host plays the all-gather -- every logical rank gets the
        same gathered x (rank r's block in slot r) and computes its own rows."""
        import torch
        elt = _elt_of(x_blocks[0])
        for op in self.ops:
            nat.check(nat.lib().dmv_replicated_setup(op._ctx))
        block = self.ops[0].info("replicated_block")
        assert all(op.info("replicated_block") == block for op in self.ops)
        x_cat = torch.zeros(block * self.num_ranks, dtype=x_blocks[0].dtype, device=x_blocks[0].device)
        for r, x in enumerate(x_blocks):
            x_cat[r * block:r * block + x.shape[0]] = x
        ys = [torch.zeros_like(x) for x in x_blocks]
        torch.cuda.synchronize()
        for r, op in enumerate(self.ops):
            nat.check(nat.lib().dmv_replicated_product(op._ctx, elt, _ptr(x_cat), _ptr(ys[r])))
        for op in self.ops:
            op.synchronize()
        return ys

    def close(self):
        for op in self.ops:
            op.close()


class DistributedOperator:
    """One rank of the real multi-GPU product.  Requires torch.distributed to be initialised."""

    def __init__(self, spec: OperatorSpec, device: int | None = None):
        import torch
        import torch.distributed as dist
        self.rank, self.num_ranks = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = torch.cuda.current_device()
        self.op = Operator(spec, device=device, rank=self.rank, num_ranks=self.num_ranks)
        if self.num_ranks > 1:
            ident = [None]
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                nat.check(nat.lib().dmv_comm_unique_id(buf))
                ident[0] = bytes(buf)
            dist.broadcast_object_list(ident, src=0)
            buf = (C.c_ubyte * 128).from_buffer_copy(ident[0])
            nat.check(nat.lib().dmv_comm_init(self.op._ctx, buf))
        self.basis = self.op.basis

    def matvec(self, x, y=None):
        return self.op.matvec(x, y)


def matrix_vector_product(matrix, x, y=None):
    """``matrixVectorProduct(matrix, x, y, representatives)`` (DMV:1072-1075) on this rank's blocks;
    `matrix` is an Operator (one rank) or a DistributedOperator."""
    return matrix.matvec(x, y)


class HostExchangedProduct:
    """``matrixVectorProduct`` with the exchange owned by the HOST through torch.distributed
    (``all_to_all_single``), driving the stepwise C ABI (dmv_plan / dmv_generate / dmv_outgoing /
    dmv_accumulate) -- what a Chapel host doing its own PUTs (DMV:361-371) would do.  Works with any
    torch.distributed backend; `rank_ops` is this rank's Operator (or, in the CPU tests, a stand-in with
    the same methods built on the oracle, so that the exchange logic is covered without a GPU).

    Protocol per product (all ranks):
      1. counts[r][q] (records r sends q) are fixed by the plan; exchanged once with all_gather
      2. generate: own records are accumulated locally, the rest land in per-destination buckets
      3. all_to_all_single of betas (int64 view of uint64) and coefficients (float64, `width` per record)
      4. accumulate the received records
    """

    def __init__(self, rank_ops, width_of=None):
        import torch
        import torch.distributed as dist
        self.ops = rank_ops
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        send = torch.from_numpy(np.asarray(self.ops.plan(), dtype=np.int64).copy())
        send[self.rank] = 0
        if dist.get_backend() == "nccl":      # NCCL moves device tensors only
            send = send.cuda(getattr(self.ops, "device", None))
        gathered = [torch.zeros_like(send) for _ in range(self.world)]
        dist.all_gather(gathered, send)
        self.send_counts = send.tolist()
        self.recv_counts = [int(gathered[q][self.rank]) for q in range(self.world)]

    def matvec(self, x, y):
        import torch
        import torch.distributed as dist
        width = self.ops.record_width(x)
        self.ops.generate(x, y)
        betas_out, coeffs_out = self.ops.outgoing_tensors(width)   # concatenated over destinations
        betas_in = torch.empty(sum(self.recv_counts), dtype=torch.int64, device=betas_out.device)
        coeffs_in = torch.empty(sum(self.recv_counts) * width, dtype=torch.float64, device=betas_out.device)
        dist.all_to_all_single(betas_in, betas_out, self.recv_counts, self.send_counts)
        dist.all_to_all_single(coeffs_in, coeffs_out, [c * width for c in self.recv_counts],
                               [c * width for c in self.send_counts])
        self.ops.accumulate_tensors(x, betas_in, coeffs_in, y)
        return y


class HostReplicatedProduct:
    """The replicated-x form of ``matrixVectorProduct`` with the all-gather owned by the HOST through torch.distributed
    (any backend): every rank contributes its block of x to a gathered vector with one equally sized slot per rank
    (slot size = the largest block, as dmv_replicated_setup reports it) and computes its own rows.  `rank_ops` is this
    rank's Operator (methods replicated_setup / replicated_rows), or in the CPU tests a stand-in with the same
    methods built on the oracle, so that the slot logic is covered without a GPU."""

    def __init__(self, rank_ops):
        import torch.distributed as dist
        self.ops = rank_ops
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.block = int(self.ops.replicated_setup())
        import torch
        dev = torch.device("cuda", getattr(self.ops, "device", 0)) if dist.get_backend() == "nccl" else torch.device("cpu")
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(sizes, torch.tensor([self.block], dtype=torch.int64, device=dev))
        if any(int(b) != self.block for b in sizes):      # every rank derives it from the same global basis
            raise RuntimeError("ranks disagree on the slot size of the gathered x")

    def matvec(self, x, y):
        import torch
        import torch.distributed as dist
        if x.shape[0] > self.block:
            raise ValueError("block larger than its slot")
        slot = torch.zeros(self.block, dtype=x.dtype, device=x.device)
        slot[:x.shape[0]] = x
        slots = [torch.empty_like(slot) for _ in range(self.world)]
        dist.all_gather(slots, slot)
        return self.ops.replicated_rows(torch.cat(slots), y)
