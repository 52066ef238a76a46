#!/usr/bin/env python3
"""Parity of the real multi-GPU product (NCCL all-to-all inside libdmv_b200) against the CPU oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py [workload ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from distributed_matvec_b200 import DistributedOperator, load_config_from_yaml  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = sys.argv[1:] or ["heisenberg_chain_16", "heisenberg_kagome_16", "heisenberg_chain_10",
                             "heisenberg_square_4x4", "heisenberg_chain_24_symm", "heisenberg_chain_20"]
    failures = 0
    for name in names:
        basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
        dop = DistributedOperator(matrix, device=local)
        dop.op.set_option("exchange", int(os.environ.get("DMV_EXCHANGE", "-1")))
        dop.basis.build()
        mine = dop.basis.representatives()
        o_reps, _ = po.enumerate_states(basis)
        masks, blocks = po.partition_by_hash(o_reps, world)
        ok_basis = np.array_equal(mine, blocks[rank])
        # block <-> hashed redistribution (collective, NCCL all-to-all-v inside the library)
        bounds = np.linspace(0, o_reps.shape[0], world + 1).astype(int)
        m_chunk = masks[bounds[rank]:bounds[rank + 1]]
        for arr in (o_reps, (np.arange(o_reps.shape[0]) * (1 + 2j)).astype(np.complex128)):
            hashed = dop.op.block_to_hashed(arr[bounds[rank]:bounds[rank + 1]], m_chunk)
            ok_basis &= bool(np.array_equal(hashed, arr[masks == rank]))
            back = dop.op.hashed_to_block(torch.from_numpy(hashed).cuda(), m_chunk)
            ok_basis &= bool(np.array_equal(back.cpu().numpy(), arr[bounds[rank]:bounds[rank + 1]]))
        for cplx in (False, True):
            rng = np.random.default_rng(42)
            x = rng.random(o_reps.shape[0]) - 0.5
            if cplx:
                x = x + 1j * (rng.random(o_reps.shape[0]) - 0.5)
            y_ref = po.matvec_global(matrix, o_reps, x, world)[masks == rank]
            x_mine = np.ascontiguousarray(x[masks == rank])
            # host vectors through the C ABI (collective call)
            y_host = dop.matvec(x_mine)
            # device-resident vectors
            xd = torch.from_numpy(x_mine).cuda()
            yd = dop.matvec(xd)
            torch.cuda.synchronize()
            dop.op.synchronize()
            scale = max(np.abs(y_ref).max(), 1e-300)
            e1 = np.abs(y_host - y_ref).max() / scale
            e2 = np.abs(yd.cpu().numpy() - y_ref).max() / scale
            good = ok_basis and e1 < 1e-12 and e2 < 1e-12
            flag = torch.tensor([0 if good else 1], device="cuda")
            dist.all_reduce(flag)
            if rank == 0:
                print(f"{name:28s} P={world} {'c128' if cplx else 'f64 '} N={o_reps.shape[0]} basis_ok={ok_basis} "
                      f"err_host={e1:.1e} err_dev={e2:.1e} exchange={'replicated-x' if dop.op.info('replicated') else ('peer-direct' if dop.op.info('peer_direct') else 'nccl')} "
                      f"{'OK' if int(flag) == 0 else 'FAIL'}", flush=True)
            failures += int(flag)
        # Lanczos across the ranks (dot products reduced with NCCL) against the single-rank oracle matrix
        if o_reps.shape[0] <= 13000:
            e0, _, iters, res = dop.op.lanczos(max_iters=200, tol=1e-11, eigenvector=False)
            e_all = torch.tensor([e0], device="cuda", dtype=torch.float64)
            lst = [torch.zeros_like(e_all) for _ in range(world)]
            dist.all_gather(lst, e_all)
            same = all(abs(float(t) - e0) <= 1e-12 * abs(e0) for t in lst)
            if rank == 0:
                print(f"{name:28s} P={world} lanczos E0={e0:.10f} iters={iters} residual={res:.1e} ranks agree={same} "
                      f"{'OK' if same else 'FAIL'}", flush=True)
            failures += 0 if same else 1
        dop.op.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
