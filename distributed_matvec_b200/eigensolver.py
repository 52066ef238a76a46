"""Block eigensolver on the batched product -- the consumer of the hot path ("next" row f3).

The reference hands its matrix-vector product to PRIMME (reference src/Diagonalize.chpl:134-225: `ls_chpl_primme_matvec`
as `matrixMatvec`, block size `kMaxBlockSize`, global sums through `primmeGlobalSumReal`, src/PRIMME.chpl:267-322).  PRIMME is
third-party and not in the image; ``ls_chpl_primme_matvec`` itself is exported by libdmv_b200 (see ChapelKernels.primme_matvec),
and this module is the stand-in for the solver above it: LOBPCG (Knyazev 2001) for the k lowest eigenpairs, every application of
H being ONE call of ``Operator.matvec_batch`` on device-resident blocks (several vectors share one walk over the terms in the
kernels).  The small Rayleigh-Ritz problems (3k x 3k) are solved with torch.linalg on the device; with several ranks the Gram
matrices are summed with torch.distributed (the analogue of primmeGlobalSumReal).
"""
from __future__ import annotations

import numpy as np


def _gram(a, b, dist_group):
    import torch
    g = a.conj() @ b.transpose(0, 1)          # rows are vectors: (ka, n) x (n, kb)
    if dist_group is not None:
        import torch.distributed as dist
        if g.is_complex():
            r = torch.view_as_real(g).contiguous()
            dist.all_reduce(r, group=dist_group if dist_group is not True else None)
            g = torch.view_as_complex(r)
        else:
            dist.all_reduce(g, group=dist_group if dist_group is not True else None)
    return g


def _orthonormalise(s, dist_group, drop=1e-10):
    """Rows of s -> an orthonormal set spanning the same space (eigen-decomposition of the Gram matrix; directions with
    relative weight below `drop` are removed)."""
    import torch
    m = _gram(s, s, dist_group)
    w, v = torch.linalg.eigh((m + m.conj().transpose(0, 1)) / 2)
    keep = w > drop * w.max().clamp_min(1e-300)
    t = (v[:, keep] / torch.sqrt(w[keep])).to(s.dtype)       # t_a = sum_j t[j, a] s_j
    return t.transpose(0, 1) @ s


def _project(x, v, dist_group):
    """Component of the rows of v inside span(rows of x) (x orthonormal): sum_i <x_i, v_j> x_i."""
    return _gram(x, v, dist_group).transpose(0, 1) @ x


def lobpcg(op, k: int = 1, max_iters: int = 200, tol: float = 1e-8, complex_vectors: bool = False, seed: int = 42,
           distributed: bool = False):
    """k lowest eigenpairs of the operator on this rank's block (collective when `distributed`): returns
    (eigenvalues [k], eigenvectors torch (k, n_local) on the device, iterations, residual norms [k]).

    `op` is an Operator or DistributedOperator (anything with matvec_batch and basis.numberStates()); every iteration
    applies H once to a block of up to 2k new directions through dmv_matvec_batch."""
    import torch
    base = op.op if hasattr(op, "op") else op
    matvec = base.matvec_batch
    n = base.basis.numberStates()
    dev = torch.device("cuda", base.device) if getattr(base, "device", None) is not None else torch.device("cpu")
    dtype = torch.complex128 if complex_vectors else torch.float64
    group = True if distributed else None
    gen = torch.Generator(device="cpu").manual_seed(seed + 7919 * getattr(base, "rank", 0))
    x = torch.randn((k, n), generator=gen, dtype=torch.float64)
    if complex_vectors:
        x = x + 1j * torch.randn((k, n), generator=gen, dtype=torch.float64)
    x = _orthonormalise(x.to(dev).to(dtype), group)
    ax = matvec(x.contiguous())
    p = None
    lam = res = None
    for it in range(1, max_iters + 1):
        # Rayleigh-Ritz on the current block, residuals
        h = _gram(x, ax, group)
        lam, c = torch.linalg.eigh((h + h.conj().transpose(0, 1)) / 2)
        ct = c.to(dtype).transpose(0, 1)
        x, ax = ct @ x, ct @ ax
        r = ax - lam.to(dtype)[:, None] * x
        res = torch.sqrt(torch.diagonal(_gram(r, r, group)).real)
        if bool(torch.all(res <= tol * torch.clamp(lam.abs(), min=1.0))):
            break
        # search directions: residuals and the previous step, orthonormalised against X
        d = r if p is None else torch.cat([r, p], dim=0)
        d = d - _project(x, d, group)
        d = _orthonormalise(d, group)
        d = d - _project(x, d, group)             # second pass: keeps [X, D] orthonormal to rounding
        d = _orthonormalise(d, group)
        ad = matvec(d.contiguous())
        s = torch.cat([x, d], dim=0)
        a_s = torch.cat([ax, ad], dim=0)
        g = _gram(s, a_s, group)
        _, vecs = torch.linalg.eigh((g + g.conj().transpose(0, 1)) / 2)
        q = vecs[:, :k].to(dtype).transpose(0, 1)    # (k, ks): new x_i = sum_j q[i, j] s_j
        p = q[:, k:] @ d                             # the part of the step outside the old X
        x, ax = q @ s, q @ a_s
    return lam.cpu().numpy(), x, it, res.cpu().numpy()
