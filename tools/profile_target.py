#!/usr/bin/env python3
"""Small target for ncu: build a basis, run a few products.  python tools/profile_target.py WORKLOAD MODE INDEX DTYPE [ITERS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_matvec_b200 import Operator, load_config_from_yaml  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "heisenberg_chain_24"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else -1
index = int(sys.argv[3]) if len(sys.argv) > 3 else -1
cplx = (sys.argv[4] if len(sys.argv) > 4 else "c128") == "c128"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
basis, matrix = load_config_from_yaml(os.path.join(ROOT, "data", name + ".yaml"))
op = Operator(matrix)
op.set_option("mode", mode)
op.set_option("index", index)
op.basis.build()
n = op.basis.numberStates()
rng = np.random.default_rng(42)
x = rng.random(n) - 0.5
if cplx:
    x = x + 1j * (rng.random(n) - 0.5)
xd = torch.from_numpy(x).cuda()
yd = torch.zeros_like(xd)
for _ in range(iters):
    op.matvec(xd, yd)
torch.cuda.synchronize()
op.synchronize()
print("done", name, n, "pull" if op.info("pull") else "push", "index_mode", op.info("index_mode"))
