"""Bit permutations as Benes networks, for the TIMED CPU arm only -- TEST INFRASTRUCTURE.

The checker (oracle_state_info in oracle.c) permutes bit by bit on purpose: it shares nothing with the GPU's networks.
Timing that loop as "the reference on CPU" would flatter the GPU, because the library behind the reference's
``ls_hs_state_info`` (reference src/FFI.chpl:181-184; lattice-symmetries) applies every group element as a Benes network
of 2 log2(64) - 1 = 11 delta-swap stages on whole 64-bit words.  ``benes(perm)`` builds those stages with the textbook
looping algorithm; oracle.c applies them (oracle_state_info_networks) in the timed product and in the parallel
enumeration.  tests/test_oracle_pins.py holds the networks against the bit-by-bit permutation.
"""
from __future__ import annotations

import numpy as np

WIDTH = 64
DELTAS = [32, 16, 8, 4, 2, 1, 2, 4, 8, 16, 32]


def _route(src, width):
    """src[i] = input position that must reach output i.  -> list of (mask, delta), application order."""
    if width == 1:
        return []
    if width == 2:
        return [(1 if src[0] == 1 else 0, 1)]
    d = width // 2
    dst = [0] * width
    for i, j in enumerate(src):
        dst[j] = i
    colour = [-1] * width          # 0: through the lower sub-network, 1: through the upper one
    for j0 in range(width):
        if colour[j0] != -1:
            continue
        j = j0
        while True:
            colour[j] = 0
            jp = j ^ d              # its input partner takes the other sub-network ...
            colour[jp] = 1
            jn = src[dst[jp] ^ d]   # ... so the output partner of where jp arrives is fed from the lower one
            if colour[jn] != -1:
                break
            j = jn
    mask_in = sum(1 << p for p in range(d) if colour[p] == 1)
    mask_out = sum(1 << i for i in range(d) if colour[src[i]] == 1)
    sub = [[0] * d, [0] * d]
    for i in range(width):
        c = colour[src[i]]
        sub[c][i % d] = src[i] % d
    lower, upper = _route(sub[0], d), _route(sub[1], d)
    inner = [(ml | (mu << d), dl) for (ml, dl), (mu, _) in zip(lower, upper)]
    return [(mask_in, d)] + inner + [(mask_out, d)]


def benes(perm) -> list[int]:
    """Masks of the 11 stages (deltas = DELTAS) that map s to g.s with (g.s)[i] = s[perm[i]]."""
    n = len(perm)
    src = list(int(p) for p in perm) + list(range(n, WIDTH))
    stages = _route(src, WIDTH)
    assert [d for _, d in stages] == DELTAS
    return [m for m, _ in stages]


def group_networks(group) -> np.ndarray:
    """uint64 [G, 11] stage masks of every element's permutation (the flip is applied separately)."""
    return np.array([benes(p) for p in group.perms], dtype=np.uint64).reshape(len(group), len(DELTAS))


def apply_network(masks, s: int) -> int:
    for m, d in zip(masks, DELTAS):
        m = int(m)
        t = ((s >> d) ^ s) & m
        s ^= t ^ (t << d)
    return s
