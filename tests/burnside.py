"""Dimension of a symmetry-adapted spin basis by Burnside's lemma -- counting, no enumeration.  Test infrastructure.

For trivial characters every orbit of G on the states of fixed Hamming weight w carries one basis state, so
N = 1/|G| sum_g fix(g).  g = (permutation pi, optional global flip):
  * without flip, a state is fixed iff it is constant on the cycles of pi: fix = number of sub-multisets of the cycle
    lengths that sum to w (a subset-sum count);
  * with flip, s_i = not s_pi(i): every cycle must have even length and alternates (2 choices each), which puts exactly
    half of the sites up: fix = 2^cycles if all cycles are even and w = n / 2, else 0.
"""
import numpy as np


def cycle_lengths(perm) -> list:
    perm = [int(p) for p in perm]
    seen, out = [False] * len(perm), []
    for i in range(len(perm)):
        if not seen[i]:
            n, j = 0, i
            while not seen[j]:
                seen[j] = True
                j = perm[j]
                n += 1
            out.append(n)
    return out


def fixed_states(perm, flip: bool, weight: int) -> int:
    cycles = cycle_lengths(perm)
    n = len(perm)
    if flip:
        return (1 << len(cycles)) if all(c % 2 == 0 for c in cycles) and 2 * weight == n else 0
    ways = [0] * (weight + 1)        # ways[k]: selections of cycles with total length k
    ways[0] = 1
    for c in cycles:
        for k in range(weight, c - 1, -1):
            ways[k] += ways[k - c]
    return ways[weight]


def dimension(perms, flips, weight: int) -> int:
    total = sum(fixed_states(p, bool(f), weight) for p, f in zip(np.asarray(perms), np.asarray(flips)))
    assert total % len(perms) == 0
    return total // len(perms)
