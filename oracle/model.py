"""Model inputs for the oracle, read WITHOUT the product's host layer -- TEST INFRASTRUCTURE ONLY.

``load_model(path)`` turns one of data/*.yaml into the flat tables oracle/oracle.c consumes.  It is a second,
independent reading of the same files (the product's is distributed_matvec_b200/config.py, expr.py, symmetry.py), so

  * ``bench.py --impl reference`` runs with nothing of the product imported or mapped, and
  * tests/test_oracle_pins.py can hold the two readings against each other.

What the reference does here lives in the third-party ``ls_hs_load_yaml_config`` (reference src/ForeignTypes.chpl:261-283;
semantics restated in SURVEY.md App. A.1-A.3).  Scope: the expression grammar of the reference's data/*.yaml --
``[number x] P_0 P_1 ...`` with P in {sigma^x, sigma^y, sigma^z, S^x, S^y, S^z, sigma^+, sigma^-} -- and ``matrix:`` terms
(data/old/*.yaml).  Conventions as in DESIGN.md section 1: site i = bit i, bit 1 = spin up, a permutation p acts as
(g s)[i] = s[p[i]].
"""
from __future__ import annotations

import cmath
import re
from fractions import Fraction
from math import comb

import numpy as np
import yaml

_SUP = {"ˣ": "x", "ʸ": "y", "ᶻ": "z", "⁺": "+", "⁻": "-"}
_SUB = {c: str(i) for i, c in enumerate("₀₁₂₃₄₅₆₇₈₉")}
# index = bit value (1 = up); rows = outgoing bit, columns = incoming bit
_ONE_SITE = {
    "x": np.array([[0, 1], [1, 0]], dtype=complex),
    "y": np.array([[0, 1j], [-1j, 0]], dtype=complex),     # up <- down element is -i
    "z": np.array([[-1, 0], [0, 1]], dtype=complex),
    "+": np.array([[0, 0], [1, 0]], dtype=complex),
    "-": np.array([[0, 1], [0, 0]], dtype=complex),
}
_FACTOR = re.compile(r"\s*([σS])([ˣʸᶻ⁺⁻])([₀₁₂₃₄₅₆₇₈₉]+)")
_NUMBER = re.compile(r"\s*([0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)\s*[×*]")


def expression_matrix(text: str):
    """(k, M): the 2^k x 2^k matrix of a product expression; placeholder site 0 is the most significant index bit."""
    pos, coeff = 0, 1.0 + 0j
    m = _NUMBER.match(text, pos)
    if m:
        coeff, pos = complex(float(m.group(1))), m.end()
    factors = []
    while True:
        m = _FACTOR.match(text, pos)
        if not m:
            break
        kind, comp, sub = m.group(1), _SUP[m.group(2)], int("".join(_SUB[c] for c in m.group(3)))
        mat = _ONE_SITE[comp] * (0.5 if kind == "S" and comp in "xyz" else 1.0)
        factors.append((sub, mat))
        pos = m.end()
    if text[pos:].strip() or not factors:
        raise ValueError(f"oracle/model.py cannot read the expression {text!r}")
    k = 1 + max(s for s, _ in factors)
    per_site = [np.eye(2, dtype=complex) for _ in range(k)]
    for s, mat in factors:
        per_site[s] = per_site[s] @ mat
    total = np.array([[coeff]])
    for s in range(k):
        total = np.kron(total, per_site[s])
    return k, total


class Terms:
    """Flat non-branching term table (v, m, r, x, s): <b|t|a> = v [a & m == r] (-1)^popcount(a & s), b = a ^ x."""

    def __init__(self, rows):
        self.v = np.array([t[0] for t in rows], dtype=np.complex128)
        self.m, self.r, self.x, self.s = (np.array([t[i] for t in rows], dtype=np.uint64) for i in (1, 2, 3, 4))

    def __len__(self):
        return int(self.v.shape[0])


class Group:
    def __init__(self, perms, flips, chars, n):
        self.perms = np.array(perms, dtype=np.int32).reshape(len(perms), n)
        self.flips = np.array(flips, dtype=np.uint8)
        self.characters = np.array(chars, dtype=np.complex128)

    def __len__(self):
        return int(self.perms.shape[0])

    @property
    def all_characters_trivial(self):
        return bool(np.all(self.characters == 1))


class Basis:
    """The basis flags the reference reads off ``ls_hs_basis`` (src/ForeignTypes.chpl:82-109)."""

    def __init__(self, d):
        self.number_sites = int(d["number_spins"])
        hw, inv = d.get("hamming_weight"), d.get("spin_inversion")
        self.hamming_weight = None if hw is None else int(hw)
        self.spin_inversion = 0 if inv is None else int(inv)
        self.generators = [([int(v) for v in g["permutation"]], int(g.get("sector", 0))) for g in (d.get("symmetries") or [])]
        self._group = None

    def is_hamming_weight_fixed(self): return self.hamming_weight is not None
    def has_spin_inversion_symmetry(self): return self.spin_inversion != 0
    def has_permutation_symmetries(self): return len(self.generators) > 0
    def requires_projection(self): return self.has_permutation_symmetries() or self.has_spin_inversion_symmetry()
    def is_state_index_identity(self): return not self.requires_projection() and not self.is_hamming_weight_fixed()

    def min_state_estimate(self):
        return 0 if self.hamming_weight is None else (1 << self.hamming_weight) - 1

    def max_state_estimate(self):
        """Largest candidate; with spin inversion the top site of a representative is never set."""
        n, w = self.number_sites, self.hamming_weight
        top_clear = 1 if self.has_spin_inversion_symmetry() else 0
        if w is None:
            return ((1 << n) - 1) >> top_clear
        if top_clear and (n == 0 or w == 0):
            top_clear = 0
        width = n - top_clear
        return ((1 << w) - 1) << (width - w) if width >= w else 0

    def number_candidates(self):
        return comb(self.number_sites, self.hamming_weight) if self.hamming_weight is not None else 1 << self.number_sites

    @property
    def group(self):
        """Closure of the generators; the character of a generator of period T in sector k is exp(-2 pi i k / T)
        (SURVEY App. A.3); spin inversion doubles the group with character ``spin_inversion``."""
        if self._group is None:
            n = self.number_sites
            ident = tuple(range(n))

            def after(p, q):        # apply q first, then p
                return tuple(q[i] for i in p)

            gens = []
            for p, sector in self.generators:
                p = tuple(p)
                if sorted(p) != list(range(n)):
                    raise ValueError("not a permutation")
                cur, period = p, 1
                while cur != ident:
                    cur, period = after(p, cur), period + 1
                gens.append((p, Fraction(sector, period) % 1))
            seen = {ident: Fraction(0)}
            frontier = [ident]
            while frontier:
                nxt = []
                for e in frontier:
                    for p, ph in gens:
                        c, phase = after(p, e), (seen[e] + ph) % 1
                        if c in seen:
                            if seen[c] != phase:
                                raise ValueError("sectors do not define a one-dimensional representation")
                        else:
                            seen[c] = phase
                            nxt.append(c)
                frontier = nxt
            perms, flips, chars = [], [], []
            for p, phase in sorted(seen.items()):
                chi = 1 + 0j if phase == 0 else (-1 + 0j if phase == Fraction(1, 2) else cmath.exp(-2j * cmath.pi * float(phase)))
                perms.append(p), flips.append(0), chars.append(chi)
            if self.spin_inversion:
                for p, chi in list(zip(perms, chars)):
                    perms.append(p), flips.append(1), chars.append(chi * self.spin_inversion)
            self._group = Group(perms, flips, chars, n)
        return self._group


class Operator:
    """Compiled Hamiltonian: what the reference holds as ``ls_hs_operator`` (src/FFI.chpl:109-119)."""

    def __init__(self, basis: Basis, term_specs):
        self.basis = basis
        n = basis.number_sites
        acc = {}
        for spec in term_specs:
            if "expression" in spec:
                k, mat = expression_matrix(spec["expression"])
            else:
                mat = np.array(spec["matrix"], dtype=complex)
                k = mat.shape[0].bit_length() - 1
            for sites in spec["sites"]:
                sites = [int(s) for s in sites]
                if len(sites) != k or len(set(sites)) != k or min(sites) < 0 or max(sites) >= n:
                    raise ValueError(f"bad site tuple {sites}")
                mask = sum(1 << s for s in sites)

                def spread(local):
                    return sum(1 << st for pos, st in enumerate(sites) if (local >> (k - 1 - pos)) & 1)

                for out_idx in range(1 << k):
                    for in_idx in range(1 << k):
                        if mat[out_idx, in_idx] != 0:
                            r = spread(in_idx)
                            key = (mask, r, r ^ spread(out_idx))
                            acc[key] = acc.get(key, 0j) + complex(mat[out_idx, in_idx])
        # equal (m, r, x) contributions of different expressions are merged at compile time:
        # sigma^x sigma^x + sigma^y sigma^y cancels on parallel spins and leaves the flip-flop term with coefficient 2
        off = [(v, m, r, x, 0) for (m, r, x), v in sorted(acc.items()) if x != 0 and v != 0]
        diag = [(v, m, r, 0, 0) for (m, r, x), v in sorted(acc.items()) if x == 0 and v != 0]
        self.off_diag, self.diag = Terms(off), Terms(diag)

    def number_off_diag_terms(self):       # ls_hs_operator_max_number_off_diag (src/ForeignTypes.chpl:228-229)
        return int(len(np.unique(self.off_diag.x)))

    def number_diag_terms(self):
        return len(self.diag)

    def is_real(self):
        return bool(np.all(self.off_diag.v.imag == 0) and np.all(self.diag.v.imag == 0))


def load_model(path: str):
    """-> (Basis, Operator) of a data/*.yaml model input."""
    with open(path, encoding="utf-8") as f:
        conf = yaml.safe_load(f)
    basis = Basis(conf["basis"])
    return basis, Operator(basis, conf["hamiltonian"]["terms"])
