"""YAML model inputs -> flat host-side descriptions of basis and operator.

Mirror of ``loadConfigFromYaml`` (reference: src/ForeignTypes.chpl:261-283), whose work is done
by the third-party ``ls_hs_load_yaml_config`` in the reference.  The YAML schema is the one of
/root/reference/data/*.yaml (and data/old/*.yaml):

    basis: {number_spins, hamming_weight, spin_inversion, symmetries: [{permutation, sector}]}
    hamiltonian: {terms: [{expression | matrix, sites}]}
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import comb

import numpy as np
import yaml

from .expr import TermTable, compile_terms, max_number_off_diag
from .symmetry import SymmetryGroup, build_group


@dataclass
class BasisSpec:
    """Flat description of a spin basis (what ``ls_hs_basis`` carries, src/FFI.chpl:94-105)."""
    number_sites: int
    hamming_weight: int | None
    spin_inversion: int            # 0 = none
    generators: list[dict] = field(default_factory=list)
    _group: SymmetryGroup | None = None

    # -- flags, named after the reference's Basis methods (src/ForeignTypes.chpl:82-100) --------
    def is_hamming_weight_fixed(self) -> bool:
        return self.hamming_weight is not None

    def has_spin_inversion_symmetry(self) -> bool:
        return self.spin_inversion != 0

    def has_permutation_symmetries(self) -> bool:
        return len(self.generators) > 0

    def requires_projection(self) -> bool:
        return self.has_permutation_symmetries() or self.has_spin_inversion_symmetry()

    def is_state_index_identity(self) -> bool:
        return not self.requires_projection() and not self.is_hamming_weight_fixed()

    @property
    def group(self) -> SymmetryGroup:
        if self._group is None:
            self._group = build_group(self.number_sites, self.generators, self.spin_inversion)
        return self._group

    def min_state_estimate(self) -> int:
        """Smallest candidate state (mirror of ``ls_hs_min_state_estimate``, src/FFI.chpl:147)."""
        if self.hamming_weight is None:
            return 0
        return (1 << self.hamming_weight) - 1

    def max_state_estimate(self) -> int:
        """Largest candidate state (mirror of ``ls_hs_max_state_estimate``, src/FFI.chpl:148).

        With spin inversion every representative satisfies s <= s ^ mask, so the top site is
        never set; the bound excludes it."""
        n, w = self.number_sites, self.hamming_weight
        if w is None:
            hi = (1 << n) - 1
            if self.has_spin_inversion_symmetry():
                hi >>= 1
            return hi
        hi = ((1 << w) - 1) << (n - w)
        if self.has_spin_inversion_symmetry() and n > 0 and w > 0:
            # largest weight-w state with the top bit clear
            hi = ((1 << w) - 1) << (n - 1 - w) if n - 1 >= w else 0
        return hi

    def number_candidates(self) -> int:
        n, w = self.number_sites, self.hamming_weight
        return comb(n, w) if w is not None else 1 << n


@dataclass
class OperatorSpec:
    """Flat description of an operator on a basis: compiled non-branching term tables."""
    basis: BasisSpec
    off_diag: TermTable
    diag: TermTable
    name: str = ""

    def number_diag_terms(self) -> int:       # src/ForeignTypes.chpl:222-226
        return len(self.diag)

    def number_off_diag_terms(self) -> int:   # src/ForeignTypes.chpl:228-233
        return max_number_off_diag(self.off_diag)

    def is_real(self) -> bool:                # src/ForeignTypes.chpl:258
        return self.off_diag.is_real() and self.diag.is_real()


def basis_from_dict(d: dict) -> BasisSpec:
    n = int(d["number_spins"])
    if not 0 < n <= 64:
        raise ValueError("bases with more than 64 bits are not yet implemented")  # DMV:1099-1100
    hw = d.get("hamming_weight", None)
    inv = d.get("spin_inversion", None)
    return BasisSpec(
        number_sites=n,
        hamming_weight=None if hw is None else int(hw),
        spin_inversion=0 if inv is None else int(inv),
        generators=[{"permutation": [int(v) for v in g["permutation"]], "sector": int(g.get("sector", 0))}
                    for g in (d.get("symmetries") or [])],
    )


def operator_from_dict(d: dict, basis: BasisSpec) -> OperatorSpec:
    off, diag = compile_terms(d["terms"], basis.number_sites)
    return OperatorSpec(basis=basis, off_diag=off, diag=diag, name=str(d.get("name", "")))


def load_config_from_yaml(filename: str, hamiltonian: bool = True):
    """``loadConfigFromYaml(filename, hamiltonian=true)`` -> (basis, hamiltonian)."""
    with open(filename, "r", encoding="utf-8") as f:
        conf = yaml.safe_load(f)
    if "basis" not in conf:
        raise ValueError(f"failed to load Config from '{filename}'")
    basis = basis_from_dict(conf["basis"])
    if not hamiltonian:
        return basis
    if "hamiltonian" not in conf:
        raise ValueError(f"'{filename}' does not contain a Hamiltonian")
    return basis, operator_from_dict(conf["hamiltonian"], basis)
