"""CPU oracle for the H.x hot path -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).

PARITY UNPINNED by reference artefacts (no runnable reference, no golden HDF5 in the tree); pinned
by independent constructions instead: Kronecker matrices (oracle/dense_pin.py), the Heisenberg definition
at full size, exact dimensions, Bethe-ansatz and literature ground-state energies (tests/test_oracle_pins.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.
"""
