"""CPU oracle for the H.x hot path -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).

PARITY UNPINNED by reference artefacts (no runnable reference, no golden HDF5 in the tree); pinned
by oracle/dense_pin.py, exact dimensions and physics known answers instead.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.
"""
