"""CPU oracle for the BLP scoring / loss / ranking hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(blp_b200/) never imports this.  Parity status: pinned against the reference
itself (see gen_golden.py and tests/golden/).
"""
