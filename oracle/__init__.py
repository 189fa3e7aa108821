"""CPU oracle of the k-mer hot path.  TEST INFRASTRUCTURE ONLY (see ref_cpu.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product never imports it.
"""
