"""CPU oracle for the csparse_cuda hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never from the product package.  See csparse_oracle.c for the contract.
"""
