"""ORACLE -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
`--impl reference` legs).  See oracle/model.py for the parity-pinning status.
"""
