"""ORACLE package -- CPU restatements of the reference hot path. TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and `--impl reference` legs.
The product package `esc_gnn_b200` must never import from here.
"""
