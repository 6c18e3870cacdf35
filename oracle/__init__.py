"""CPU oracle for the Style-SeqCVAE `var_updown` decoder hot path — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product package (style-seqcvae_b200/) must never import anything from here.
"""
