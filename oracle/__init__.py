"""Parity oracle for the BRISK hot path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this package; the product (ethzasl_brisk_b200) never does.
"""
