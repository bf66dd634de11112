"""Numerical operators of the flow hot path; every function here is a thin NumPy-in / NumPy-out shim over a
stage-level entry point of libb200flow.so (same names and argument meaning as optical_flow/utils of the reference)."""
