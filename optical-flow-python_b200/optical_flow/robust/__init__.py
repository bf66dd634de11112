"""Robust penalty functions (device-evaluated).  Mirrors optical_flow/robust of the reference."""
from optical_flow.robust.robust_function import RobustFunction, PENALTY_MAP

__all__ = ["RobustFunction", "PENALTY_MAP"]
