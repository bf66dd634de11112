"""Minimal stand-in so the reference package imports without matplotlib (golden generation only)."""
def use(*a, **k):
    pass
