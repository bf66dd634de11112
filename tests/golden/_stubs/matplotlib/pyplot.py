"""Stub: the reference's viz module imports pyplot at package import; nothing here is called."""
def __getattr__(name):
    raise AttributeError("matplotlib stub: %s is not available" % name)
