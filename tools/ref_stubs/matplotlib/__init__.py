"""Import-only stub: plotting is never executed by the golden generator."""
def use(*a, **k):
    pass
