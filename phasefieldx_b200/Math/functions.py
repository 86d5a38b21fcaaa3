"""Macaulay brackets, API of reference `src/phasefieldx/Math/functions.py:10-41`."""


def macaulay_bracket_positive(x):
    """<x>+ = (x + |x|)/2 for floats or numpy arrays."""
    return 0.5 * (x + abs(x))


def macaulay_bracket_negative(x):
    """<x>- = (x - |x|)/2 for floats or numpy arrays."""
    return 0.5 * (x - abs(x))
