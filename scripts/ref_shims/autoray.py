"""Stand-in for autoray (numpy only).  Golden-generation infrastructure only (see README.md)."""
import numpy as _np


def do(fn, *a, like=None, **k):
    return getattr(_np, fn)(*a, **k)
