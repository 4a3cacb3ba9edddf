import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_relerr(a, b):
    """max over the batch axis of the relative Frobenius error (the parity metric of BASELINE.md)."""
    return max(relerr(x, y) for x, y in zip(a, b))
