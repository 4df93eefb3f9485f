import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_l1(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300))


def interior(a):
    return a[3:-3, 3:-3] if a.ndim >= 2 and a.shape[1] > 9 else a[3:-3]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
