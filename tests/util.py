"""Shared helpers for the parity tests (the checker side: oracle + golden fixtures)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances: rel <= 1e-2 in bf16 mode, <= 1e-5 in the fp32 check mode, measured as
# max|a-b| / max|b| per tensor.  Where the reference's OWN fp32 run deviates from its fp64 run by more
# than that (gradients through train-mode BatchNorm at batch 4, see tests/golden/*.npz
# ref_fp32_dev_*), the fp32 bound is widened to 3x the reference's own deviation.
TOL_BF16 = 1e-2
TOL_FP32 = 1e-5


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.sqrt(((a - b) ** 2).sum()) / (np.sqrt((b ** 2).sum()) + 1e-300))


def digest(a, nsamp=64):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def ref_dev(g):
    return dict(zip([str(k) for k in g["ref_fp32_dev_keys"]], [float(v) for v in g["ref_fp32_dev_vals"]]))
