import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def mano_tables(side):
    t = load_golden("mano_" + side)
    return t


def check_grad_digest(golden, name, g, rtol, floor=1e-9, outliers=0.0):
    """Compare a gradient tensor with its record in a golden made by make_golden.grad_digest:
    element-wise within rtol of the tensor's largest magnitude (``floor`` is the absolute
    tolerance for gradients that are analytically zero, e.g. conv biases in front of BatchNorm).
    ``outliers``: fraction of elements allowed outside that band (but inside 10x of it / 5 % of the scale) - in fp32 a
    near-tie in a max-pool or a ReLU at 0 can route one gradient differently from the fp64 golden."""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    full = name in golden
    ref = (golden[name] if full else golden[name + "@s97"]).astype(np.float64).reshape(-1)
    got = g if full else g[::97]
    scale = float(golden[name + "@norm"]) / np.sqrt(g.size) if not full else np.abs(ref).max()
    atol = max(rtol * max(np.abs(ref).max(), scale), floor)
    bad = np.abs(got - ref) > atol + rtol * np.abs(ref)
    assert bad.mean() <= outliers, "%s: %d of %d elements outside tolerance (max diff %.3g, atol %.3g)" % (
        name, bad.sum(), bad.size, np.abs(got - ref).max(), atol)
    # re-routed elements carry a whole different contribution: bounded by a fraction of the tensor's scale
    hard = np.maximum(10 * (atol + rtol * np.abs(ref)), 0.05 * max(np.abs(ref).max(), scale))
    assert (np.abs(got - ref) <= hard).all(), name
    if not full:
        norm = float(golden[name + "@norm"])
        assert abs(np.linalg.norm(g) - norm) <= max(rtol * norm, floor * np.sqrt(g.size)), name
