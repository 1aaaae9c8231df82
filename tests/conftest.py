import json
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["config"] = json.loads(str(g.pop("config_json")))
    g["steps"] = [int(s) for s in g["steps"]]
    return g


def block_rel_err(a, b):
    """max over cells of |a - b| / max|b| over the cell's (block, field) -- the per-cell parity norm:
    momenta cross zero inside a block, so each cell is compared against its block's scale."""
    scale = np.abs(b).max(axis=(-2, -1), keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b) / scale).max())


def block_rel_err_q(a, b):
    """The same norm for conserved_q = (sigma, Sr, Lz) (conserve_linear_p=0): Sr = x.p is a difference of two O(r |p|)
    terms that nearly cancel in a Keplerian disk (it can be 1e-3 of them over a whole block), so Sr and Lz = x cross p
    are both measured against their common scale r |p| = hypot(Sr, Lz), maximised over the block."""
    scale = np.abs(b).max(axis=(-2, -1), keepdims=True)
    rp = np.hypot(b[:, 1], b[:, 2]).max(axis=(-2, -1))
    scale[:, 1, 0, 0] = rp
    scale[:, 2, 0, 0] = rp
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b) / scale).max())


@pytest.fixture(scope="session")
def golden():
    return load_golden
