"""Test configuration. `-m "not gpu"` runs the oracle/golden, host-logic and C-ABI export checks on the CPU;
`-m gpu` runs the parity tests proper (CUDA path vs oracle / golden vectors) on a B200."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu")
    config.addinivalue_line("markers", "slow: full-width oracle runs (about a minute of CPU)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / f"{name}.npz")
    return load


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """Element-wise figure next to rel_l2: max |a - b| / max |b| (the largest single-element deviation in units of the
    reference's dynamic range)."""
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
