"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on a CPU-only container (oracle vs golden vectors, host logic, C-ABI symbol
check, gloo multi-process tests).  `-m gpu` are the parity tests proper: they call the CUDA path
through the C-ABI library on a real B200 and compare with the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


def rel_l2(a, b):
    import torch
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def free_port() -> int:
    """A TCP port that is free right now on 127.0.0.1 (rendezvous port of the gloo multi-process tests)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]
