"""pytest config: registers the `gpu` marker and puts the product package + oracle on sys.path.

`-m "not gpu"`: oracle vs golden fixtures, host logic, C-ABI symbol export check (no compute calls).
`-m gpu`      : parity tests proper -- CUDA kernels through the C-ABI vs the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clip-for-dl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "head_golden.npz")))
