import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name, inputs):
    """Golden outputs of the reference for case `name`; fails if the regenerated inputs differ from the minted ones."""
    from oracle import cases
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    assert np.uint32(g.pop("input_crc")) == cases.input_crc(inputs), \
        f"{name}: synthetic inputs regenerated on this host differ from the ones the golden was minted with"
    return g


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
