import os
import sys

# the reference's scripted morphology breaks under torch 2.11 (SURVEY B#13); nothing in the product uses
# torch.jit, so run the whole suite in eager mode.  Must happen before torch is imported.
os.environ.setdefault("PYTORCH_JIT", "0")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack_mask(fix, key="mask"):
    shape = tuple(int(v) for v in fix["shape"])
    n = int(np.prod(shape))
    return np.unpackbits(fix[key])[:n].reshape(shape)


@pytest.fixture(scope="session")
def golden():
    return load_golden
