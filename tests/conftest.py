import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    w = {k[2:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("w.")}
    g = {k[2:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("g.")}
    return rec, w, g


def split_steps(flat, sizes):
    """[sum(sizes), ...] -> list of per-step tensors."""
    out, o = [], 0
    for n in sizes:
        out.append(flat[o:o + n])
        o += n
    return out


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
