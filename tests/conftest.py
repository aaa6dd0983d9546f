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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz fixture written by tests/golden/make_golden.py from the reference."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
        nb = len([k for k in z.files if k.startswith("batch/")])
        self.batch = [torch.from_numpy(z[f"batch/{i}"]) for i in range(nb)]
        self.ratings = torch.from_numpy(z["ratings"])
        self.out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out/")}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
        self.meta = {k[5:]: z[k].tolist() for k in z.files if k.startswith("meta/")}
        self.model = self.meta["model"]


GOLDEN_CASES = ["deepconn_small", "deepconn_edge", "deepconn_multik", "deepconn_odd", "narre_small",
                "narre_h150ish", "dual_att_small"]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


def rel_err(a: torch.Tensor, b: torch.Tensor, floor: float = 1e-12) -> float:
    """max |a-b| / max(|b|_inf, floor): the relative error the parity tolerances are stated in.

    `floor` keeps quantities that are mathematically ~0 (e.g. d loss / d b_2 of NARRE's attention, which
    vanishes by softmax shift-invariance up to the 1e-8 epsilon) from being compared as rounding noise."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor)) if b.numel() else 0.0


def grad_floor(key: str) -> float:
    """d loss / d b_2 of NARRE's LinearAttention is mathematically ~0 (softmax is shift-invariant up to the
    1e-8 epsilon, narre.py:58), so what the reference stores there is fp32 cancellation noise (~1e-10):
    compare it on an absolute scale instead of relative to itself."""
    return 1e-3 if key.endswith("att.b_2") else 1e-12
