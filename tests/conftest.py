import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def normwise(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a - ref| / max|ref| -- the 'rel 1e-5' definition of SURVEY.md sec.8c."""
    a, ref = a.double().cpu(), ref.double().cpu()
    den = float(ref.abs().max())
    return float((a - ref).abs().max()) / (den if den > 0 else 1.0)


# Post-optimiser weights cannot be held to 1e-5 normwise: Adam's update lr*m/(sqrt(v)+eps) has
# slope lr/eps = 1e5 w.r.t. gradient noise on elements with |g| <~ eps = 1e-8, so fp32
# summation-order noise of ~1e-10 in a gradient moves a weight by up to ~1e-5.  Measured here:
# the reference's own CPU path differs run-to-run by 4e-6 and from its fp64 run by 6e-7 after
# two steps (tiny shape).  Weights after N optimiser steps are therefore compared with an
# ABSOLUTE tolerance of 1e-2 * lr per step; the Adam kernel itself is checked tightly on
# identical gradients.
ADAM_LR = 1e-3
ADAM_STEP_ATOL = 1e-2 * ADAM_LR


def max_abs(a: torch.Tensor, ref: torch.Tensor) -> float:
    return float((a.double().cpu() - ref.double().cpu()).abs().max())
