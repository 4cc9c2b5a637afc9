"""pytest configuration: markers, repo on sys.path, golden loaders."""
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


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def env_cases():
    z = load_golden("env_cases.npz")
    cases = {}
    for name in z["names"]:
        name = str(name)
        cases[name] = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
    return cases


@pytest.fixture(scope="session")
def model_logits():
    return load_golden("model_logits.npz")


@pytest.fixture(scope="session")
def train_step():
    return load_golden("train_step.npz")


@pytest.fixture(scope="session")
def graph_layers():
    return load_golden("graph_layers.npz")


def load_weights(name):
    z = load_golden(f"weights_{name}.npz")
    return {k: z[k] for k in z.files}
