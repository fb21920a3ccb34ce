"""pytest configuration: the `gpu` marker, library builds, golden-vector access."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared libraries exist before any test touches them (nvcc cross-compiles without a GPU)."""
    from kindel_b200 import build

    build.build_engine()
    build.build_oracle()


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_npz():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def clip_golden():
    """Reference outputs for the deterministic clip-heavy cases (oracle/make_clip_golden.py)."""
    with open(os.path.join(GOLDEN, "clip_cases.json")) as fh:
        return json.load(fh)


def golden_input(entry):
    return os.path.join(GOLDEN, entry["input"])


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    if os.environ.get("KDL_SHIM_ENGINE"):  # development aid: GPU test bodies against the oracle (tests/shim_engine.py)
        import shim_engine

        shim_engine.install()
        return
    skip = pytest.mark.skip(reason="no CUDA device here; run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
