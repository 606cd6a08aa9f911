import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are SKIPPED (not failed) where they cannot run: no CUDA device, or the CUDA library has not been built.
    On a GPU box a missing library is an error of its own (tests/test_host_api.py) — the product never falls back."""
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    have_so = os.path.exists(os.path.join(ROOT, "infernos_b200", "libinfernos_b200.so"))
    if have_cuda and have_so:
        return
    why = "no CUDA device" if not have_cuda else "libinfernos_b200.so has not been built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
