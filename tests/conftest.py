import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/wsimod")
    skip_ref = pytest.mark.skip(reason="/root/reference not mounted")
    have_cuda = None
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords:
            if have_cuda is None:  # only pay for `import torch` when a GPU test was selected
                import torch

                have_cuda = torch.cuda.is_available()
            if not have_cuda:
                # no device: skip.  On a GPU box the tests run and FAIL when libwsi_b200.so is missing
                # (ExtensionMissingError) -- there is no CPU fallback to fall back to.
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
