import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed_global_rng():
    """Module constructors (nn.Linear, xavier_normal_) draw from torch's global RNG: fix it per test so that every run of the
    suite sees the same weights (results then repeat to fp32 atomic-order noise)."""
    import torch
    torch.manual_seed(1234)
    yield


# hypothesis: the suite runs a FIXED set of examples (derandomize) so that results repeat; CNF_HYPOTHESIS=explore searches wider
try:
    from hypothesis import settings as _hs
    _hs.register_profile("ci", derandomize=True, deadline=None)
    _hs.register_profile("explore", max_examples=150, derandomize=False, deadline=None, print_blob=True)
    _hs.load_profile(os.environ.get("CNF_HYPOTHESIS", "ci"))
except ImportError:      # property tests are skipped at import time without hypothesis
    pass
