import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if not has_cuda:
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(autouse=True)
def _poison_before_gpu_tests(request):
    """MMSUM_TEST_POISON=1: before every GPU test, fill the caching allocator's free blocks and every SM's shared / tensor
    memory with NaN bit patterns — a kernel or a torch.empty workspace whose result depends on what was there before then
    fails deterministically instead of once in ten runs ("A latent NaN source")."""
    if os.environ.get("MMSUM_TEST_POISON") == "1" and "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from multimodalsum_b200 import ops
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 29, 1 << 28, 1 << 27, 1 << 26, 1 << 26, 1 << 24, 1 << 24, 1 << 22, 1 << 20)]
            junk += [torch.full((1 << 16,), float("nan"), device="cuda") for _ in range(64)]
            junk += [torch.full((1 << 12,), float("nan"), device="cuda") for _ in range(256)]
            del junk
            ops.debug_poison()
            torch.cuda.synchronize()
    yield
