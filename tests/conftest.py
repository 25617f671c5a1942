import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    import mray_b200
    assert torch.cuda.is_available(), "gpu-marked test on a box without CUDA"
    ctx = mray_b200.Context(0)
    ctx.set_stream(torch.cuda.current_stream())
    yield ctx
    ctx.close()
