import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


os.environ.setdefault("CAPDEC_POISON_WORKSPACE", "1")   # catch reads of uninitialised workspace


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product library must exist before the package can be imported (no fallback): build it if absent
    lib = os.path.join(ROOT, "image-captioning-ml-project_b200", "csrc", "libcapdec.so")
    if not os.path.isfile(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
