import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The oracle restatement is test infrastructure; build it on demand (gcc only).
    if not os.path.exists(os.path.join(ROOT, "oracle", "libvoxoracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "voxplat_b200", "libvpworldgen.so")):
        from voxplat_b200 import build
        build.build()


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
