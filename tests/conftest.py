"""pytest configuration: marker registration and shared fixtures."""
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import kp_oracle

    kp_oracle.build()
    return kp_oracle


def free_gpu_memory():
    """Drop every cached plan and hand torch's cached blocks back to the driver: the tests that launch torchrun children
    on the same GPUs need the memory this process has accumulated (full-size tables are tens of GB)."""
    import gc

    import torch

    from kmerpapa_b200 import engine

    engine.clear_plans()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

