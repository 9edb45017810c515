import os
import sys

import pytest

# The loopback multi-rank tests run several ranks' streams on ONE device, some of them holding a spinning wait kernel: give every
# stream its own hardware queue so a push kernel never queues behind another rank's wait (must be set before CUDA initialises).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


@pytest.fixture(scope="session")
def gpu_engine_factory():
    """ReplayEngine factory for -m gpu tests. No GPU => the test FAILS (never a silent CPU path),
    unless it was collected by a run that did not ask for GPU tests."""
    from surge_b200 import ReplayEngine

    def make(program=None):
        e = ReplayEngine(0)
        if program is not None:
            e.register_program(program)
        return e

    return make
