import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture
def cpu_backend():
    """Installs the CPU test double of libvqvae_b200 (tests/fake_backend.py) for host-logic tests."""
    import vqvae_b200 as V
    from tests.fake_backend import FakeBackend
    prev = (V._lib._BACKEND, V._lib._DEVICE)
    V._lib.set_backend(FakeBackend(), "cpu")
    V.keras_compat.reset_name_counters()
    V.keras_compat.set_seed(0)
    yield V
    V._lib._BACKEND, V._lib._DEVICE = prev


@pytest.fixture
def gpu():
    """The real library on cuda:0; fails (never skips silently) if the extension is not the one running."""
    import torch
    import vqvae_b200 as V
    assert torch.cuda.is_available(), "gpu-marked test run without a GPU"
    V._lib._BACKEND = None
    V._lib.lib()
    assert V._lib.is_native(), "libvqvae_b200.so is not the active backend"
    V.keras_compat.reset_name_counters()
    V.keras_compat.set_seed(0)
    return V
