"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu and call the product only through
the C ABI (ctypes binding in <package>/fdr.py); the oracle (oracle/) is used only as checker."""
import importlib.util
import os
import sys

import numpy as np
import pytest

# Several shards in one process, each with three streams and spinning barrier kernels: with the default 8 hardware queues
# unrelated streams share a queue and a waiting barrier could block the kernel it waits for.  Must be set before CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-implementation-of-frequency-domain-image-restoration-using-fft_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_fdr():
    return _load("fdr_b200_binding", os.path.join(PKG, "fdr.py"))


def load_oracle():
    return _load("fdr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fdr():
    return load_fdr()


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


@pytest.fixture(scope="session")
def gpu(fdr):
    """The binding, after checking the CUDA library loads and sees a device (no fallback)."""
    if fdr.device_count() < 1:
        pytest.fail("no CUDA device visible to libfdr_b200.so: " + fdr.lib().fdr_last_error().decode())
    return fdr


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def u8_gate(got, want):
    """BASELINE.json gate: |delta| <= 1 LSB on >= 99.9 % of pixels.  Returns (n_exact, n_off1, n_worse)."""
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    return int((d == 0).sum()), int((d == 1).sum()), int((d > 1).sum())
