"""pytest configuration: registers the `gpu` marker and makes the harness-side bindings
importable.  `-m "not gpu"` covers the oracle (tests/oracle_binding.py) against the reference's
golden vectors, host-side planning and the C-ABI symbol table; `-m gpu` are the parity tests of
the CUDA path proper and call through the C ABI in include/swgn.h."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)


# Build the test infrastructure (oracle) and the product libraries before collection: test modules
# bind the libraries at import time.  Incremental (mtime based), a no-op when everything is fresh.
import __graft_entry__ as _ge  # noqa: E402

_ge.build_if_needed()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
