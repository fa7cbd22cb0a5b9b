"""The streamed Schur kernel is opt-in (SWGN_SCHUR_STREAM=1, read once per process by the planner), so its GPU parity
check runs in a subprocess with the variable set: tests/stream_gpu_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_streamed_schur_kernel_matches_the_oracle():
    env = dict(os.environ, SWGN_SCHUR_STREAM="1")
    r = subprocess.run([sys.executable, os.path.join(HERE, "stream_gpu_check.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "stream mixed batch ok" in r.stdout
