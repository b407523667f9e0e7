import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    """Register the marker and build the product library and the test-side oracle if they
    are stale -- before collection, because test modules query the library at import time
    (both builds are incremental and take seconds; nvcc cross-compiles on CPU)."""
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    from nanorq_b200 import build as b
    b.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
