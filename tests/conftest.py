import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the CUDA extension is built in-tree and git-ignored: build it on a fresh checkout (nvcc cross-compiles
    # sm_100a without a GPU); the tests themselves fail loudly if it is still missing
    lib = os.path.join(ROOT, "ssr_eval_b200", "lib", "libssr_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-j4", "all"], cwd=ROOT, check=False, capture_output=True)


@pytest.fixture(scope="session")
def golden():
    """Fixtures produced by tests/golden/make_golden.py from the reference's own code."""
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
