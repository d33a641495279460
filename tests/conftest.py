import os
import sys

import pytest

# cap the BLAS / OpenMP pools before NumPy is imported: their threads busy-wait, so on a loaded host an un-capped pool makes the
# SuperLU / ARPACK heavy oracle tests crawl (measured: 8 s -> minutes with four other busy cores).  An explicit setting wins.
# Only on a GPU-less host (the shared build container): the GPU box is dedicated, and there the suite ran -- and was timed -- un-capped.
if not os.path.exists("/dev/nvidiactl"):
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_v, str(min(4, os.cpu_count() or 1)))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a minute on CPU")


@pytest.fixture(scope="session")
def fdfd():
    import fdfd_jl_b200 as m
    return m


@pytest.fixture(scope="session")
def ctx(fdfd):
    return fdfd.default_context()
