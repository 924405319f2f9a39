import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without a CUDA device, so a plain
    `pytest tests` works on a CPU box; on the GPU box they run and the library must load."""
    try:
        import irspack_b200

        have_gpu = irspack_b200.device_count() > 0
    except Exception:  # library missing / not loadable: the CPU-side tests report that themselves
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture()
def X_small() -> sps.csr_matrix:
    """The 4 x 5 interaction matrix the reference's tests train on
    (/root/reference/tests/conftest.py:8-16): a dense first user, two sparse ones and an
    empty last row.  Given here as (row, column, value) triplets."""
    triplets = [(0, 0, 1), (0, 1, 1), (0, 2, 2), (0, 3, 3), (0, 4, 4), (1, 1, 1), (1, 3, 1), (2, 2, 1)]
    r, c, v = (np.asarray(x) for x in zip(*triplets))
    return sps.csr_matrix((v.astype(float), (r, c)), shape=(4, 5))
