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


@pytest.fixture()
def X_small() -> sps.csr_matrix:
    # the reference's fixture, /root/reference/tests/conftest.py:8-16
    return sps.csr_matrix(
        np.asarray(
            [[1, 1, 2, 3, 4], [0, 1, 0, 1, 0], [0, 0, 1, 0, 0], [0, 0, 0, 0, 0]],
            dtype=float,
        )
    )
