"""``get_n_threads`` as in /root/reference/src/irspack/_threading.py:5-17.

The GPU kernels ignore the thread count; it is still validated (``n_threads``
must be > 0, IALSTrainer.hpp:81-83) and it is what the CPU baseline uses."""
import os
from typing import Optional


def get_n_threads(n_threads: Optional[int]) -> int:
    if n_threads is not None:
        return n_threads
    try:
        cand = os.environ.get("IRSPACK_NUM_THREADS_DEFAULT", os.cpu_count())
        return int(cand or 1)
    except Exception:
        raise ValueError('failed to interpret "IRSPACK_NUM_THREADS_DEFAULT" as an integer.')
