import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irspack_b200._lib import lib, check
from irspack_b200._ials_core import _ptr
n, K, m = 138493, 128, 148 * 32768
rng = np.random.default_rng(1)
Y = rng.standard_normal((n, K)).astype(np.float32)
idx = rng.integers(0, n, m).astype(np.int32)
w = np.ones(m, np.float32)
names = {0: "full", 4: "no convert/store", 16: "no gather loads", 20: "no gather, no convert", 8: "no MMA", 28: "barriers only"}
for jobs in (148 * 27, 148 * 4):
    for flags, name in names.items():
        G = np.zeros((K, K), np.float32); b = np.zeros(K, np.float32); T = np.zeros(128 * 512 + 16, np.float32)
        check(lib.ials_weighted_gram_debug(_ptr(Y), n, K, _ptr(idx), _ptr(w), m, jobs, ctypes.c_float(0.0), 0, _ptr(G), _ptr(b), _ptr(T), flags))
        ms = T[128 * 512 + 1]
        print(f"jobs {jobs:5d} (len {m // jobs:6d}) {name:24s}: {ms:8.3f} ms  -> {ms * 1e6 / (m / 148):7.1f} ns per neighbour per SM", flush=True)
