import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from irspack_b200._lib import lib, check
from irspack_b200._ials_core import _ptr
np.set_printoptions(precision=4, linewidth=200, suppress=True)
n, K = 32, 128
rng = np.random.default_rng(1)
Y = rng.standard_normal((n, K)).astype(np.float32)
G64 = Y.astype(np.float64).T @ Y.astype(np.float64)
for flags in (0, 2, 1, 3):
    G = np.zeros((K, K), np.float32); b = np.zeros(K, np.float32); T = np.zeros(128 * 512 + 16, np.float32)
    check(lib.ials_weighted_gram_debug(_ptr(Y), n, K, None, None, n, 1, ctypes.c_float(0.0), 0, _ptr(G), _ptr(b), _ptr(T), flags))
    base = T[128 * 512:].view(np.uint32)[0]
    T = T[:128 * 512].reshape(128, 512)
    nzc = np.flatnonzero(np.abs(T).max(axis=0) > 0); nzl = np.flatnonzero(np.abs(T).max(axis=1) > 0)
    print("flags", flags, "tmem_base", hex(base), "G err", np.abs(G - G64).max(), "nonzero cols", (nzc.min(), nzc.max(), len(nzc)) if len(nzc) else None,
          "lanes", (nzl.min(), nzl.max(), len(nzl)) if len(nzl) else None)
    if len(nzc):
        print("  col300 lanes 0..3,127:", T[[0, 1, 2, 3, 127], 300], " T[0:3,0:5]", T[0:3, 0:5].ravel())
