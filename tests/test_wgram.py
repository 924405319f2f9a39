"""tcgen05 weighted Gram (irspack_b200/csrc/wgram.cu) against float64 numpy.

Tolerance: the kernel evaluates u u^T with u = sqrt(w) y split into TF32 hi/lo parts and
drops only the lo*lo term (2^-22 relative), accumulating in fp32 inside the tensor core,
which truncates once per MMA (8 neighbours): a job of L neighbours carries a relative
bias of at most (L / 8) * 2^-24 on a sum of same-signed terms.  Stated tolerance:
    |G - G64| <= (2e-6 + 6e-8 * L / 8) * max|G64|,   L = longest job
(the diagonal of a Gram matrix is a sum of non-negative terms, so max|G64| bounds every
entry's sum of absolute terms).  The trainer keeps L <= 1024 (IALS_HEAVY_JOB_LEN) for the
per-row Grams and ~n/148 for K1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tol(m, jobs):
    return 2e-6 + 6e-8 * (-(-m // jobs)) / 8


def ref(Y, idx, w, bias):
    Y64 = Y.astype(np.float64)
    rows = Y64 if idx is None else Y64[idx]
    ww = np.ones(rows.shape[0]) if w is None else w.astype(np.float64)
    return (rows * ww[:, None]).T @ rows, ((bias + ww)[:, None] * rows).sum(axis=0)


@pytest.mark.parametrize("n,K,jobs", [(32, 128, 1), (8, 128, 1), (1000, 128, 3), (5000, 64, 7),
                                      (257, 20, 2), (40000, 128, 148), (3, 5, 4)])
def test_plain_gram(n, K, jobs):
    from irspack_b200.ops import weighted_gram

    rng = np.random.default_rng(n + K)
    Y = (rng.standard_normal((n, K)) * rng.uniform(0.01, 3.0, size=(1, K))).astype(np.float32)
    G, b = weighted_gram(Y, n_jobs=jobs, bias=0.25)
    G64, b64 = ref(Y, None, None, 0.25)
    assert np.abs(G - G64).max() <= tol(n, jobs) * np.abs(G64).max()
    assert np.abs(b - b64).max() <= 1e-5 * (np.abs(Y).astype(np.float64).sum(axis=0).max() * 1.25)
    np.testing.assert_array_equal(G, G.T)


@pytest.mark.parametrize("n,m,K,jobs", [(500, 37, 128, 1), (500, 4096, 128, 5), (26744, 35000, 128, 9), (26744, 35000, 128, 40),
                                        (100, 1, 128, 1), (100, 0, 128, 2), (64, 333, 48, 4)])
def test_gathered_weighted_gram(n, m, K, jobs):
    from irspack_b200.ops import weighted_gram

    rng = np.random.default_rng(m + 7)
    Y = (rng.standard_normal((n, K)) * 0.1).astype(np.float32)
    idx = rng.integers(0, n, m).astype(np.int32)
    w = rng.choice([0.0, 0.5, 1.0, 2.0, 4.7], size=m).astype(np.float32)
    G, b = weighted_gram(Y, idx, w, n_jobs=jobs, bias=0.1)
    G64, b64 = ref(Y, idx, w, 0.1)
    scale = max(np.abs(G64).max(), 1e-30)
    assert np.abs(G - G64).max() <= tol(m, jobs) * scale
    bscale = max((np.abs(Y[idx]).astype(np.float64) * (0.1 + w)[:, None]).sum(axis=0).max(), 1e-30)
    assert np.abs(b - b64).max() <= 1e-5 * bscale


def test_tf32_alone_would_fail_this_tolerance():
    """Documents why the hi/lo split is there: a single-pass TF32 product is ~1e-3 off."""
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((4096, 128)).astype(np.float32)
    hi = (Y.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
    G64 = Y.astype(np.float64).T @ Y.astype(np.float64)
    assert np.abs(hi.T @ hi - G64).max() > 1e-4 * np.abs(G64).max()


def test_negative_weights_rejected():
    from irspack_b200.ops import weighted_gram

    Y = np.ones((4, 8), np.float32)
    with pytest.raises(ValueError):
        weighted_gram(Y, np.array([0, 1], np.int32), np.array([1.0, -1.0], np.float32))


@pytest.mark.parametrize("n,m,K,jobs", [(300, 200, 256, 1), (300, 37, 256, 1), (2000, 5000, 256, 3),
                                        (500, 1000, 240, 2), (64, 8, 136, 1)])
def test_gram_of_256_column_factors(n, m, K, jobs):
    """K = 256 Cholesky rank updates on the tensor cores (wgram.cu wgram256_kernel: one pass,
    W = 1/2 hi (hi + 2 lo)^T, G = W + W^T) against float64 numpy.  The tolerance is the
    128-column operator's with TWO accumulating instructions per 8 neighbours instead of one:
    the large and the small products share an accumulator here (all 512 TMEM columns are in
    use), and the tensor core's fp32 accumulation truncates, so the error of the sign-definite
    diagonal grows with the number of accumulations (observed 2.2e-5 at 1667 neighbours per job,
    r02m; the cross block, whose terms cancel, stays below 1e-6)."""
    from irspack_b200.ops import weighted_gram256

    rng = np.random.default_rng(m + K)
    Y = (rng.standard_normal((n, K)) * rng.uniform(0.05, 2.0, size=(1, K))).astype(np.float32)
    idx = rng.integers(0, n, m).astype(np.int32)
    w = rng.choice([0.5, 1.0, 2.0, 4.7], size=m).astype(np.float32)
    G, b = weighted_gram256(Y, idx, w, n_jobs=jobs, bias=0.1)
    G64, b64 = ref(Y, idx, w, 0.1)
    scale = np.abs(G64).max()
    err = {name: float(np.abs(G[r, c] - G64[r, c]).max() / scale)
           for name, (r, c) in {"G00": (slice(0, 128), slice(0, 128)), "G11": (slice(128, K), slice(128, K)),
                                "G01": (slice(0, 128), slice(128, K))}.items()}
    assert max(err.values()) <= 2e-6 + 2 * (tol(m, jobs) - 2e-6), err
    np.testing.assert_array_equal(G, G.T)
    assert np.abs(b - b64).max() <= 1e-5 * (np.abs(Y[idx]).astype(np.float64) * (0.1 + w)[:, None]).sum(axis=0).max()
