"""IALSRecommender.get_score for a large arbitrary index set takes one device GEMM over the
gathered embeddings (ials.py `_score_embeddings`) instead of one launch per user; both routes
must return what the reference's numpy expression returns (ials.py:477-481)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_get_score_of_a_large_arbitrary_index_set():
    from irspack_b200 import IALSRecommender
    from irspack_b200.synth import synth_csr

    X = synth_csr(900, 400, 20000, seed=4)
    rec = IALSRecommender(X, n_components=24, alpha0=0.1, reg=0.02, train_epochs=2).learn()
    rng = np.random.default_rng(0)
    idx = rng.permutation(900)[:300]
    assert idx.size >= rec._GATHER_MIN_ROWS and not np.all(np.diff(idx) == 1)
    want = rec.get_user_embedding()[idx] @ rec.get_item_embedding().T
    got = rec.get_score(idx)
    assert got.shape == (300, 400) and got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    few = idx[:5]  # the per-user route
    np.testing.assert_allclose(rec.get_score(few), want[:5], rtol=2e-5, atol=2e-5)
    seen = rec.get_score_remove_seen(idx)
    assert np.all(np.isneginf(seen[X[idx].nonzero()]))
    mask = np.ones_like(seen, dtype=bool)
    mask[X[idx].nonzero()] = False
    np.testing.assert_allclose(seen[mask], want[mask], rtol=2e-5, atol=2e-5)
    # scores from embeddings: cold users and arbitrary item embeddings also take the device GEMM
    emb = rec.compute_user_embedding(X[:70])
    np.testing.assert_allclose(rec.get_score_from_user_embedding(emb), emb @ rec.get_item_embedding().T,
                               rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(rec.get_score_cold_user(X[:70]), emb @ rec.get_item_embedding().T,
                               rtol=2e-5, atol=2e-5)
    new_items = rng.standard_normal((37, 24)).astype(np.float32)
    np.testing.assert_allclose(rec.get_score_from_item_embedding(idx[:50], new_items),
                               rec.get_user_embedding()[idx[:50]] @ new_items.T, rtol=2e-5, atol=2e-5)
    assert rec.get_score_from_user_embedding(np.zeros((0, 24), np.float32)).shape == (0, 400)
