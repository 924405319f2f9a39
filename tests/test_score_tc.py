"""The fused tcgen05 score + seen-mask + top-k kernel (csrc/score_tc.cu) against the CPU oracle
(float64 scores + the reference's (-score, index) ordering, evaluator.cpp:324-355) and against
the three-kernel FP32 SIMT path (cutoffs above 128).

Tolerances: score blocks rtol = atol = 2e-5 (the reference's own, tests/recommenders/
test_ials.py:564-570); top-k lists identical except that two items whose float64 scores differ
by less than 1e-5 * max|score| may swap; on exactly representable inputs (small-integer
factors: every product and sum is exact in any order) the lists must be IDENTICAL, ties
included.
"""
import numpy as np
import pytest
import scipy.sparse as sps

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core():
    import irspack_b200

    if irspack_b200.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible")
    from irspack_b200 import _ials_core

    return _ials_core


def trainer(core, X, K, user, item):
    cfg = core.IALSModelConfigBuilder().set_K(K).set_alpha0(0.1).set_reg(0.05).build()
    g = core.IALSTrainer(cfg, X)
    g.user, g.item = user, item
    return g


def oracle_topk(user, item, k, mask=None):
    s = user.astype(np.float64) @ item.astype(np.float64).T
    if mask is not None:
        s[sps.csr_matrix(mask).nonzero()] = -np.inf
    # every row is evaluated: give each one ground-truth item
    gt = sps.csr_matrix((np.ones(s.shape[0]), (np.arange(s.shape[0]), np.zeros(s.shape[0], int))),
                        shape=s.shape)
    _, rec, cnt = oracle.topk_metrics(s, gt, k)
    return s, rec, cnt


def check_lists(got, cnt, want, want_cnt, s64, rel=1e-5):
    np.testing.assert_array_equal(cnt, want_cnt)
    bad = np.flatnonzero((got != want).any(axis=1))
    for r in bad:
        finite = s64[r][np.isfinite(s64[r])]
        tol = rel * (np.abs(finite).max() if finite.size else 0.0) + 1e-12
        for a, b in zip(got[r], want[r]):
            if a != b:
                assert a >= 0 and b >= 0 and abs(s64[r, a] - s64[r, b]) <= tol, (r, a, b)
    return len(bad)


@pytest.mark.parametrize("U,I,K,k", [
    (300, 5000, 128, 10),    # several catalogue splits (3 user tiles, 40 item tiles)
    (1000, 777, 128, 10),    # partial last item tile, partial last user tile
    (130, 3000, 96, 100),    # three feature chunks, 256-key candidate buffers
    (257, 2100, 64, 50),     # two feature chunks, 128-key buffers
    (64, 100, 32, 16),       # one tile of everything
    (500, 26744, 128, 10),   # the configs[1] catalogue
    (200, 1500, 120, 1),     # K not a multiple of 32 (zero-padded features), k = 1
])
def test_fused_topk_matches_oracle(core, U, I, K, k, monkeypatch):
    rng = np.random.default_rng(U + I + K)
    X = sps.random(U, I, density=min(0.05, 200.0 / I), random_state=3, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    user = rng.standard_normal((U, K)).astype(np.float32)
    item = rng.standard_normal((I, K)).astype(np.float32)
    g = trainer(core, X, K, user, item)
    s64, want, want_cnt = oracle_topk(user, item, k, X)
    got, cnt, sc = g.recommend(0, U, k, mask="train", return_scores=True)
    n_diff = check_lists(got, cnt, want, want_cnt, s64)
    assert n_diff <= max(2, U // 50)
    ok = got >= 0
    rows = np.repeat(np.arange(U), k).reshape(U, k)
    np.testing.assert_allclose(sc[ok], s64[rows[ok], got[ok]], rtol=2e-5, atol=2e-5)
    # sub-block, no mask
    b, e = U // 3, U // 3 + min(U - U // 3, 77)
    s64n, want, want_cnt = oracle_topk(user[b:e], item, k)
    got, cnt = g.recommend(b, e, k, mask=None)
    check_lists(got, cnt, want, want_cnt, s64n)
    # a cutoff above what the fused kernel keeps per row (128) takes the three-kernel FP32 SIMT
    # path (score.cu): same tie rule, so its leading k columns are the same lists
    if I >= 150:
        s64w, want_w, want_cnt_w = oracle_topk(user[b:e], item, 150)
        got2, cnt2 = g.recommend(b, e, 150, mask=None)
        check_lists(got2, cnt2, want_w, want_cnt_w, s64w)


def test_fused_topk_exact_on_integer_factors(core):
    """Small-integer factors: all arithmetic is exact, scores tie massively; the lists must be
    exactly the reference's (-score, index) order."""
    rng = np.random.default_rng(5)
    U, I, K = 300, 4000, 128
    user = rng.integers(-1, 2, size=(U, K)).astype(np.float32)
    item = rng.integers(-1, 2, size=(I, K)).astype(np.float32)
    X = sps.random(U, I, density=0.02, random_state=4, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    g = trainer(core, X, K, user, item)
    for k in (1, 10, 40, 128):
        _, want, want_cnt = oracle_topk(user, item, k, X)
        got, cnt = g.recommend(0, U, k, mask="train")
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(cnt, want_cnt)


def test_fused_topk_mask_edge_cases(core):
    """Stored zeros do not mask (scipy's .nonzero()); fully masked users return nothing;
    fewer candidates than k pads with -1; a custom mask replaces the training rows."""
    rng = np.random.default_rng(6)
    U, I, K, k = 140, 300, 128, 20
    user = rng.integers(-2, 3, size=(U, K)).astype(np.float32)
    item = rng.integers(-2, 3, size=(I, K)).astype(np.float32)
    dense = (rng.random((U, I)) < 0.1).astype(np.float32)
    dense[3, :] = 1.0             # everything seen
    dense[4, :] = 1.0
    dense[4, 7::50] = 0.0         # 6 candidates left
    X = sps.csr_matrix(dense)
    # explicit stored zeros in row 5
    data = X.data.copy()
    row5 = slice(X.indptr[5], X.indptr[6])
    data[row5] = 0.0
    Xz = sps.csr_matrix((data, X.indices.copy(), X.indptr.copy()), shape=X.shape)
    g = trainer(core, Xz, K, user, item)
    mask_eff = Xz.copy()
    mask_eff.eliminate_zeros()
    _, want, want_cnt = oracle_topk(user, item, k, mask_eff)
    got, cnt = g.recommend(0, U, k, mask="train")
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)
    assert cnt[3] == 0 and (got[3] == -1).all()
    assert cnt[4] == 6 and (got[4, 6:] == -1).all()
    # custom mask for a sub-block
    custom = sps.csr_matrix((rng.random((40, I)) < 0.3).astype(np.float32))
    _, want, want_cnt = oracle_topk(user[100:140], item, k, custom)
    got, cnt = g.recommend(100, 140, k, mask=custom)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)


@pytest.mark.parametrize("K", [128, 96, 40])
def test_dense_scores_on_tensor_cores(core, K):
    rng = np.random.default_rng(K)
    U, I = 333, 1001
    user = rng.standard_normal((U, K)).astype(np.float32)
    item = rng.standard_normal((I, K)).astype(np.float32)
    g = trainer(core, sps.csr_matrix((U, I), dtype=np.float32), K, user, item)
    sc = core.IALSSolverConfigBuilder().build()
    want = user.astype(np.float64) @ item.astype(np.float64).T
    for b, e in [(0, U), (17, 193), (332, 333)]:
        got = g.user_scores(b, e, sc)
        np.testing.assert_allclose(got, want[b:e], rtol=2e-5, atol=2e-5)
    # 3xTF32 is as accurate as an FP32 dot product: compare the error with numpy's sgemm
    err_tc = np.abs(g.user_scores(0, U, sc) - want).max()
    err_f32 = np.abs(user @ item.T - want).max()
    assert err_tc <= 4 * err_f32 + 1e-6, (err_tc, err_f32)


# ---------------------------------------------------------------------------------------
# allow-lists (the Evaluator's recommendable items, evaluator.cpp:168-180) fused into the
# same kernel: ials_trainer_recommend_allowed
# ---------------------------------------------------------------------------------------
def lists_csr(lists):
    indptr = np.zeros(len(lists) + 1, np.int64)
    np.cumsum([len(x) for x in lists], out=indptr[1:])
    flat = np.concatenate([np.asarray(x, np.int32) for x in lists]) if lists else np.zeros(0, np.int32)
    return len(lists), indptr, flat.astype(np.int32)


def oracle_topk_allowed(user, item, k, lists, mask=None):
    s = user.astype(np.float64) @ item.astype(np.float64).T
    if mask is not None:
        s[sps.csr_matrix(mask).nonzero()] = -np.inf
    ok = np.zeros(s.shape, bool)
    for r in range(s.shape[0]):
        ok[r, np.asarray(lists[0] if len(lists) == 1 else lists[r], np.int64)] = True
    s[~ok] = -np.inf
    gt = sps.csr_matrix((np.ones(s.shape[0]), (np.arange(s.shape[0]), np.zeros(s.shape[0], int))),
                        shape=s.shape)
    _, rec, cnt = oracle.topk_metrics(s, gt, k)
    return s, rec, cnt


def test_fused_topk_allow_lists_exact(core):
    """Integer factors (exact arithmetic, massive ties): shared and per-user lists, with the
    training mask, a custom mask and none; sub-blocks; empty, full and short lists."""
    rng = np.random.default_rng(11)
    U, I, K = 200, 4500, 128
    user = rng.integers(-1, 2, size=(U, K)).astype(np.float32)
    item = rng.integers(-1, 2, size=(I, K)).astype(np.float32)
    X = sps.random(U, I, density=0.03, random_state=8, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    g = trainer(core, X, K, user, item)
    shared = [np.sort(rng.choice(I, 1700, replace=False))]
    per_user = [np.sort(rng.choice(I, int(n), replace=False))
                for n in rng.integers(0, 3000, size=U)]
    per_user[0] = np.zeros(0, np.int64)              # nothing recommendable
    per_user[1] = np.arange(I)                       # everything
    per_user[2] = np.array([5, 4499])                # fewer than k, first and last tiles
    per_user[3] = X[3].indices.astype(np.int64)      # only seen items: all masked by "train"
    per_user[3].sort()
    for k in (1, 10, 40, 128):
        for lists in (shared, per_user):
            _, want, want_cnt = oracle_topk_allowed(user, item, k, lists, X)
            got, cnt = g.recommend(0, U, k, mask="train", allowed=lists_csr(lists))
            np.testing.assert_array_equal(cnt, want_cnt)
            np.testing.assert_array_equal(got, want)
    assert cnt[0] == 0 and cnt[2] == 2 and cnt[3] == 0
    # sub-block: the lists are those of the block's rows
    b, e, k = 37, 171, 20
    custom = sps.csr_matrix((rng.random((e - b, I)) < 0.2).astype(np.float32))
    for mask, oracle_mask in (("train", X[b:e]), (None, None), (custom, custom)):
        for lists in (shared, per_user[b:e]):
            _, want, want_cnt = oracle_topk_allowed(user[b:e], item, k, lists, oracle_mask)
            got, cnt = g.recommend(b, e, k, mask=mask, allowed=lists_csr(lists))
            np.testing.assert_array_equal(cnt, want_cnt)
            np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("U,I,K,k", [(300, 26744, 128, 10), (130, 3000, 96, 100), (1000, 777, 64, 50)])
def test_fused_topk_allow_lists_random_factors(core, U, I, K, k):
    rng = np.random.default_rng(U + K)
    X = sps.random(U, I, density=min(0.05, 200.0 / I), random_state=3, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    user = rng.standard_normal((U, K)).astype(np.float32)
    item = rng.standard_normal((I, K)).astype(np.float32)
    g = trainer(core, X, K, user, item)
    per_user = [np.sort(rng.choice(I, int(n), replace=False)) for n in rng.integers(0, I, size=U)]
    for lists in ([np.sort(rng.choice(I, I // 3, replace=False))], per_user):
        s64, want, want_cnt = oracle_topk_allowed(user, item, k, lists, X)
        got, cnt, sc = g.recommend(0, U, k, mask="train", return_scores=True, allowed=lists_csr(lists))
        assert check_lists(got, cnt, want, want_cnt, s64) <= max(2, U // 50)
        ok = got >= 0
        rows = np.repeat(np.arange(U), k).reshape(U, k)
        np.testing.assert_allclose(sc[ok], s64[rows[ok], got[ok]], rtol=2e-5, atol=2e-5)


def test_allow_list_argument_checks(core):
    rng = np.random.default_rng(2)
    U, I, K = 40, 300, 64
    X = sps.random(U, I, density=0.05, random_state=1, format="csr", dtype=np.float32)
    g = trainer(core, X, K, rng.standard_normal((U, K)).astype(np.float32),
                rng.standard_normal((I, K)).astype(np.float32))
    with pytest.raises(ValueError):   # not ascending
        g.recommend(0, U, 5, allowed=lists_csr([[3, 2, 9]]))
    with pytest.raises(ValueError):   # duplicate
        g.recommend(0, U, 5, allowed=lists_csr([[3, 3, 9]]))
    with pytest.raises(ValueError):   # out of range
        g.recommend(0, U, 5, allowed=lists_csr([[3, I]]))
    with pytest.raises(ValueError):   # neither one list nor one per row
        g.recommend(0, U, 5, allowed=lists_csr([[1, 2]] * 3))
    with pytest.raises(ValueError):   # indptr / indices disagree
        g.recommend(0, U, 5, allowed=(1, np.array([0, 4]), np.array([1, 2], np.int32)))
    with pytest.raises(NotImplementedError):  # past the fused kernel's cutoff
        g.recommend(0, U, 150, allowed=lists_csr([np.arange(200)]))


@pytest.mark.parametrize("cutoff", [10, 150])
def test_evaluator_allow_lists_take_the_fused_path(core, cutoff, monkeypatch):
    """Evaluator(recommendable_items= / per_user_recommendable_items=) over an IALSRecommender:
    the fused path (cutoff <= 128) and the host score-block path give the same metrics, and
    with cutoff 150 the Evaluator falls back by itself."""
    from irspack_b200 import Evaluator, EvaluatorWithColdUser, IALSRecommender

    rng = np.random.default_rng(21)
    U, I, K = 260, 900, 32
    X = sps.csr_matrix((rng.random((U, I)) < 0.04).astype(np.float32))
    te = sps.csr_matrix((rng.random((U, I)) < 0.02).astype(np.float32))
    rec = IALSRecommender(X, n_components=K, alpha0=0.1, reg=0.05, train_epochs=1).learn()
    t = rec.trainer_as_ials.core_trainer   # integer factors: no near-ties between the two paths
    t.user = rng.integers(-2, 3, size=(U, K)).astype(np.float32)
    t.item = rng.integers(-2, 3, size=(I, K)).astype(np.float32)

    class HostOnly:  # no recommend_block: score blocks on the host + select_topk
        X_train_all, n_users, n_items = rec.X_train_all, U, I
        get_score_block = staticmethod(rec.get_score_block)

    shared = [int(i) for i in rng.permutation(I)[:400]]                    # unsorted on purpose
    per_user = [[int(i) for i in rng.choice(I, int(n))] for n in rng.integers(0, 700, U)]  # duplicates too
    other = sps.csr_matrix((rng.random((U, I)) < 0.1).astype(np.float32))
    calls = []
    orig = type(t).recommend
    monkeypatch.setattr(type(t), "recommend",
                        lambda self, *a, **kw: calls.append(kw.get("allowed") is not None) or orig(self, *a, **kw))
    for kw in (dict(recommendable_items=shared), dict(per_user_recommendable_items=per_user),
               dict(per_user_recommendable_items=per_user, masked_interactions=other)):
        ev = Evaluator(te, cutoff=cutoff, mb_size=64, **kw)
        calls.clear()
        got = ev.get_score(rec)
        assert calls and all(calls)
        want = ev.get_score(HostOnly())
        for key, v in want.items():
            assert got[key] == pytest.approx(v, abs=1e-12), key
    # offset blocks
    ev = Evaluator(te[100:], offset=100, cutoff=cutoff, per_user_recommendable_items=per_user[100:])
    assert ev.get_score(rec) == pytest.approx(ev.get_score(HostOnly()), abs=1e-12)
    # cold users: fold-in + fused lists against the host path
    if cutoff <= 128:
        evc = EvaluatorWithColdUser(X[:50], te[:50], cutoff=cutoff, per_user_recommendable_items=per_user[:50])

        class ColdHost:
            n_items = I
            get_score_cold_user = staticmethod(rec.get_score_cold_user)

        got, want = evc.get_score(rec), evc.get_score(ColdHost())
        for key, v in want.items():
            assert got[key] == pytest.approx(v, abs=1e-9), key


def test_recommend_users_picked_by_index(core):
    """ials_trainer_recommend_users: arbitrary order, repeats, each user's own training row as the
    mask (row map), custom masks and per-user allow-lists by position in the request."""
    rng = np.random.default_rng(17)
    U, I, K, k = 400, 3000, 96, 20
    user = rng.integers(-1, 2, size=(U, K)).astype(np.float32)
    item = rng.integers(-1, 2, size=(I, K)).astype(np.float32)
    X = sps.random(U, I, density=0.03, random_state=9, format="csr", dtype=np.float32)
    X.data[:] = 1.0
    g = trainer(core, X, K, user, item)
    pick = np.concatenate([rng.permutation(U)[:170], [5, 5, 399, 0]])
    _, want, want_cnt = oracle_topk(user[pick], item, k, X[pick])
    got, cnt, sc = g.recommend_users(pick, k, mask="train", return_scores=True)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)
    _, want, want_cnt = oracle_topk(user[pick], item, k)
    got, cnt = g.recommend_users(pick, k, mask=None)
    np.testing.assert_array_equal(got, want)
    custom = sps.csr_matrix((rng.random((pick.size, I)) < 0.2).astype(np.float32))
    lists = [np.sort(rng.choice(I, int(n), replace=False)) for n in rng.integers(0, 1500, pick.size)]
    for allowed in ([np.sort(rng.choice(I, 900, replace=False))], lists):
        for mask, omask in (("train", X[pick]), (custom, custom)):
            _, want, want_cnt = oracle_topk_allowed(user[pick], item, k, allowed, omask)
            got, cnt = g.recommend_users(pick, k, mask=mask, allowed=lists_csr(allowed))
            np.testing.assert_array_equal(cnt, want_cnt)
            np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        g.recommend_users([0, U], k)
    with pytest.raises(ValueError):
        g.recommend_users([-1], k)
    got, cnt = g.recommend_users(np.zeros(0, np.int64), k)
    assert got.shape == (0, k) and cnt.shape == (0,)
    # past the fused kernel's cutoff: the SIMT path takes gathered users too (explicit or no mask)
    _, want, want_cnt = oracle_topk(user[pick], item, 150, custom)
    got, cnt = g.recommend_users(pick, 150, mask=custom)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(NotImplementedError):
        g.recommend_users(pick, 150, mask="train")


def test_recommend_embeddings_against_the_resident_item_factors(core):
    """ials_trainer_recommend_embeddings: rows that are not users of the model (fold-in results),
    K < row stride (padding on the device), explicit masks, allow-lists, the SIMT path past 128."""
    rng = np.random.default_rng(23)
    U, I, K, k = 50, 2500, 40, 15
    user = rng.integers(-1, 2, size=(U, K)).astype(np.float32)
    item = rng.integers(-1, 2, size=(I, K)).astype(np.float32)
    X = sps.random(U, I, density=0.02, random_state=2, format="csr", dtype=np.float32)
    g = trainer(core, X, K, user, item)
    emb = rng.integers(-2, 3, size=(333, K)).astype(np.float32)
    mask = sps.csr_matrix((rng.random((333, I)) < 0.1).astype(np.float32))
    _, want, want_cnt = oracle_topk(emb, item, k, mask)
    got, cnt, sc = g.recommend_embeddings(emb, k, mask=mask, return_scores=True)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)
    lists = [np.sort(rng.choice(I, int(n), replace=False)) for n in rng.integers(0, 900, 333)]
    _, want, want_cnt = oracle_topk_allowed(emb, item, k, lists, None)
    got, cnt = g.recommend_embeddings(emb, k, allowed=lists_csr(lists))
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cnt, want_cnt)
    _, want, want_cnt = oracle_topk(emb, item, 200, mask)
    got, cnt = g.recommend_embeddings(emb, 200, mask=mask)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        g.recommend_embeddings(emb, k, mask="train")
    with pytest.raises(ValueError):
        g.recommend_embeddings(emb[:, :-1], k)
    np.testing.assert_array_equal(g.item, item)  # the model is untouched
