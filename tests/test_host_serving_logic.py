"""Host-side logic of the fused allow-list / serving paths, without a GPU: what the Evaluator and
the ID mappers hand to a recommender's fused entry points (`recommend_block`,
`recommend_cold_block`, `recommend_users`), and how they fall back when those answer
NotImplementedError.  The fake recommender below selects with numpy, in the reference's order
(descending score, ties to the smaller index, -inf never; evaluator.cpp:324-355)."""
import numpy as np
import pytest
import scipy.sparse as sps

from irspack_b200 import evaluation
from irspack_b200.evaluation import Evaluator, EvaluatorWithColdUser, _canonical_lists
from irspack_b200.id_mapping import IDMapper


def numpy_topk(scores, k, mask, allowed):
    s = np.array(scores, dtype=np.float64, copy=True)
    rows, n = s.shape
    if mask is not None:
        s[sps.csr_matrix(mask).nonzero()] = -np.inf
    if allowed is not None:
        n_lists, indptr, idx = allowed
        assert idx.dtype == np.int32 and indptr.dtype == np.int64
        ok = np.zeros((rows, n), bool)
        for r in range(rows):
            l = 0 if n_lists == 1 else r
            seg = idx[indptr[l]: indptr[l + 1]]
            assert (np.diff(seg) > 0).all() and (seg >= 0).all() and (seg < n).all()  # canonical
            ok[r, seg] = True
        s[~ok] = -np.inf
    out = np.full((rows, k), -1, np.int32)
    cnt = np.zeros(rows, np.int32)
    val = np.zeros((rows, k), np.float32)
    for r in range(rows):
        order = np.lexsort((np.arange(n), -s[r]))
        order = order[np.isfinite(s[r, order])][:k]
        out[r, : order.size] = order
        val[r, : order.size] = s[r, order]
        cnt[r] = order.size
    return out, cnt, val


class FakeFused:
    """Scores = a fixed matrix; the fused entry points record their arguments."""

    def __init__(self, scores, X, fail=False):
        self.scores, self.X_train_all = scores, sps.csr_matrix(X)
        self.n_users, self.n_items = scores.shape
        self.calls, self.fail = [], fail

    def recommend_block(self, begin, end, cutoff, mask="train", allowed=None):
        self.calls.append(("block", begin, end, cutoff, allowed is not None))
        if self.fail and allowed is not None:
            raise NotImplementedError("not here")
        m = self.X_train_all[begin:end] if isinstance(mask, str) else mask
        return numpy_topk(self.scores[begin:end], cutoff, m, allowed)[:2]

    def recommend_users(self, u, cutoff, mask="train", allowed=None, return_scores=True):
        self.calls.append(("users", len(u), cutoff, allowed is not None, isinstance(mask, str)))
        if self.fail:
            raise NotImplementedError("not here")
        m = self.X_train_all[u] if isinstance(mask, str) else mask
        return numpy_topk(self.scores[u], cutoff, m, allowed)

    def get_score_block(self, begin, end):
        return self.scores[begin:end].astype(np.float32)

    def get_score_remove_seen(self, u):
        s = self.scores[u].astype(np.float32)
        s[self.X_train_all[u].nonzero()] = -np.inf
        return s


def test_canonical_lists():
    indptr = np.array([0, 5, 5, 8], np.int64)
    flat = np.array([7, 3, 3, -1, 40, 2, 2, 0], np.int64)
    ip, ix = _canonical_lists(indptr, flat, 10)
    assert ip.tolist() == [0, 2, 2, 4] and ix.tolist() == [3, 7, 0, 2] and ix.dtype == np.int32
    ip, ix = _canonical_lists(np.array([0], np.int64), np.zeros(0, np.int64), 10)
    assert ip.tolist() == [0] and ix.size == 0


@pytest.fixture
def data():
    rng = np.random.default_rng(4)
    U, I = 70, 90
    scores = rng.integers(0, 12, size=(U, I)).astype(np.float64)  # many ties
    X = (rng.random((U, I)) < 0.1).astype(np.float32)
    gt = sps.csr_matrix((rng.random((U, I)) < 0.08).astype(np.float32))
    return rng, U, I, scores, X, gt


def test_evaluator_hands_canonical_lists_to_the_fused_path_and_falls_back(data, monkeypatch):
    rng, U, I, scores, X, gt = data
    shared = [int(i) for i in rng.permutation(I)[:40]] + [3, 3]
    per_user = [[int(i) for i in rng.choice(I, int(n))] for n in rng.integers(0, 60, U)]
    host_calls = []
    orig = evaluation.select_topk

    def fake_select(scores_, cutoff, mask=None, allowed=None):  # the host path, on numpy
        host_calls.append(allowed is not None)
        a = None
        if allowed is not None:
            ip, ix = _canonical_lists(np.asarray(allowed[1], np.int64), np.asarray(allowed[2], np.int64), I)
            a = (allowed[0], ip, ix)
        return numpy_topk(scores_, min(cutoff, I), mask, a)[:2]

    monkeypatch.setattr(evaluation, "select_topk", fake_select)
    for kw in (dict(recommendable_items=shared), dict(per_user_recommendable_items=per_user)):
        ev = Evaluator(gt, cutoff=7, mb_size=16, **kw)
        fused = FakeFused(scores, X)
        got = ev.get_score(fused)
        assert fused.calls and all(c[0] == "block" and c[4] for c in fused.calls) and not host_calls
        assert len(fused.calls) == 1  # one call of up to 32768 users, not mb_size blocks
        refused = FakeFused(scores, X, fail=True)
        want = ev.get_score(refused)  # NotImplementedError -> host score blocks of mb_size users
        assert len(refused.calls) == 1 and len(host_calls) == -(-U // 16) and all(host_calls)
        host_calls.clear()
        assert got == pytest.approx(want, abs=1e-12)
    # no lists: a NotImplementedError of the model is not swallowed
    class Broken(FakeFused):
        def recommend_block(self, *a, **k):
            raise NotImplementedError("broken")
    with pytest.raises(NotImplementedError):
        Evaluator(gt, cutoff=7).get_score(Broken(scores, X))
    monkeypatch.setattr(evaluation, "select_topk", orig)


def test_id_mapper_fused_serving_arguments_and_fallback(data, monkeypatch):
    rng, U, I, scores, X, gt = data
    from irspack_b200 import id_mapping

    def fake_retrieve(score, allowed, cutoff, n_threads=1):  # the reference's flow, on numpy
        a = None
        if len(allowed):
            ip = np.zeros(len(allowed) + 1, np.int64)
            np.cumsum([len(x) for x in allowed], out=ip[1:])
            ip, ix = _canonical_lists(ip, np.asarray([i for x in allowed for i in x], np.int64), I)
            a = (len(allowed), ip, ix)
        idx, cnt, val = numpy_topk(np.where(np.isinf(score), -np.inf, score), min(cutoff, I), None, a)
        return [[(int(i), float(v)) for i, v in zip(idx[r, :cnt[r]], val[r, :cnt[r]])] for r in range(score.shape[0])]

    monkeypatch.setattr(id_mapping, "retrieve_recommend_from_score", fake_retrieve)
    users = [f"u{i}" for i in range(U)]
    items = [f"i{j}" for j in range(I)]
    mapper = IDMapper(users, items)
    picked = ["u5", "u0", "u69", "u5"]
    shared = [items[j] for j in rng.permutation(I)[:30]] + ["unknown item"]
    per_user = [[items[j] for j in rng.choice(I, 25)] for _ in picked]
    forbidden = [[items[j] for j in rng.choice(I, 10)] for _ in picked]
    for kw in (dict(), dict(allowed_item_ids=shared), dict(per_user_allowed_item_ids=per_user),
               dict(forbidden_item_ids=forbidden), dict(per_user_allowed_item_ids=per_user, forbidden_item_ids=forbidden)):
        fused, refused = FakeFused(scores, X), FakeFused(scores, X, fail=True)
        got = mapper.recommend_for_known_user_batch(fused, picked, cutoff=6, **kw)
        want = mapper.recommend_for_known_user_batch(refused, picked, cutoff=6, **kw)
        assert [c[0] for c in fused.calls] == ["users"] and fused.calls[0][3] == ("allowed_item_ids" in kw or "per_user_allowed_item_ids" in kw)
        # without forbidden items the kernel masks with its own copy of the training rows
        assert fused.calls[0][4] == ("forbidden_item_ids" not in kw)
        assert got == want
    one = mapper.recommend_for_known_user_id(FakeFused(scores, X), "u9", cutoff=5, forbidden_item_ids=forbidden[0])
    assert one == mapper.recommend_for_known_user_id(FakeFused(scores, X, fail=True), "u9", cutoff=5,
                                                     forbidden_item_ids=forbidden[0])
    # past the fused kernel's cutoff the reference's flow is taken without asking
    wide = np.tile(scores, (1, 4))  # 360 items: min(cutoff, n_items) = 200 > 128
    big = FakeFused(wide, np.tile(X, (1, 4)))
    monkeypatch.setattr(id_mapping, "retrieve_recommend_from_score",
                        lambda score, allowed, cutoff, n_threads=1: [[] for _ in range(score.shape[0])])
    IDMapper(users, [f"i{j}" for j in range(wide.shape[1])]).recommend_for_known_user_batch(big, picked, cutoff=200)
    assert not big.calls


def test_csr_row_block_is_a_view_equal_to_the_slice():
    from irspack_b200.evaluation import _csr_row_block

    m = sps.random(50, 30, density=0.2, random_state=1, format="csr", dtype=np.float32)
    m.sort_indices()
    for b, e in ((0, 50), (7, 19), (49, 50), (10, 10), (0, 0)):
        v = _csr_row_block(m, b, e)
        assert v.shape == (e - b, 30) and v.has_sorted_indices
        if v.nnz:  # (checked first: scipy's comparison below may canonicalise v in place)
            assert np.shares_memory(v.indices, m.indices) and np.shares_memory(v.data, m.data)
        assert (v != m[b:e]).nnz == 0
    empty = sps.csr_matrix((5, 30), dtype=np.float32)
    assert _csr_row_block(empty, 1, 4).nnz == 0
