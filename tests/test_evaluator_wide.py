"""The callers on the far side of the scoring path ("next" rows): the
``Evaluator`` features beyond the hot-user loop -- recommendable-item lists, score-matrix /
score-chunk entry points, float64 score blocks, ``EvaluatorWithColdUser``.

The cases restate /root/reference/tests/evaluation/test_evaluator.py (:87-152 cutoff metrics,
:155-228 cold users, :231-244 shape checks, :247-275 score matrix, :358-368 ``-inf``,
:371-438 chunks) against a pure-Python restatement of ``EvaluatorCore::get_metrics_local``
(/root/reference/cpp_source/evaluator.cpp:292-367) written here.

Every case runs twice: ``fake`` replaces the two device selections by numpy (host logic
only, runs without a GPU) and ``gpu`` goes through the C ABI on the B200.
"""
import math
import pickle

import numpy as np
import pytest
import scipy.sparse as sps

import oracle
from irspack_b200 import evaluation
from irspack_b200.evaluation import Evaluator, EvaluatorWithColdUser, select_topk

KEYS = ["hit", "recall", "ndcg", "map", "precision", "gini_index", "entropy", "appeared_item",
        "catalog_coverage"]


# ---------------------------------------------------------------------------------------
# numpy stand-ins for the two device entry points (host-logic runs only)
# ---------------------------------------------------------------------------------------
def _fake_topk(s32, cutoff, mask):
    s = np.array(s32, dtype=np.float32, copy=True)
    rows = s.shape[0]
    if mask is not None:
        r = np.repeat(np.arange(rows), np.diff(mask[0]))
        s[r, mask[1]] = -np.inf
    lists = oracle.retrieve_recommend_from_score(s, [], cutoff)
    return _pack(lists, rows, cutoff)


def _fake_retrieve(s32, cutoff, n_lists, indptr, flat):
    allowed = [list(flat[indptr[i]: indptr[i + 1]]) for i in range(n_lists)]
    lists = oracle.retrieve_recommend_from_score(np.asarray(s32), allowed, cutoff)
    return _pack(lists, s32.shape[0], cutoff)


def _pack(lists, rows, cutoff):
    idx = np.full((rows, cutoff), -1, np.int32)
    val = np.full((rows, cutoff), -np.inf, np.float32)
    cnt = np.zeros(rows, np.int32)
    for r, lst in enumerate(lists):
        cnt[r] = len(lst)
        for j, (i, v) in enumerate(lst):
            idx[r, j], val[r, j] = i, v
    return idx, val, cnt


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    if request.param == "fake":
        monkeypatch.setattr(evaluation, "_device_topk", _fake_topk)
        monkeypatch.setattr(evaluation, "_device_retrieve", _fake_retrieve)
    return request.param


# ---------------------------------------------------------------------------------------
# pure-Python restatement of get_metrics_local + Metrics (evaluator.cpp:87-166, 292-367)
# ---------------------------------------------------------------------------------------
def brute_metrics(scores, gt, cutoff, mask=None, allowed=None, recall_with_cutoff=False,
                  n_recommendable=None):
    scores = np.array(scores, copy=True)
    gt = sps.csr_matrix(gt)
    U, I = gt.shape
    if mask is not None:
        scores[sps.csr_matrix(mask).nonzero()] = -np.inf
    acc = dict(hit=0.0, recall=0.0, ndcg=0.0, map=0.0, precision=0.0)
    valid = 0
    item_cnt = np.zeros(I, np.int64)
    for u in range(U):
        g = set(gt[u].indices.tolist())
        if not g:
            continue
        if allowed is None or len(allowed) == 0:
            cand = range(I)
        else:
            cand = sorted(set(allowed[0] if len(allowed) == 1 else allowed[u]))
        pairs = sorted((-scores[u, i], i) for i in cand if scores[u, i] != -np.inf)
        rec = [i for _, i in pairs[:cutoff]]
        valid += 1
        if not rec:
            continue
        hits, dcg, ap = 0, 0.0, 0.0
        for rank, i in enumerate(rec):
            item_cnt[i] += 1
            if i in g:
                hits += 1
                dcg += 1.0 / math.log2(2 + rank)
                ap += hits / (rank + 1.0)
        idcg = sum(1.0 / math.log2(2 + r) for r in range(min(len(g), len(rec))))
        acc["ndcg"] += dcg / idcg
        acc["map"] += ap / len(g)
        acc["precision"] += hits / len(rec)
        acc["recall"] += hits / (min(len(g), len(rec)) if recall_with_cutoff else len(g))
        acc["hit"] += float(hits > 0)
    out = {k: v / max(valid, 1) for k, v in acc.items()}
    cnt = np.sort(item_cnt)
    total = float(cnt.sum())
    out["appeared_item"] = float((cnt > 0).sum())
    out["entropy"] = float(sum(-math.log(c / total) * (c / total) for c in cnt if c > 0))
    gini = sum((2 * i - I + 1) * float(c) for i, c in enumerate(cnt) if c > 0)
    out["gini_index"] = gini / (I * total) if total > 0 else 0.0
    nrec = I if n_recommendable is None else n_recommendable
    out["catalog_coverage"] = out["appeared_item"] / nrec if nrec else float("nan")
    return out


def assert_same(got, want, keys=KEYS, prefix=""):
    for k in keys:
        assert got[k + prefix] == pytest.approx(want[k], abs=1e-12), k


class MockRecommender:  # /root/reference/tests/mock_recommender.py
    def __init__(self, X_all, scores):
        assert X_all.shape == scores.shape
        self.X_train_all = sps.csr_matrix(X_all)
        self.n_users, self.n_items = X_all.shape
        self.scores = scores

    def get_score(self, user_indices):
        return self.scores[user_indices]

    def get_score_block(self, begin, end):
        raise NotImplementedError("get_score_block not implemented!")  # base.py:296-306

    def get_score_cold_user(self, X):  # a "model": the score only depends on the input row
        X = sps.csr_matrix(X)
        return np.asarray(X @ self.item_sim, dtype=self.scores.dtype)


def _problem(U, I, dtype, seed=42, gt_density=0.3):
    rns = np.random.RandomState(seed)
    scores = rns.randn(U, I).astype(dtype)
    gt = sps.csr_matrix((rns.rand(U, I) >= 1 - gt_density).astype(np.float64))
    seen = sps.csr_matrix((rns.rand(U, I) >= 0.8).astype(np.float64))
    return rns, scores, gt, seen


# ---------------------------------------------------------------------------------------
# hot users, any recommender's score block
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("U, I, C, dtype", [(10, 5, 5, "float32"), (10, 30, 29, "float64"),
                                            (300, 40, 7, "float32")])
def test_metrics_with_cutoff(device, U, I, C, dtype):  # test_evaluator.py:87-152
    _, scores, gt, seen = _problem(U, I, dtype, gt_density=0.7)
    rec = MockRecommender(seen, scores)
    for rwc in (False, True):
        ev = Evaluator(gt, cutoff=C, n_threads=2, recall_with_cutoff=rwc)
        fine = Evaluator(gt, cutoff=C, n_threads=2, recall_with_cutoff=rwc, mb_size=1)
        want = brute_metrics(scores, gt, C, mask=seen, recall_with_cutoff=rwc)
        assert_same(ev.get_score(rec), want)
        assert_same(fine.get_score(rec), want)
    multi = Evaluator(gt, cutoff=C).get_scores(rec, [1, C])
    assert_same(multi, brute_metrics(scores, gt, 1, mask=seen), prefix="@1")
    assert_same(multi, brute_metrics(scores, gt, C, mask=seen), prefix=f"@{C}")
    with pytest.raises(ValueError):  # float16 is not a score type (:33-37)
        Evaluator(gt, cutoff=C).get_score(MockRecommender(seen, scores.astype(np.float16)))
    with pytest.raises(ValueError):  # :66-76
        Evaluator(gt, cutoff=C, masked_interactions=sps.csr_matrix((U + 1, I)))


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_recommendable_items_shared_and_per_user(device, dtype):
    U, I, C = 37, 23, 5
    rns, scores, gt, seen = _problem(U, I, dtype)
    rec = MockRecommender(seen, scores)
    shared = [int(i) for i in rns.choice(I, 9, replace=False)]
    ev = Evaluator(gt, cutoff=C, recommendable_items=shared, mb_size=8)
    assert ev.n_recommendable_items == 9
    assert_same(ev.get_score(rec), brute_metrics(scores, gt, C, mask=seen, allowed=[shared],
                                                 n_recommendable=9))
    per_user = [[int(i) for i in rns.choice(I, rns.randint(0, 12), replace=False)] for _ in range(U)]
    n_union = len({i for l in per_user for i in l})
    want = brute_metrics(scores, gt, C, mask=seen, allowed=per_user, n_recommendable=n_union)
    for mb in (1, 8, 4096):  # 1: every block is a single row (the C ABI's "shared list" case)
        ev = Evaluator(gt, cutoff=C, per_user_recommendable_items=per_user, mb_size=mb)
        assert ev.n_recommendable_items == n_union
        assert_same(ev.get_score(rec), want)
    as_matrix = sps.lil_matrix((U, I))
    for u, l in enumerate(per_user):
        for i in l:
            as_matrix[u, i] = 1.0
    ev = Evaluator(gt, cutoff=C, per_user_recommendable_items=sps.csr_matrix(as_matrix))
    assert_same(ev.get_score(rec), want)
    # masked_interactions replaces the training matrix as the mask (evaluator.py:426-431)
    other = sps.csr_matrix((rns.rand(U, I) >= 0.5).astype(np.float64))
    ev = Evaluator(gt, cutoff=C, per_user_recommendable_items=per_user, masked_interactions=other)
    assert_same(ev.get_score(rec), brute_metrics(scores, gt, C, mask=other, allowed=per_user,
                                                 n_recommendable=n_union))
    with pytest.raises(ValueError, match="inconsistent shapes"):  # evaluator.py:130-133
        Evaluator(gt, per_user_recommendable_items=per_user[:-1])


def test_offset_block_with_allow_lists(device):
    U, I, C, off = 40, 17, 4, 13
    rns, scores, gt, seen = _problem(U, I, "float32")
    rec = MockRecommender(seen, scores)
    gt_tail = gt[off:]
    per_user = [[int(i) for i in rns.choice(I, 6, replace=False)] for _ in range(U - off)]
    ev = Evaluator(gt_tail, offset=off, cutoff=C, per_user_recommendable_items=per_user, mb_size=7)
    want = brute_metrics(scores[off:], gt_tail, C, mask=seen[off:], allowed=per_user,
                         n_recommendable=len({i for l in per_user for i in l}))
    assert_same(ev.get_score(rec), want)


def test_recommender_check(device):  # test_evaluator.py:231-244
    U, I, C = 10, 30, 29
    _, scores, gt, _ = _problem(U, I, "float64")
    ev = Evaluator(gt, cutoff=C, n_threads=2)
    with pytest.raises(ValueError):
        ev.get_score(MockRecommender(sps.csr_matrix((U - 1, I)), scores[1:]))
    with pytest.raises(ValueError):
        ev.get_score(MockRecommender(sps.csr_matrix((U, I - 1)), scores[:, 1:]))
    ev.get_score(MockRecommender(sps.csr_matrix((U, I)), scores))
    with pytest.raises(ValueError):  # evaluator.cpp:265-266
        ev.get_scores(MockRecommender(sps.csr_matrix((U, I)), scores), [0])
    with pytest.raises(ValueError):
        ev.get_scores(MockRecommender(sps.csr_matrix((U, I)), scores), [I + 1])


# ---------------------------------------------------------------------------------------
# float64 blocks: float32-resolution selection + exact re-ranking of the ties
# ---------------------------------------------------------------------------------------
def test_float64_scores_that_tie_at_float32_resolution(device):
    rns = np.random.RandomState(5)
    U, I, C = 24, 50, 6
    base = rns.randint(1, 7, size=(U, I)).astype(np.float64)  # many float32-level ties ...
    scores = base + rns.rand(U, I) * 1e-12                    # ... broken only in float64
    assert (scores.astype(np.float32) == base.astype(np.float32)).all()
    gt = sps.csr_matrix((rns.rand(U, I) >= 0.6).astype(np.float64))
    seen = sps.csr_matrix((rns.rand(U, I) >= 0.8).astype(np.float64))
    idx, cnt = select_topk(scores, C, seen)
    masked = scores.copy()
    masked[seen.nonzero()] = -np.inf
    for u in range(U):
        want = [i for _, i in sorted((-masked[u, i], i) for i in range(I) if masked[u, i] != -np.inf)][:C]
        assert idx[u, :cnt[u]].tolist() == want
    shared = [int(i) for i in rns.choice(I, 20, replace=False)]
    ev = Evaluator(gt, cutoff=C, recommendable_items=shared)
    assert_same(ev.get_score(MockRecommender(seen, scores)),
                brute_metrics(scores, gt, C, mask=seen, allowed=[shared], n_recommendable=20))
    per_user = [[int(i) for i in rns.choice(I, 15, replace=False)] for _ in range(U)]
    ev = Evaluator(gt, cutoff=C, per_user_recommendable_items=per_user, mb_size=5)
    assert_same(ev.get_score(MockRecommender(seen, scores)),
                brute_metrics(scores, gt, C, mask=seen, allowed=per_user,
                              n_recommendable=len({i for l in per_user for i in l})))


def test_retrieve_recommend_float64_is_exact(device):
    """irspack_b200.id_mapping.retrieve_recommend_from_score on float64 blocks: the same
    device selection + tie re-ranking, against the float64 oracle (util.hpp:426-504)."""
    from irspack_b200.id_mapping import retrieve_recommend_from_score

    rns = np.random.RandomState(9)
    rows, n_items = 17, 40
    score = rns.randint(1, 5, size=(rows, n_items)).astype(np.float64) + rns.rand(rows, n_items) * 1e-12
    score[rns.rand(rows, n_items) < 0.15] = -np.inf
    score[3, :] = -np.inf
    shared = [[int(i) for i in rns.randint(-3, n_items + 3, size=25)]]
    per_row = [[int(i) for i in rns.randint(-3, n_items + 3, size=rns.randint(0, 60))] for _ in range(rows)]
    for allowed in ([], shared, per_row):
        for cutoff in (1, 7, n_items + 2):
            got = retrieve_recommend_from_score(score, allowed, cutoff)
            want = oracle.retrieve_recommend_from_score(score, allowed, cutoff)
            assert got == want
    s32 = score.astype(np.float32)
    got = retrieve_recommend_from_score(s32, per_row, 7)
    want = oracle.retrieve_recommend_from_score(s32, per_row, 7)
    assert [[i for i, _ in r] for r in got] == [[i for i, _ in r] for r in want]


# ---------------------------------------------------------------------------------------
# score-matrix / score-chunk entry points
# ---------------------------------------------------------------------------------------
def test_score_from_score_matrix(device):  # test_evaluator.py:247-261
    scores = np.array([[0.1, 0.9], [0.8, 0.2]], dtype=np.float32)
    original = scores.copy()
    gt = sps.csr_matrix([[0, 1], [1, 0]])
    mask = sps.csr_matrix([[1, 0], [0, 0]])
    ev = Evaluator(gt, cutoff=1, masked_interactions=mask, mb_size=1)
    assert ev.get_score_from_score_matrix(scores)["recall"] == 1.0
    assert ev.get_scores_from_score_matrix(scores, [1])["recall@1"] == 1.0
    np.testing.assert_array_equal(scores, original)
    with pytest.raises(ValueError, match="shape"):
        ev.get_score_from_score_matrix(scores[:, :1])
    with pytest.raises(ValueError, match="dtype"):
        ev.get_score_from_score_matrix(scores.astype(np.float16))


def test_negative_infinity_scores_are_not_recommendations(device):  # :358-368
    ev = Evaluator(sps.csr_matrix([[1, 0, 0]]), cutoff=3)
    score = ev.get_score_from_score_matrix(np.array([[1.0, -np.inf, -np.inf]], dtype=np.float64))
    assert score["recall"] == 1.0
    assert score["precision"] == 1.0
    assert score["appeared_item"] == 1.0
    assert score["catalog_coverage"] == pytest.approx(1 / 3)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_score_from_score_chunks_matches_matrix(device, dtype):  # :371-398
    rns = np.random.RandomState(0)
    U, I = 11, 7
    scores = rns.randn(U, I).astype(dtype)
    original = scores.copy()
    gt = sps.csr_matrix((rns.rand(U, I) >= 0.5).astype(np.float64))
    mask = sps.csr_matrix((rns.rand(U, I) >= 0.5).astype(np.float64))
    ev = Evaluator(gt, cutoff=3, masked_interactions=mask, mb_size=2)
    expected = ev.get_scores_from_score_matrix(scores, [1, 3])
    assert_same(expected, brute_metrics(scores, gt, 1, mask=mask), prefix="@1")
    assert_same(expected, brute_metrics(scores, gt, 3, mask=mask), prefix="@3")
    cuts = [0, 1, 1, 3, 3, 6, 10, 11]
    chunks = [scores[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if a != b]
    got = ev.get_scores_from_score_chunks(iter(chunks), [1, 3])
    for k, v in expected.items():
        assert got[k] == pytest.approx(v, abs=1e-12), k
    np.testing.assert_array_equal(scores, original)
    # an empty chunk in the stream is skipped, not an error (evaluator.py:384-385)
    with_empty = [chunks[0], scores[:0]] + chunks[1:]
    got = ev.get_scores_from_score_chunks(iter(with_empty), [1, 3])
    for k, v in expected.items():
        assert got[k] == pytest.approx(v, abs=1e-12), k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_score_from_score_chunks_errors(device, dtype):  # :414-438
    U, I = 3, 4
    ev = Evaluator(sps.csr_matrix(np.eye(U, I, dtype=np.float64)), cutoff=2)
    with pytest.raises(ValueError, match="n_items"):
        ev.get_score_from_score_chunks(iter([np.zeros((U, I - 1), dtype=dtype)]))
    with pytest.raises(ValueError, match="dtype"):
        ev.get_score_from_score_chunks(iter([np.zeros((U, I), dtype=np.float16)]))
    with pytest.raises(ValueError, match="did not cover"):
        ev.get_score_from_score_chunks(iter([np.zeros((U - 1, I), dtype=dtype)]))
    with pytest.raises(ValueError, match="more rows"):
        ev.get_score_from_score_chunks(iter([np.zeros((U + 1, I), dtype=dtype)]))
    with pytest.raises(ValueError, match="2-D ndarray"):
        ev.get_score_from_score_chunks(iter([np.zeros(I, dtype=dtype)]))


# ---------------------------------------------------------------------------------------
# cold users
# ---------------------------------------------------------------------------------------
def _cold_model(U, I, dtype, seed=3):
    rns, scores, gt, seen = _problem(U, I, dtype, seed=seed)
    m = MockRecommender(sps.csr_matrix((U, I)), scores)
    m.item_sim = rns.randn(I, I)
    return rns, m, gt, seen


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cold_user_evaluator(device, dtype):  # test_evaluator.py:155-228
    U, I, C = 19, 12, 4
    rns, model, gt, seen = _cold_model(U, I, dtype)
    cold_scores = model.get_score_cold_user(seen)
    ev = EvaluatorWithColdUser(seen, gt, cutoff=C, mb_size=5)
    want = brute_metrics(cold_scores, gt, C, mask=seen)
    assert_same(ev.get_score(model), want)
    # shuffling the cold users does not change the metrics (:217-224)
    perm = rns.permutation(U)
    assert_same(EvaluatorWithColdUser(seen[perm], gt[perm], cutoff=C).get_score(model), want)
    # masked_interactions replaces the input as the mask (:170-181); the object pickles
    other = sps.csr_matrix((rns.rand(U, I) >= 0.6).astype(np.float64))
    ev2 = EvaluatorWithColdUser(seen, gt, cutoff=C, masked_interactions=other, recall_with_cutoff=True)
    want2 = brute_metrics(cold_scores, gt, C, mask=other, recall_with_cutoff=True)
    assert_same(ev2.get_score(model), want2)
    assert_same(pickle.loads(pickle.dumps(ev2)).get_score(model), want2)
    # allow-lists
    per_user = [[int(i) for i in rns.choice(I, 5, replace=False)] for _ in range(U)]
    ev3 = EvaluatorWithColdUser(seen, gt, cutoff=C, per_user_recommendable_items=per_user, mb_size=4)
    assert_same(ev3.get_score(model),
                brute_metrics(cold_scores, gt, C, mask=seen, allowed=per_user,
                              n_recommendable=len({i for l in per_user for i in l})))
    # a user whose input is its own ground truth can never be hit (:161-163)
    vicious = EvaluatorWithColdUser(gt, gt, cutoff=1)
    assert vicious.get_score(model)["hit"] == 0.0
    with pytest.raises(ValueError, match="same number of rows"):
        EvaluatorWithColdUser(seen[:-1], gt)
    with pytest.raises(ValueError, match="training items"):
        bad = MockRecommender(sps.csr_matrix((U, I + 1)), np.zeros((U, I + 1), dtype))
        ev.get_score(bad)
    with pytest.raises(NotImplementedError):
        EvaluatorWithColdUser(seen, gt, cold_item_features=sps.csr_matrix((2, 3)))


def test_score_matrix_and_chunks_of_cold_users_mask_the_input(device):  # :264-275, :401-411
    ev = EvaluatorWithColdUser(sps.csr_matrix([[1, 0]]), sps.csr_matrix([[0, 1]]), cutoff=1)
    s = np.array([[1.0, 0.0]], dtype=np.float64)
    assert ev.get_score_from_score_matrix(s)["recall"] == 1.0
    assert ev.get_score_from_score_chunks(iter([s]))["recall"] == 1.0


# ---------------------------------------------------------------------------------------
# the iALS recommender on the device: cold-user block = fold-in + fused score/mask/top-k
# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_ials_cold_user_evaluator_matches_host_score_path():
    from irspack_b200 import IALSRecommender
    from irspack_b200.synth import holdout_split, synth_csr

    X = synth_csr(700, 300, 21000, seed=11)
    X_learn, X_test = holdout_split(X, 0.3, seed=1)
    rec = IALSRecommender(X_learn[:500], n_components=32, alpha0=0.1, reg=1e-2,
                          train_epochs=3).learn()
    cold_in, cold_gt = X_learn[500:], X_test[500:]
    C = 10
    ev = EvaluatorWithColdUser(cold_in, cold_gt, cutoff=C, mb_size=64)
    got = ev.get_score(rec)  # recommend_cold_block: nothing but `cutoff` indices leaves the device
    scores = rec.get_score_cold_user(cold_in)  # score GEMM of the folded-in embedding (ials.py:486-490)
    idx, cnt = rec.recommend_cold_block(cold_in, C)
    # the fused kernel computes the same scores in 3xTF32: lists agree except for float32 near-ties
    masked = scores.astype(np.float64)
    masked[cold_in.nonzero()] = -np.inf
    tol = 1e-5 * np.abs(scores).max()
    for u in range(cold_in.shape[0]):
        order = np.lexsort((np.arange(masked.shape[1]), -masked[u]))
        want = [i for i in order[:C] if masked[u, i] != -np.inf]
        have = idx[u, :cnt[u]].tolist()
        assert len(have) == len(want)
        for a, b in zip(have, want):
            assert a == b or abs(masked[u, a] - masked[u, b]) <= tol
    want_metrics = brute_metrics(scores, cold_gt, C, mask=cold_in)
    for k in ("ndcg", "recall", "hit", "map", "precision"):
        assert got[k] == pytest.approx(want_metrics[k], abs=2e-3), k
    # allow-lists take the host score block through ials_retrieve_recommend
    shared = list(range(0, 300, 3))
    ev2 = EvaluatorWithColdUser(cold_in, cold_gt, cutoff=C, recommendable_items=shared)
    want2 = brute_metrics(scores, cold_gt, C, mask=cold_in, allowed=[shared], n_recommendable=100)
    got2 = ev2.get_score(rec)
    for k in ("ndcg", "recall", "hit", "map", "precision"):
        assert got2[k] == pytest.approx(want2[k], abs=2e-3), k
    hot = Evaluator(X_test[:500], cutoff=C, recommendable_items=shared).get_score(rec)
    hot_scores = rec.get_score_block(0, 500)
    want3 = brute_metrics(hot_scores, X_test[:500], C, mask=X_learn[:500], allowed=[shared],
                          n_recommendable=100)
    for k in ("ndcg", "recall", "hit", "map", "precision"):
        assert hot[k] == pytest.approx(want3[k], abs=1e-9), k


# ---------------------------------------------------------------------------------------
# randomised host-logic check (numpy stand-ins for the device): float32 / float64 blocks with
# heavy ties, -inf entries, masks and allow-lists against a direct float64 selection
# ---------------------------------------------------------------------------------------
def test_select_topk_random_blocks_match_a_direct_selection(monkeypatch):
    monkeypatch.setattr(evaluation, "_device_topk", _fake_topk)
    monkeypatch.setattr(evaluation, "_device_retrieve", _fake_retrieve)
    rng = np.random.default_rng(123)
    for trial in range(120):
        rows, n_items = int(rng.integers(1, 9)), int(rng.integers(1, 40))
        cutoff = int(rng.integers(1, n_items + 3))
        dtype = np.float64 if trial % 2 else np.float32
        levels = rng.integers(1, 5, size=(rows, n_items)).astype(np.float64)
        scores = levels + (rng.random((rows, n_items)) * 1e-12 if dtype == np.float64 and trial % 4 == 1 else 0)
        scores = scores.astype(dtype)
        scores[rng.random((rows, n_items)) < 0.2] = -np.inf
        mask = sps.csr_matrix((rng.random((rows, n_items)) < 0.25).astype(np.float64)) if trial % 3 else None
        kind = trial % 5
        allowed = lists = None
        if kind == 1:
            lists = [[int(i) for i in rng.integers(-2, n_items + 2, size=rng.integers(0, 2 * n_items))]]
        elif kind == 2:
            lists = [[int(i) for i in rng.integers(-2, n_items + 2, size=rng.integers(0, 2 * n_items))]
                     for _ in range(rows)]
        if lists is not None:
            indptr = np.zeros(len(lists) + 1, np.int64)
            np.cumsum([len(l) for l in lists], out=indptr[1:])
            flat = np.asarray([i for l in lists for i in l], dtype=np.int64)
            allowed = (len(lists), indptr, flat)
        before = scores.copy()
        idx, cnt = select_topk(scores, cutoff, mask, allowed)
        np.testing.assert_array_equal(scores, before)  # the caller's block is never modified
        work = scores.astype(np.float64)
        if mask is not None:
            work[mask.nonzero()] = -np.inf
        for r in range(rows):
            cand = range(n_items)
            if lists is not None:
                src = lists[0] if len(lists) == 1 else lists[r]
                cand = sorted({i for i in src if 0 <= i < n_items})
            want = [i for _, i in sorted((-work[r, i], i) for i in cand if work[r, i] != -np.inf)][:cutoff]
            assert idx[r, :cnt[r]].tolist() == want, (trial, r)
            assert (idx[r, cnt[r]:] == -1).all() or cnt[r] == idx.shape[1]


# ---------------------------------------------------------------------------------------
# metrics of big blocks on the device (ials_metrics_accumulate) == the numpy bookkeeping
# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cutoff,recall_with_cutoff", [(10, False), (100, True), (1, False), (37, False)])
def test_device_metrics_equal_the_host_bookkeeping(cutoff, recall_with_cutoff):
    from irspack_b200.evaluation import Metrics

    rng = np.random.default_rng(cutoff)
    U, I = 70000 // max(cutoff // 10, 1) + 7000, 3000
    gt = sps.random(U, I, density=0.004, random_state=5, format="csr", dtype=np.float32)
    gt.data[:] = 1.0
    keep = np.ones(U, bool)
    keep[rng.choice(U, 50, replace=False)] = False  # users without ground truth are skipped
    gt = sps.csr_matrix(sps.diags(keep.astype(np.float32)) @ gt)
    gt.eliminate_zeros()
    gt.sort_indices()
    rec = rng.integers(0, I, size=(U, cutoff)).astype(np.int32)  # repeats inside a list are allowed
    has = np.flatnonzero(np.diff(gt.indptr) > 0)
    plant = has[rng.random(has.size) < 0.5]  # make hits likely: a ground-truth item at a random position
    rec[plant, rng.integers(0, cutoff, size=plant.size)] = gt.indices[gt.indptr[plant]]
    n_rec = rng.integers(0, cutoff + 1, size=U).astype(np.int32)
    n_rec[rng.random(U) < 0.7] = cutoff
    rec[np.arange(cutoff)[None, :] >= n_rec[:, None]] = -1
    assert U * cutoff >= Metrics.DEVICE_MIN_PAIRS
    dev = Metrics(I)
    dev.update_block(rec, n_rec, gt, recall_with_cutoff)            # device (size above the threshold)
    host = Metrics(I)
    step = max(1, (Metrics.DEVICE_MIN_PAIRS - 1) // cutoff)
    for b in range(0, U, step):                                       # host arithmetic, block by block
        host.update_block(rec[b: b + step], n_rec[b: b + step], gt[b: b + step], recall_with_cutoff)
    assert dev.total_user == host.total_user == U and dev.valid_user == host.valid_user
    np.testing.assert_array_equal(dev.item_cnt, host.item_cnt)
    for name in ("hit", "recall", "ndcg", "map", "precision"):
        assert getattr(dev, name) == pytest.approx(getattr(host, name), rel=1e-12), name
    assert host.hit > 100  # the data does exercise the hit paths
