"""Serving top-k with allow / forbid lists: ``retrieve_recommend_from_score``
(/root/reference/cpp_source/util.hpp:426-504) and ``IDMapper``
(/root/reference/src/irspack/utils/id_mapping.py).  The GPU tests follow
/root/reference/tests/utils/test_id_mapper.py with a mock recommender."""
import uuid

import numpy as np
import pytest
import scipy.sparse as sps

import oracle


def brute_force(score, allowed, cutoff):
    out = []
    for r in range(score.shape[0]):
        cand = range(score.shape[1]) if not allowed else sorted(
            {i for i in (allowed[0] if len(allowed) == 1 else allowed[r]) if 0 <= i < score.shape[1]})
        pairs = sorted(((-float(score[r, i]), i) for i in cand if score[r, i] != -np.inf))
        out.append([(i, -s) for s, i in pairs[:cutoff]])
    return out


def test_oracle_matches_brute_force():
    rng = np.random.default_rng(0)
    score = rng.standard_normal((7, 23)).astype(np.float32)
    score[2, :5] = -np.inf
    score[3, :] = -np.inf
    per_row = [list(rng.integers(-3, 30, size=rng.integers(0, 12))) for _ in range(7)]
    for allowed in ([], [[1, 5, 5, 22, 40, -1]], per_row):
        for cutoff in (0, 1, 4, 50):
            assert oracle.retrieve_recommend_from_score(score, allowed, cutoff) == brute_force(score, allowed, cutoff)
    with pytest.raises(ValueError):
        oracle.retrieve_recommend_from_score(score, [[1], [2]], 3)


def test_library_exports_retrieve_recommend():
    from irspack_b200._lib import lib

    assert lib.ials_retrieve_recommend is not None


def test_argument_errors_need_no_gpu():
    from irspack_b200.id_mapping import IDMapper, ItemIDMapper, retrieve_recommend_from_score

    s = np.zeros((3, 4), dtype=np.float32)
    with pytest.raises(ValueError):  # id_mapping.py:44-45
        retrieve_recommend_from_score(s.astype(np.float16), [], 2)
    with pytest.raises(ValueError):  # util.hpp:434
        retrieve_recommend_from_score(s, [], 2, n_threads=0)
    with pytest.raises(ValueError):  # util.hpp:436-439
        retrieve_recommend_from_score(s, [[0], [1]], 2)
    assert retrieve_recommend_from_score(s, [], 0) == [[], [], []]
    with pytest.raises(ValueError):
        ItemIDMapper(["a", "a"])
    with pytest.raises(ValueError):
        IDMapper(["u", "u"], ["a"])
    m = ItemIDMapper(["a", "b", "c"])
    X = m.list_of_user_profile_to_matrix([["a", "zzz", "c"], {"b": 2.5, "nope": 1.0}])
    assert X.shape == (2, 3)
    assert X.toarray().tolist() == [[1.0, 0.0, 1.0], [0.0, 2.5, 0.0]]
    with pytest.raises(ValueError):
        m.score_to_recommended_items_batch(np.zeros((1, 5), dtype=np.float32), 2)


class MockRecommender:
    """The reference test's mock (tests/utils/test_id_mapper.py:14-30) without its base class."""

    def __init__(self, X, scores):
        self.X = sps.csr_matrix(X)
        self.scores = scores
        self.n_users, self.n_items = X.shape

    def get_score_remove_seen(self, user_indices):
        s = np.array(self.scores[user_indices], copy=True)
        s[self.X[user_indices].nonzero()] = -np.inf
        return s

    def get_score_cold_user_remove_seen(self, X):
        s = np.exp(X.toarray())
        s /= s.sum(axis=1)[:, None]
        s[X.nonzero()] = -np.inf
        return s


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rows,n_items", [(1, 1), (5, 17), (64, 3000), (3, 40000)])
def test_retrieve_recommend_matches_oracle(rows, n_items, dtype):
    from irspack_b200.id_mapping import retrieve_recommend_from_score

    rng = np.random.default_rng(rows * 131 + n_items)
    score = rng.standard_normal((rows, n_items)).astype(dtype)
    score[rng.random(score.shape) < 0.2] = -np.inf
    if rows > 2:
        score[1, :] = -np.inf
    shared = [list(rng.integers(-5, n_items + 5, size=min(n_items, 50)))]
    per_row = [list(rng.integers(-5, n_items + 5, size=rng.integers(0, min(2 * n_items, 200)))) for _ in range(rows)]
    for allowed in ([], shared, per_row):
        for cutoff in (1, 10, n_items + 3):
            if min(cutoff, n_items) > 1024:
                with pytest.raises(NotImplementedError):
                    retrieve_recommend_from_score(score, allowed, cutoff)
                continue
            got = retrieve_recommend_from_score(score, allowed, cutoff, n_threads=2)
            want = oracle.retrieve_recommend_from_score(score, allowed, cutoff)
            assert [[i for i, _ in row] for row in got] == [[i for i, _ in row] for row in want]
            for g_row, w_row in zip(got, want):
                assert [v for _, v in g_row] == [v for _, v in w_row]  # scores are copied, not recomputed


@pytest.mark.gpu
def test_ties_come_back_in_index_order():
    from irspack_b200.id_mapping import retrieve_recommend_from_score

    score = np.zeros((2, 100), dtype=np.float32)
    score[1, 50:] = 1.0
    got = retrieve_recommend_from_score(score, [], 5)
    assert [i for i, _ in got[0]] == [0, 1, 2, 3, 4]
    assert [i for i, _ in got[1]] == [50, 51, 52, 53, 54]
    got = retrieve_recommend_from_score(score, [[99, 7, 7, 3, 60]], 4)
    assert [i for i, _ in got[0]] == [3, 7, 60, 99]
    assert [i for i, _ in got[1]] == [60, 99, 3, 7]


def check_descending(pairs):
    assert all(a[1] >= b[1] for a, b in zip(pairs, pairs[1:]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
def test_id_mapper_usecase_of_the_reference(dtype):  # tests/utils/test_id_mapper.py:44-247
    from irspack_b200.id_mapping import IDMapper

    rns = np.random.RandomState(0)
    n_users, n_items = 31, 42
    user_ids = [str(uuid.uuid4()) for _ in range(n_users)]
    item_ids = [str(uuid.uuid4()) for _ in range(n_items)]
    score = rns.randn(n_users, n_items).astype(dtype)
    X = sps.csr_matrix((score + rns.randn(*score.shape)) > 0).astype(np.float64)
    rec = MockRecommender(X, score)
    with pytest.raises(ValueError):
        IDMapper(user_ids, item_ids + [str(uuid.uuid4())]).recommend_for_known_user_id(rec, user_ids[0])
    with pytest.raises(ValueError):
        IDMapper(user_ids + [str(uuid.uuid4())], item_ids).recommend_for_known_user_id(rec, user_ids[0])
    mapper = IDMapper(user_ids, item_ids)
    with pytest.raises(RuntimeError):
        mapper.recommend_for_known_user_id(rec, str(uuid.uuid4()))

    individual = []
    for i, uid in enumerate(user_ids):
        seen = [item_ids[j] for j in X[i].nonzero()[1]]
        full = mapper.recommend_for_known_user_id(rec, uid, cutoff=n_items)
        check_descending(full)
        ids = {p[0] for p in full}
        assert not ids.intersection(seen) and len(ids.union(seen)) == n_items
        half = mapper.recommend_for_known_user_id(rec, uid, cutoff=n_items // 2)
        individual.append(half)
        check_descending(half)
        assert len(half) <= n_items // 2
        unseen = sorted(set(item_ids).difference(seen))
        forbidden = list(rns.choice(unseen, replace=False, size=len(unseen) // 2))
        restricted = mapper.recommend_for_known_user_id(rec, uid, cutoff=n_items, forbidden_item_ids=forbidden)
        check_descending(restricted)
        assert not {p[0] for p in restricted}.intersection(forbidden)
        allowed = list(rns.choice(unseen, size=min(len(unseen), n_items // 3))) + [str(uuid.uuid1())]
        with_allowed = mapper.recommend_for_known_user_id(rec, uid, cutoff=n_items, allowed_item_ids=allowed)
        check_descending(with_allowed)
        assert {p[0] for p in with_allowed}.issubset(allowed)
        assert not mapper.recommend_for_known_user_id(rec, uid, cutoff=n_items, forbidden_item_ids=forbidden,
                                                      allowed_item_ids=forbidden)
        cold = {p[0] for p in mapper.recommend_for_new_user(rec, seen, cutoff=n_items)}
        assert not cold.intersection(seen) and len(cold.union(seen)) == n_items

    if dtype == "float16":
        with pytest.raises(ValueError):
            mapper.recommend_for_known_user_batch(rec, user_ids, cutoff=n_items)
        return
    batch = mapper.recommend_for_known_user_batch(rec, user_ids, cutoff=n_items // 2)
    assert len(batch) == len(individual)
    for b_row, i_row in zip(batch, individual):
        assert [p[0] for p in b_row] == [p[0] for p in i_row]
        assert [p[1] for p in b_row] == pytest.approx([p[1] for p in i_row])

    profiles = [[item_ids[j] for j in X[i].nonzero()[1]] for i in range(n_users)]
    cold_batch = mapper.recommend_for_new_user_batch(rec, profiles, cutoff=3, n_threads=2)
    for row, seen in zip(cold_batch, profiles):
        assert len(row) == 3
        check_descending(row)
        assert not {p[0] for p in row}.intersection(seen)
    per_user_allowed = [list(rns.choice(item_ids, size=5, replace=False)) for _ in range(n_users)]
    forbidden = [a[:2] for a in per_user_allowed]
    rows = mapper.recommend_for_known_user_batch(rec, user_ids, cutoff=n_items, per_user_allowed_item_ids=per_user_allowed,
                                                 forbidden_item_ids=forbidden)
    for row, a, f, i in zip(rows, per_user_allowed, forbidden, range(n_users)):
        seen = {item_ids[j] for j in X[i].nonzero()[1]}
        assert {p[0] for p in row} == set(a) - set(f) - seen


def test_no_gpu_means_loud_failure_not_fallback():
    import irspack_b200
    from irspack_b200.id_mapping import retrieve_recommend_from_score

    if irspack_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        retrieve_recommend_from_score(np.zeros((2, 4), np.float32), [], 2)


@pytest.mark.gpu
def test_id_mapper_serves_an_ials_recommender_without_a_score_block(monkeypatch):
    """IDMapper over an IALSRecommender: the fused device path (users by index / folded-in
    profiles + seen and forbidden items as the mask + allowed items as allow-lists, one kernel)
    returns what the reference's flow returns (host score block -> -inf scatter ->
    retrieve_recommend_from_score, utils/id_mapping.py:225-453)."""
    from irspack_b200 import IALSRecommender
    from irspack_b200.id_mapping import IDMapper

    rng = np.random.default_rng(31)
    U, I, K = 150, 700, 32
    X = sps.csr_matrix((rng.random((U, I)) < 0.05).astype(np.float32))
    rec = IALSRecommender(X, n_components=K, alpha0=0.1, reg=0.05, train_epochs=1).learn()
    t = rec.trainer_as_ials.core_trainer  # small-integer factors: exact scores, massive ties
    t.user = rng.integers(-2, 3, size=(U, K)).astype(np.float32)
    t.item = rng.integers(-2, 3, size=(I, K)).astype(np.float32)
    user_ids = [f"u{i}" for i in range(U)]
    item_ids = [f"i{j}" for j in range(I)]
    mapper = IDMapper(user_ids, item_ids)

    class HostOnly:  # the same model without the fused entry points
        n_users, n_items, X_train_all = U, I, rec.X_train_all
        get_score_remove_seen = staticmethod(rec.get_score_remove_seen)
        get_score_cold_user_remove_seen = staticmethod(rec.get_score_cold_user_remove_seen)

    fused_calls = []
    orig = type(t)._recommend
    monkeypatch.setattr(type(t), "_recommend", lambda self, *a: fused_calls.append(1) or orig(self, *a))

    def same(got, want):
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert [i for i, _ in g] == [i for i, _ in w]
            np.testing.assert_allclose([v for _, v in g], [v for _, v in w], rtol=1e-5, atol=1e-5)

    picked = [user_ids[i] for i in (7, 3, 149, 3, 0, 88)]          # any order, a repeat
    shared = [item_ids[j] for j in rng.permutation(I)[:300]] + ["no such item"]
    per_user = [[item_ids[j] for j in rng.choice(I, int(n))] for n in rng.integers(0, 400, len(picked))]
    forbidden = [[item_ids[j] for j in rng.choice(I, 50)] for _ in picked]
    for kw in (dict(), dict(allowed_item_ids=shared), dict(per_user_allowed_item_ids=per_user),
               dict(forbidden_item_ids=forbidden), dict(allowed_item_ids=shared, forbidden_item_ids=forbidden),
               dict(per_user_allowed_item_ids=per_user, forbidden_item_ids=forbidden)):
        for cutoff in (1, 10, 128):
            fused_calls.clear()
            got = mapper.recommend_for_known_user_batch(rec, picked, cutoff=cutoff, **kw)
            assert fused_calls
            same(got, mapper.recommend_for_known_user_batch(HostOnly(), picked, cutoff=cutoff, **kw))
    # one user
    for kw in (dict(), dict(allowed_item_ids=shared), dict(forbidden_item_ids=forbidden[0]),
               dict(allowed_item_ids=shared, forbidden_item_ids=forbidden[0])):
        same([mapper.recommend_for_known_user_id(rec, "u42", cutoff=15, **kw)],
             [mapper.recommend_for_known_user_id(HostOnly(), "u42", cutoff=15, **kw)])
    with pytest.raises(RuntimeError):
        mapper.recommend_for_known_user_id(rec, "nobody")
    # a cutoff past the fused kernel's 128 takes the reference's flow by itself
    fused_calls.clear()
    same(mapper.recommend_for_known_user_batch(rec, picked, cutoff=200),
         mapper.recommend_for_known_user_batch(HostOnly(), picked, cutoff=200))
    assert not fused_calls
    # new users: profiles folded in on the device, their own items masked
    profiles = [[item_ids[j] for j in rng.choice(I, 12, replace=False)] for _ in range(5)]
    profiles.append({item_ids[1]: 2.0, item_ids[5]: 1.0})
    profiles.append([])
    # fold-in gives real-valued embeddings: compare the fused lists with the host flow tie-aware
    for kw in (dict(), dict(allowed_item_ids=shared),
               dict(forbidden_item_ids=[forbidden[0]] * len(profiles))):
        got = mapper.recommend_for_new_user_batch(rec, profiles, cutoff=10, **kw)
        want = mapper.recommend_for_new_user_batch(HostOnly(), profiles, cutoff=10, **kw)
        for g, w in zip(got, want):
            assert len(g) == len(w)
            np.testing.assert_allclose([v for _, v in g], [v for _, v in w], rtol=2e-5, atol=2e-5)
            assert not {i for i, _ in g} & {i for i in (kw.get("forbidden_item_ids") or [[]])[0]}
    g1 = mapper.recommend_for_new_user(rec, profiles[0], cutoff=10, allowed_item_ids=shared)
    w1 = mapper.recommend_for_new_user(HostOnly(), profiles[0], cutoff=10, allowed_item_ids=shared)
    np.testing.assert_allclose([v for _, v in g1], [v for _, v in w1], rtol=2e-5, atol=2e-5)
    assert not {i for i, _ in g1} & set(profiles[0]) and {i for i, _ in g1} <= set(shared)
