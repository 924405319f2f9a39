"""The oracle's restatement of the Evaluator core (oracle.topk_metrics / oracle.Metrics,
oracle/ials_oracle.cpp) against the REFERENCE'S OWN evaluator: /root/reference/cpp_source/
evaluator.cpp compiled unmodified, where it lies, into oracle/_ref (oracle.build_ref; Eigen and
nanobind are replaced by the container stand-ins of oracle/ref_shim, which hold no evaluator
arithmetic).  This is what pins SURVEY.md 8 row a14 -- candidate selection, the (-score, index)
partial sort with ties and -inf, every metric formula, merge and as_dict -- to the reference's
code rather than to a reading of it.

Runs where oracle/_ref can be built (this container) or travelled prebuilt (the GPU box)."""
import numpy as np
import pytest
import scipy.sparse as sps

import oracle

KEYS = ("total_user", "valid_user", "n_items", "hit", "ndcg", "recall", "map", "precision",
        "appeared_item", "entropy", "gini_index")


@pytest.fixture(scope="module")
def ref():
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return oracle.ref_evaluator_metrics


def oracle_metrics(scores, gt, cutoff, offset=0, recall_with_cutoff=False):
    m, _, _ = oracle.topk_metrics(scores, gt, cutoff, offset, recall_with_cutoff)
    return m.as_dict()


def assert_same(a, b):
    for k in KEYS:  # sums of the same doubles in a possibly different thread order
        assert a[k] == pytest.approx(b[k], rel=1e-13, abs=1e-13), (k, a[k], b[k])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cutoff", [1, 3, 10, 57])
@pytest.mark.parametrize("recall_with_cutoff", [False, True])
def test_random_scores_with_ties_and_minus_inf(ref, dtype, cutoff, recall_with_cutoff):
    rng = np.random.default_rng(cutoff + 100 * int(recall_with_cutoff))
    U, I = 97, 57
    scores = rng.integers(-3, 4, size=(U, I)).astype(dtype)  # heavy ties
    scores[rng.random(scores.shape) < 0.25] = -np.inf
    scores[4] = -np.inf            # a user with no candidate at all
    scores[5, :] = 2.0             # every item tied
    scores[6, 1], scores[6, 2] = -0.0, 0.0
    gt = sps.random(U, I, density=0.08, random_state=3, format="csr", dtype=np.float64)
    gt.data[:] = 1.0
    gt = gt.tolil()
    gt[9, :] = 0                   # a user without ground truth: counted in total_user only
    gt = sps.csr_matrix(gt)
    gt.eliminate_zeros()
    want = ref(scores, gt, cutoff, recall_with_cutoff=recall_with_cutoff)
    got = oracle_metrics(scores, gt, cutoff, 0, recall_with_cutoff)
    assert_same(got, want)
    assert want["total_user"] == U and want["valid_user"] < U


def test_blocks_with_offset_merge_like_the_reference(ref):
    rng = np.random.default_rng(7)
    U, I, cutoff = 300, 41, 5
    scores = rng.standard_normal((U, I)).astype(np.float32)
    scores[rng.random(scores.shape) < 0.1] = -np.inf
    gt = sps.random(U, I, density=0.1, random_state=5, format="csr", dtype=np.float64)
    want_all = ref(scores, gt, cutoff, n_threads=4)
    total = oracle.Metrics(I)
    for b in range(0, U, 128):  # Evaluator._get_scores_as_list: chunks, offset, merge
        e = min(b + 128, U)
        m, _, _ = oracle.topk_metrics(scores[b:e], gt, cutoff, b)
        total.merge(m)
        part = ref(scores[b:e], gt, cutoff, offset=b)  # the same chunk through the reference
        assert_same(m.as_dict(), part)
    assert_same(total.as_dict(), want_all)


def test_float64_scores_are_compared_as_float64(ref):
    rng = np.random.default_rng(11)
    U, I = 50, 30
    base = rng.integers(0, 3, size=(U, I)).astype(np.float64)
    scores = base + rng.integers(0, 3, size=(U, I)) * 1e-12  # ties at float32, ordered at float64
    gt = sps.random(U, I, density=0.15, random_state=2, format="csr", dtype=np.float64)
    assert_same(oracle_metrics(scores, gt, 7), ref(scores, gt, 7))


def test_reference_argument_errors(ref):
    gt = sps.csr_matrix(np.eye(4))
    s = np.zeros((4, 4), np.float32)
    with pytest.raises(ValueError, match="cutoff must be strictly"):
        ref(s, gt, 0)
    with pytest.raises(ValueError, match="cutoff must not exeeed"):
        ref(s, gt, 5)
    with pytest.raises(ValueError, match="offset"):
        ref(s, gt, 2, offset=4)
