"""Evaluator for the B200 iALS path.

Mirror of the hot-user ``Evaluator`` flow of the reference
(/root/reference/src/irspack/evaluation/evaluator.py:98-205, 400-441) with the
per-chunk work re-cut for the GPU:

* the score GEMM, the seen-item mask and the top-``cutoff`` selection run on the
  device (``IALSTrainer.recommend`` -- fused, the U x I score matrix never
  reaches the host -- or ``ials_topk_scores`` for any other recommender's
  float32 score block);
* only ``cutoff`` indices per user come back, and the metric arithmetic of
  ``Metrics::update`` / ``as_dict`` (cpp_source/evaluator.cpp:87-166) is done
  here on the host in float64, vectorised with numpy.

Not mirrored (outside the hot path, raise ``NotImplementedError``):
``recommendable_items`` / ``per_user_recommendable_items``, float64 score
blocks, ``EvaluatorWithColdUser``.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from enum import Enum, auto
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sps

from ._ials_core import _current_device_and_stream, _ptr
from ._lib import check, lib
from ._threading import get_n_threads


class TargetMetric(Enum):  # evaluator.py:17-22
    ndcg = auto()
    recall = auto()
    hit = auto()
    map = auto()
    precision = auto()


METRIC_NAMES = [  # evaluator.py:25-35
    "hit", "recall", "ndcg", "map", "precision", "gini_index", "entropy", "appeared_item",
    "catalog_coverage",
]


class Metrics:
    """Accumulator equal to the reference's ``Metrics`` (evaluator.cpp:49-179)."""

    def __init__(self, n_item: int) -> None:
        self.n_item = int(n_item)
        self.valid_user = 0
        self.total_user = 0
        self.hit = 0.0
        self.recall = 0.0
        self.ndcg = 0.0
        self.precision = 0.0
        self.map = 0.0
        self.item_cnt = np.zeros(self.n_item, dtype=np.int64)

    def merge(self, other: "Metrics") -> None:  # evaluator.cpp:76-85
        self.hit += other.hit
        self.recall += other.recall
        self.ndcg += other.ndcg
        self.total_user += other.total_user
        self.valid_user += other.valid_user
        self.item_cnt += other.item_cnt
        self.precision += other.precision
        self.map += other.map

    def update_block(self, rec: np.ndarray, n_rec: np.ndarray, gt: sps.csr_matrix,
                     recall_with_cutoff: bool) -> None:
        """``get_metrics_local`` bookkeeping + ``Metrics::update`` for a block of users
        (evaluator.cpp:308-361, 127-166).  ``rec`` int32 [rows, cutoff] (-1 padded),
        ``n_rec`` valid entries per row, ``gt`` the block's ground-truth rows."""
        rows, cutoff = rec.shape
        self.total_user += rows
        if rows == 0:
            return
        n_gt = np.diff(gt.indptr).astype(np.int64)
        valid = n_gt > 0  # users with empty ground truth are skipped, :319-321
        self.valid_user += int(valid.sum())
        use = valid & (n_rec > 0)  # Metrics::update returns early when nothing is recommendable
        if not use.any():
            return
        rec = rec[use].astype(np.int64)
        n_rec = n_rec[use].astype(np.int64)
        n_gt = n_gt[use]
        r = rec.shape[0]
        pos = np.arange(cutoff)
        in_list = pos[None, :] < n_rec[:, None]
        # membership of every recommended (user, item) pair in the ground truth
        gt_use = gt[np.flatnonzero(use)]
        gt_keys = (np.repeat(np.arange(r, dtype=np.int64), np.diff(gt_use.indptr)) * self.n_item
                   + gt_use.indices.astype(np.int64))
        rec_keys = np.arange(r, dtype=np.int64)[:, None] * self.n_item + np.where(in_list, rec, 0)
        hits = np.isin(rec_keys, gt_keys) & in_list
        discount = 1.0 / np.log2(2.0 + pos)  # prepare_dcg_discount, :42-48
        cum_discount = np.cumsum(discount)
        dcg = (hits * discount[None, :]).sum(axis=1)
        idcg = cum_discount[np.minimum(n_gt, n_rec) - 1]
        cum_hit = np.cumsum(hits, axis=1)
        ap = (hits * (cum_hit / (pos[None, :] + 1.0))).sum(axis=1)
        total_hit = cum_hit[:, -1]
        self.hit += float((total_hit > 0).sum())
        self.precision += float((total_hit / n_rec).sum())
        denom = np.minimum(n_gt, n_rec) if recall_with_cutoff else n_gt
        self.recall += float((total_hit / denom).sum())
        self.ndcg += float((dcg / idcg).sum())
        self.map += float((ap / n_gt).sum())
        self.item_cnt += np.bincount(rec[in_list], minlength=self.n_item)

    def as_dict(self) -> Dict[str, float]:  # evaluator.cpp:87-123
        cnt = np.sort(self.item_cnt)
        total_item = float(cnt.sum())
        nz = cnt > 0
        appeared = float(nz.sum())
        entropy = 0.0
        gini = 0.0
        n = cnt.shape[0]
        if appeared:
            p = cnt[nz] / total_item
            entropy = float((-np.log(p) * p).sum())
            idx = np.flatnonzero(nz).astype(np.float64)
            gini = float(((2 * idx - n + 1) * cnt[nz]).sum())
        if total_item > 0:
            gini /= n * total_item
        den = self.valid_user if self.valid_user > 0 else 1
        return {
            "total_user": float(self.total_user), "valid_user": float(self.valid_user),
            "n_items": float(self.n_item), "hit": self.hit / den, "ndcg": self.ndcg / den,
            "recall": self.recall / den, "map": self.map / den,
            "precision": self.precision / den, "appeared_item": appeared, "entropy": entropy,
            "gini_index": gini,
        }


def topk_scores(scores: np.ndarray, cutoff: int, mask: Optional[sps.spmatrix] = None
                ) -> Tuple[np.ndarray, np.ndarray]:
    """Device top-``cutoff`` of a host float32 score block with the reference's
    ordering (descending score, ties to the smaller index, ``-inf`` excluded)."""
    if scores.dtype == np.float64:
        raise NotImplementedError("float64 score blocks are outside the B200 hot path")
    if scores.dtype != np.float32:
        raise ValueError("score must be either float32 or float64.")  # evaluator.py:183
    scores = np.ascontiguousarray(scores)
    rows, n_items = scores.shape
    idx = np.empty((rows, cutoff), dtype=np.int32)
    cnt = np.empty((rows,), dtype=np.int32)
    mi = mx = None
    if mask is not None:
        m = sps.csr_matrix(mask, copy=True)
        m.eliminate_zeros()
        m.sort_indices()
        mi = np.ascontiguousarray(m.indptr, dtype=np.int64)
        mx = np.ascontiguousarray(m.indices, dtype=np.int32)
    dev, stream = _current_device_and_stream()
    check(lib.ials_topk_scores(_ptr(scores), rows, n_items, int(cutoff), _ptr(mi), _ptr(mx), dev,
                               ctypes.c_void_p(stream), _ptr(idx), _ptr(None), _ptr(cnt)))
    return idx, cnt


class Evaluator:
    """See the module docstring; arguments as in evaluator.py:98-161."""

    def __init__(self, ground_truth: Any, offset: int = 0, cutoff: int = 10,
                 target_metric: str = "ndcg", recommendable_items: Optional[List[int]] = None,
                 per_user_recommendable_items: Any = None, masked_interactions: Any = None,
                 n_threads: Optional[int] = None, recall_with_cutoff: bool = False,
                 mb_size: int = 4096) -> None:
        if recommendable_items is not None or per_user_recommendable_items is not None:
            raise NotImplementedError("recommendable-item lists are outside the B200 hot path")
        gt = sps.csr_matrix(ground_truth).astype(np.float64)
        gt.sort_indices()
        self.ground_truth = gt
        self.n_recommendable_items = gt.shape[1]
        self.offset = offset
        self.n_users, self.n_items = gt.shape
        self.target_metric = TargetMetric[target_metric]
        self.cutoff = cutoff
        self.target_metric_name = f"{self.target_metric.name}@{self.cutoff}"
        self.n_threads = get_n_threads(n_threads)
        # the reference defaults to 128-row chunks sized for CPU caches; the GPU
        # path only ships `cutoff` indices per user, so bigger chunks are free
        self.mb_size = mb_size
        if masked_interactions is None:
            self.masked_interactions = None
        else:
            if masked_interactions.shape != gt.shape:
                raise ValueError("ground_truth and masked_interactions have different shapes. ")
            self.masked_interactions = sps.csr_matrix(masked_interactions)
        self.recall_with_cutoff = recall_with_cutoff

    def get_target_score(self, model: Any) -> float:
        return self.get_score(model)[self.target_metric.name]

    def get_score(self, model: Any) -> Dict[str, float]:
        return self._get_scores_as_list(model, [self.cutoff])[0]

    def get_scores(self, model: Any, cutoffs: List[int]) -> Dict[str, float]:
        result: Dict[str, float] = OrderedDict()
        for cutoff, score in zip(cutoffs, self._get_scores_as_list(model, cutoffs)):
            for name in METRIC_NAMES:
                result[f"{name}@{cutoff}"] = score[name]
        return result

    def _metrics_as_dict(self, metrics: Metrics) -> Dict[str, float]:  # evaluator.py:326-334
        result = metrics.as_dict()
        result["catalog_coverage"] = (
            result["appeared_item"] / self.n_recommendable_items
            if self.n_recommendable_items else float("nan"))
        return result

    def _block_topk(self, model: Any, begin: int, end: int, cutoff: int
                    ) -> Tuple[np.ndarray, np.ndarray]:
        custom = None
        if self.masked_interactions is not None:
            custom = self.masked_interactions[begin - self.offset: end - self.offset]
        if hasattr(model, "recommend_block"):  # fused GPU path (IALSRecommender)
            return model.recommend_block(begin, end, cutoff, mask="train" if custom is None else custom)
        try:
            scores = model.get_score_block(begin, end)
        except NotImplementedError:
            scores = model.get_score(np.arange(begin, end))
        mask = model.X_train_all[begin:end] if custom is None else custom
        return topk_scores(np.asarray(scores), cutoff, mask)

    def _get_scores_as_list(self, model: Any, cutoffs: List[int]) -> List[Dict[str, float]]:
        if self.offset + self.n_users > model.n_users:  # evaluator.py:403-406
            raise ValueError("evaluator offset + n_users exceeds the model's n_users.")
        if self.n_items != model.n_items:
            raise ValueError("The model and evaluator assume different n_items.")
        for c in cutoffs:  # EvaluatorCore::get_metrics, evaluator.cpp:265-266
            if c <= 0:
                raise ValueError("cutoff must be strictly greather than 0.")
            if c > self.n_items:
                raise ValueError("cutoff must not exeeed the number of items.")
        metrics = [Metrics(self.n_items) for _ in cutoffs]
        cmax = max(cutoffs)
        block_end = self.offset + self.n_users
        for b in range(self.offset, block_end, self.mb_size):
            e = min(b + self.mb_size, block_end)
            # one top-max(cutoffs) pass serves every cutoff: a shorter list is a prefix
            rec, n_rec = self._block_topk(model, b, e, cmax)
            gt = self.ground_truth[b - self.offset: e - self.offset]
            for m, c in zip(metrics, cutoffs):
                m.update_block(rec[:, :c], np.minimum(n_rec, c), gt, self.recall_with_cutoff)
        return [self._metrics_as_dict(m) for m in metrics]
