"""CPU-side checks of the product's host layer: the C-ABI library loads and
exports every symbol include/ials_b200.h declares, the `_ials_core` mirror has
the reference's names and defaults, the host Metrics arithmetic equals the
oracle's restatement of the reference, and there is no silent CPU fallback."""
import os
import pickle
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sps

import oracle
import irspack_b200
from irspack_b200 import _ials_core as core
from irspack_b200 import _lib
from irspack_b200.evaluation import Evaluator, Metrics
from irspack_b200.synth import holdout_split, init_factors, synth_csr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ials_b200.h")).read()
    return sorted(set(re.findall(r"IALS_API[^;(]*?\b(ials_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (ials_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    for n in names:  # and the ctypes binding knows each of them
        assert hasattr(_lib.lib, n)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md section 3 maps every exported symbol to the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols()
               if n not in doc and n.replace("ials_trainer", "") not in doc]
    assert not missing, missing


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure_not_fallback():
    if irspack_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    X = sps.csr_matrix(np.eye(4, dtype=np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.IALSTrainer(core.IALSModelConfigBuilder().build(), X)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        irspack_b200.topk_scores(np.zeros((2, 4), np.float32), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "irspack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "ials_oracle" not in text, f


def test_enums_and_builder_defaults():
    # wrapper.cpp:25-40, IALSLearningConfig.hpp:33-43, 114-120
    assert core.LossType.ORIGINAL.value == 0 and core.LossType.IALSPP.value == 1
    assert [core.SolverType.CHOLESKY.value, core.SolverType.CG.value, core.SolverType.IALSPP.value] == [0, 1, 2]
    assert core.IALSPP is core.SolverType.IALSPP and core.ORIGINAL is core.LossType.ORIGINAL
    m = core.IALSModelConfigBuilder().build()
    assert (m.K, m.random_seed, m.loss_type) == (16, 42, core.LossType.IALSPP)
    assert (m.reg, m.alpha0, m.nu, m.init_stdev) == (0.1, 0.1, 1.0, 0.1)
    s = core.IALSSolverConfigBuilder().build()
    assert (s.n_threads, s.solver_type, s.max_cg_steps) == (1, core.SolverType.CG, 3)
    assert (s.ialspp_subspace_dimension, s.ialspp_iteration) == (64, 1)
    m2 = pickle.loads(pickle.dumps(core.IALSModelConfigBuilder().set_K(7).set_reg(3.0).build()))
    assert (m2.K, m2.reg) == (7, 3.0)
    s2 = pickle.loads(pickle.dumps(core.IALSSolverConfigBuilder().set_max_cg_steps(9).build()))
    assert s2.max_cg_steps == 9
    st = m._as_struct()
    assert (st.K, st.loss_type) == (16, 1)


def test_recommender_argument_handling():
    X = sps.csr_matrix(np.eye(3))
    with pytest.raises(ValueError, match="IALSPP"):  # ials.py:430-433
        irspack_b200.IALSRecommender(X, user_features=np.zeros((3, 2)), solver_type="IALSPP")
    rec = irspack_b200.IALSRecommender(X, n_components=4, nu=0.5, nu_star=1.0, alpha0=0.3, reg=2.0)
    from irspack_b200.ials import compute_reg_scale
    assert rec.scaled_reg == pytest.approx(
        2.0 * compute_reg_scale(rec.X_train_all, 0.3, 1.0) / compute_reg_scale(rec.X_train_all, 0.3, 0.5))
    with pytest.raises(RuntimeError):
        rec.trainer_as_ials
    Xs = irspack_b200.IALSRecommender._scale_X(sps.csr_matrix(np.array([[0.0, 3.0]])),
                                               rec.confidence_scaling.__class__["log"], 3.0)
    assert Xs.data[0] == pytest.approx(np.log(2.0))


def test_metrics_host_arithmetic_equals_oracle():
    """Feed the product's Metrics the oracle's own top lists: every metric must
    come out identical (Metrics::update / as_dict, evaluator.cpp:87-166)."""
    rns = np.random.RandomState(7)
    for U, I, C, rwc in [(40, 30, 7, False), (25, 12, 12, True), (10, 5, 3, False)]:
        scores = rns.randn(U, I).astype(np.float32)
        scores[rns.rand(U, I) < 0.2] = -np.inf
        scores[3] = -np.inf  # a user with nothing recommendable
        gt = sps.csr_matrix((rns.rand(U, I) > 0.8).astype(np.float64))
        ref, rec, cnt = oracle.topk_metrics(scores, gt, C, recall_with_cutoff=rwc)
        m = Metrics(I)
        m.update_block(rec, np.maximum(cnt, 0), gt, rwc)
        d, r = m.as_dict(), ref.as_dict()
        assert d.keys() == r.keys()
        for k in r:
            assert d[k] == pytest.approx(r[k], rel=1e-12, abs=1e-12), k
        # chunked accumulation + merge == one shot (evaluator.py:417-439)
        parts = Metrics(I)
        for b in range(0, U, 9):
            e = min(b + 9, U)
            mm = Metrics(I)
            mm.update_block(rec[b:e], np.maximum(cnt[b:e], 0), gt[b:e], rwc)
            parts.merge(mm)
        for k in r:
            assert parts.as_dict()[k] == pytest.approx(r[k], rel=1e-12, abs=1e-12), k


def test_evaluator_argument_checks():
    gt = sps.csr_matrix(np.eye(4))
    assert Evaluator(gt, recommendable_items=[0, 1]).n_recommendable_items == 2
    with pytest.raises(ValueError):  # evaluator.py:130-133
        Evaluator(gt, per_user_recommendable_items=[[0]])
    with pytest.raises(ValueError):
        Evaluator(gt, masked_interactions=sps.csr_matrix((5, 4)))

    class Fake:
        n_users, n_items = 3, 4

    with pytest.raises(ValueError):
        Evaluator(gt).get_score(Fake())


def test_synth_is_deterministic_and_well_formed():
    a = synth_csr(300, 200, 5000, seed=11)
    b = synth_csr(300, 200, 5000, seed=11)
    assert a.nnz == 5000 and (a != b).nnz == 0
    assert a.has_canonical_format and a.indices.dtype == np.int32
    assert np.all(a.data == 1.0)
    c = synth_csr(300, 200, 5000, seed=11, values="counts")
    assert set(np.unique(c.data)) <= {1.0, 2.0, 3.0, 4.0, 5.0}
    tr, te = holdout_split(a, 0.2, 5)
    assert tr.nnz + te.nnz == a.nnz and tr.multiply(te).nnz == 0
    f = init_factors(10, 8, 3)
    assert f.dtype == np.float32 and f.shape == (10, 8)
