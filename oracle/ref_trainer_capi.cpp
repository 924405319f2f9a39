// oracle/_ref, trainer half: the reference's OWN iALS trainer -- /root/reference/cpp_source/als/
// IALSTrainer.hpp with IALSLearningConfig.hpp / definitions.hpp, compiled from where they lie
// (this file only #includes them) against the Eigen stand-in of oracle/ref_shim/ -- behind C
// entry points for ctypes.  TEST INFRASTRUCTURE: it pins the oracle's restatement
// (oracle/ials_oracle.cpp) in tests/test_oracle_vs_reference_trainer.py; nothing under
// irspack_b200/ may load it, and it is not a performance baseline (the stand-in's products are
// plain loops, not Eigen's kernels).
//
// Everything above the level of a matrix product here is the reference's code: Solver::prepare_p,
// step_cg, step_cholesky + BatchedRankUpdater, step_ialspp / step_icd, compute_reg, the exits and
// failure tests, IALSTrainer::step / transform_* / compute_loss / user_scores, initialize.
#include "als/IALSTrainer.hpp"  // -I /root/reference/cpp_source

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

using namespace irspack::ials;

namespace {
thread_local std::string g_err;

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::invalid_argument &e) {
    g_err = e.what();
    return 1;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}

SolverConfig solver(int64_t n_threads, int solver_type, int64_t max_cg_steps, int64_t subspace, int64_t iters) {
  // the product's C ABI numbers the solvers like wrapper.cpp:29-32: CHOLESKY 0, CG 1, IALSPP 2
  const SolverType st = solver_type == 0 ? SolverType::Cholesky : (solver_type == 1 ? SolverType::CG : SolverType::IALSPP);
  return SolverConfig((size_t)n_threads, st, (size_t)max_cg_steps, (size_t)subspace, (size_t)iters);
}

void copy_out(const DenseMatrix &m, float *out) { std::memcpy(out, m.data(), sizeof(float) * (size_t)m.size()); }
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char *ref_trainer_last_error() { return g_err.c_str(); }

// loss_type: 0 ORIGINAL, 1 IALSPP (wrapper.cpp:25-27)
REF_API int ref_trainer_create(int64_t K, float alpha0, float reg, float nu, float init_stdev, int32_t seed,
                               int loss_type, int64_t n_users, int64_t n_items, const int64_t *indptr,
                               const int32_t *indices, const float *data, void **out) {
  return guarded([&] {
    IALSModelConfig cfg((size_t)K, alpha0, reg, nu, init_stdev, seed,
                        loss_type == 0 ? LossType::ORIGINAL : LossType::IALSPP);
    SparseMatrix X(n_users, n_items, indptr, indices, data);
    *out = new IALSTrainer(cfg, X);
  });
}
REF_API void ref_trainer_destroy(void *h) { delete static_cast<IALSTrainer *>(h); }

REF_API int ref_trainer_get(void *h, int side, float *out) {
  return guarded([&] { copy_out(side == 0 ? static_cast<IALSTrainer *>(h)->user : static_cast<IALSTrainer *>(h)->item, out); });
}
REF_API int ref_trainer_set(void *h, int side, const float *in) {
  return guarded([&] {
    DenseMatrix &m = side == 0 ? static_cast<IALSTrainer *>(h)->user : static_cast<IALSTrainer *>(h)->item;
    std::memcpy(m.data(), in, sizeof(float) * (size_t)m.size());
  });
}
REF_API int ref_trainer_step(void *h, int64_t n_threads, int solver_type, int64_t max_cg_steps, int64_t subspace,
                             int64_t iters) {
  return guarded([&] { static_cast<IALSTrainer *>(h)->step(solver(n_threads, solver_type, max_cg_steps, subspace, iters)); });
}
// Solver::prepare_p of `side`'s solver (0: users, P = alpha0 item^T item), K x K out
REF_API int ref_trainer_gram(void *h, int side, int64_t n_threads, float *out) {
  return guarded([&] {
    auto *t = static_cast<IALSTrainer *>(h);
    const SolverConfig sc = solver(n_threads, 1, 3, 64, 1);
    Solver &s = side == 0 ? t->user_solver : t->item_solver;
    s.prepare_p(side == 0 ? t->item : t->user, t->config_, sc);
    copy_out(s.P, out);
  });
}
REF_API int ref_trainer_transform(void *h, int side, int64_t n_rows, int64_t n_cols, const int64_t *indptr,
                                  const int32_t *indices, const float *data, int64_t n_threads, int solver_type,
                                  int64_t max_cg_steps, int64_t subspace, int64_t iters, float *out) {
  return guarded([&] {
    auto *t = static_cast<IALSTrainer *>(h);
    SparseMatrix X(n_rows, n_cols, indptr, indices, data);
    const SolverConfig sc = solver(n_threads, solver_type, max_cg_steps, subspace, iters);
    copy_out(side == 0 ? t->transform_user(X, sc) : t->transform_item(X, sc), out);
  });
}
REF_API int ref_trainer_compute_loss(void *h, int64_t n_threads, float *out) {
  return guarded([&] { *out = static_cast<IALSTrainer *>(h)->compute_loss(solver(n_threads, 1, 3, 64, 1)); });
}
REF_API int ref_trainer_user_scores(void *h, int64_t begin, int64_t end, int64_t n_threads, float *out) {
  return guarded([&] {
    copy_out(static_cast<IALSTrainer *>(h)->user_scores((size_t)begin, (size_t)end, solver(n_threads, 1, 3, 64, 1)), out);
  });
}

// ---- the feature-aware model (IALSTrainer(config, X, user_features, item_features), :722-743) ----
// dense row-major feature matrices (n x F; F may be 0), the reference's DenseMatrix alternative of
// FeatureMatrix; csr_* != nullptr selects the SparseMatrix alternative instead.
namespace {
FeatureMatrix make_features(int64_t rows, int64_t cols, const float *dense, const int64_t *indptr,
                            const int32_t *indices, const float *data) {
  if (indptr != nullptr) return FeatureMatrix(SparseMatrix(rows, cols, indptr, indices, data));
  DenseMatrix m(rows, cols);
  if (rows * cols) std::memcpy(m.data(), dense, sizeof(float) * (size_t)(rows * cols));
  return FeatureMatrix(m);
}
}  // namespace

REF_API int ref_trainer_create_features(int64_t K, float alpha0, float reg, float nu, float init_stdev, int32_t seed,
                                        int loss_type, float lambda_user, float lambda_item, int64_t warmup,
                                        int64_t n_users, int64_t n_items, const int64_t *indptr,
                                        const int32_t *indices, const float *data, int64_t uf_cols,
                                        const float *uf_dense, const int64_t *uf_indptr, const int32_t *uf_indices,
                                        const float *uf_data, int64_t if_cols, const float *if_dense,
                                        const int64_t *if_indptr, const int32_t *if_indices, const float *if_data,
                                        void **out) {
  return guarded([&] {
    IALSModelConfig cfg((size_t)K, alpha0, reg, nu, init_stdev, seed,
                        loss_type == 0 ? LossType::ORIGINAL : LossType::IALSPP, lambda_user, lambda_item,
                        (size_t)warmup);
    SparseMatrix X(n_users, n_items, indptr, indices, data);
    *out = new IALSTrainer(cfg, X, make_features(n_users, uf_cols, uf_dense, uf_indptr, uf_indices, uf_data),
                           make_features(n_items, if_cols, if_dense, if_indptr, if_indices, if_data));
  });
}
REF_API int64_t ref_trainer_feature_weight_rows(void *h, int side) {
  auto *t = static_cast<IALSTrainer *>(h);
  return side == 0 ? t->user_feature_weight.rows() : t->item_feature_weight.rows();
}
REF_API int ref_trainer_get_feature_weight(void *h, int side, float *out) {
  return guarded([&] {
    auto *t = static_cast<IALSTrainer *>(h);
    copy_out(side == 0 ? t->user_feature_weight : t->item_feature_weight, out);
  });
}
REF_API int ref_trainer_transform_with_feature(void *h, int side, int64_t n_rows, int64_t n_cols,
                                               const int64_t *indptr, const int32_t *indices, const float *data,
                                               int64_t f_rows, int64_t f_cols, const float *f_dense,
                                               int64_t n_threads, int solver_type, int64_t max_cg_steps, float *out) {
  return guarded([&] {
    auto *t = static_cast<IALSTrainer *>(h);
    SparseMatrix X(n_rows, n_cols, indptr, indices, data);
    const SolverConfig sc = solver(n_threads, solver_type, max_cg_steps, 64, 1);
    const FeatureMatrix F = make_features(f_rows, f_cols, f_dense, nullptr, nullptr, nullptr);
    copy_out(side == 0 ? t->transform_user_with_feature(X, F, sc) : t->transform_item_with_feature(X, F, sc), out);
  });
}
