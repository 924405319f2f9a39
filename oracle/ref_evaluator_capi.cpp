// oracle/_ref, evaluator half: the reference's OWN evaluator -- /root/reference/cpp_source/
// evaluator.cpp, compiled from where it lies (this file only #includes it) against the container
// stand-ins of oracle/ref_shim/ (Eigen and nanobind are absent from this image) -- behind a C
// entry point for ctypes.  TEST INFRASTRUCTURE: it pins oracle.topk_metrics / oracle.Metrics
// (tests/test_oracle_vs_reference_evaluator.py); nothing under irspack_b200/ may load it.
// Everything that decides a result here is the reference's code: candidate selection
// (score != -inf, recommendable lists), std::partial_sort on (-score, index), Metrics::update,
// merge and as_dict (evaluator.cpp:42-179, 292-367).
#include "evaluator.cpp"  // -I /root/reference/cpp_source

#include <cstdint>
#include <stdexcept>
#include <string>

namespace {
thread_local std::string g_err;
}

extern "C" __attribute__((visibility("default"))) const char *ref_evaluator_last_error() { return g_err.c_str(); }

// out[11] = total_user, valid_user, n_items, hit, ndcg, recall, map, precision, appeared_item,
// entropy, gini_index (Metrics::as_dict).  n_lists in {0, 1, n_users}: recommendable items
// (rec_indptr / rec_indices, CSR-like).  Returns 0, or 1 (std::invalid_argument) / 2 (other).
extern "C" __attribute__((visibility("default"))) int ref_evaluator_metrics(
    int is_f64, const void *scores, int64_t rows, int64_t n_users, int64_t n_items, const int64_t *gt_indptr,
    const int32_t *gt_indices, int64_t n_lists, const int64_t *rec_indptr, const int64_t *rec_indices,
    int64_t cutoff, int64_t offset, int64_t n_threads, int recall_with_cutoff, double *out) {
  using namespace irspack::evaluation;
  try {
    SparseMatrix X(n_users, n_items, gt_indptr, gt_indices, static_cast<const double *>(nullptr));
    std::vector<std::vector<size_t>> rec((size_t)n_lists);
    for (int64_t l = 0; l < n_lists; l++)
      for (int64_t j = rec_indptr[l]; j < rec_indptr[l + 1]; j++) rec[(size_t)l].push_back((size_t)rec_indices[j]);
    EvaluatorCore core(X, rec);
    Metrics m = is_f64 ? core.get_metrics<double>(Eigen::Ref<DenseMatrix<double>>(static_cast<const double *>(scores), rows, n_items),
                                                  (size_t)cutoff, (size_t)offset, (size_t)n_threads, recall_with_cutoff != 0)
                       : core.get_metrics<float>(Eigen::Ref<DenseMatrix<float>>(static_cast<const float *>(scores), rows, n_items),
                                                 (size_t)cutoff, (size_t)offset, (size_t)n_threads, recall_with_cutoff != 0);
    const auto d = m.as_dict();
    const char *keys[11] = {"total_user", "valid_user", "n_items", "hit", "ndcg", "recall", "map", "precision",
                            "appeared_item", "entropy", "gini_index"};
    for (int i = 0; i < 11; i++) out[i] = d.at(keys[i]);
    return 0;
  } catch (const std::invalid_argument &e) {
    g_err = e.what();
    return 1;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}
