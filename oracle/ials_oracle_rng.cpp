// CPU oracle, second translation unit: the helpers whose results depend on libstdc++'s random
// number algorithms (Solver::initialize, the row-wise split).  TEST INFRASTRUCTURE like
// ials_oracle.cpp -- only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs load it.
//
// Compiled WITHOUT -march=native and with -ffp-contract=off: the reference ships portable
// Release wheels (CMakeLists.txt:11-13, x86-64 baseline, no FMA), and std::normal_distribution's
// `ret * stddev + mean` rounds differently once the compiler may contract it into an FMA
// (1 ulp on a few values; found by tests/test_gpu_parity.py::
// test_default_init_matches_libstdcxx_reference_rng, where the product's host code -- built by
// nvcc's host compiler for the x86-64 baseline -- is compared bit for bit).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>

#define ORACLE_API extern "C" __attribute__((visibility("default")))
namespace {
constexpr int STATUS_OK = 0, STATUS_INVALID = 1;
}

// ---- helpers that let the oracle reproduce the reference's docstring known-answer ----
// (src/irspack/recommenders/ials.py:345-353; tests/test_oracle_known_answer.py)

// Solver::initialize, IALSTrainer.hpp:64-76: mt19937(seed) + normal_distribution<float>
// (libstdc++'s Marsaglia polar method; the value depends on the C++ standard library).
ORACLE_API int oracle_init_factors_f32(float *factor, int64_t n, int64_t K, float init_stdev,
                                       int32_t seed) {
  if (!(init_stdev > 0)) return STATUS_OK;
  std::mt19937 gen(seed);
  std::normal_distribution<float> dist(0.0, init_stdev / std::sqrt((double)K));
  for (int64_t i = 0; i < n * K; i++) factor[i] = dist(gen);
  return STATUS_OK;
}

// SplitFunction::split_imple + SplitByRatioFunction, cpp_source/util.hpp:72-141:
// per row, std::shuffle of the positions with ONE mt19937(seed) carried across rows;
// the first floor|ceil(nnz * ratio) shuffled positions go to test.  is_test[j] is
// written for every stored element j (the caller rebuilds the two CSR matrices).
ORACLE_API int oracle_rowwise_split(const int64_t *indptr, int64_t n_rows, int64_t random_seed,
                                    double test_ratio, int ceil_n, uint8_t *is_test) {
  if (!(test_ratio <= 1.0 && test_ratio >= 0.0)) return STATUS_INVALID;
  std::mt19937 random_state(random_seed);
  std::vector<uint64_t> index_;
  for (int64_t row = 0; row < n_rows; row++) {
    index_.clear();
    const int64_t s = indptr[row], cnt = indptr[row + 1] - indptr[row];
    for (int64_t c = 0; c < cnt; c++) index_.push_back((uint64_t)c);
    std::shuffle(index_.begin(), index_.end(), random_state);
    const size_t n_test = ceil_n ? (size_t)std::ceil(cnt * test_ratio) : (size_t)std::floor(cnt * test_ratio);
    for (size_t i = 0; i < (size_t)cnt; i++) is_test[s + index_[i]] = i < n_test ? 1 : 0;
  }
  return STATUS_OK;
}
