// =============================================================================
// oracle/ials_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17, no Eigen, no CUDA) of the iALS hot path of
// tohtsky/irspack.  It exists only as (a) the parity checker for the sm_100a
// CUDA path in irspack_b200/ and (b) the timed host-CPU baseline of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product never calls it.
//
// PARITY STATUS: pinned to the reference's own sources at the algorithm level, unpinned at bit
// level.  The reference's arithmetic kernels live in Eigen 5.0.1 (CMakeLists.txt:16-25, fetched
// at build time, absent from /root/reference and from this machine, no network), whose blocked
// SIMD summation order cannot be reproduced, and the reference holds no golden vectors for
// this path.  What IS compiled from /root/reference, unmodified and where it lies, is the code
// above those kernels: oracle/_ref/libref_trainer.so (als/IALSTrainer.hpp + its config headers)
// and libref_evaluator.so (evaluator.cpp), built against the container / plain-loop stand-ins of
// oracle/ref_shim (oracle.build_ref*, recipe in oracle/__init__.py and oracle/Makefile).  This
// restatement is checked against them in tests/test_oracle_vs_reference_trainer.py (CG,
// Cholesky, iALS++ / iCD, fold-in, loss, scores, Solver::initialize bit for bit, error
// behaviour; float32 summation-order tolerance 5e-5 after three epochs) and
// tests/test_oracle_vs_reference_evaluator.py (metrics exact), and, at tolerance level, against
// the reference's own closed-form tests restated in tests/test_oracle_reference_invariants.py
// (tests/recommenders/test_ials.py:54-76, 456-570, 627-697;
//  tests/evaluation/test_evaluator.py:19-152, 358-368).
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference).  Templates are instantiated for float (the reference's
// `Real`, cpp_source/als/definitions.hpp:6) and double (an f64 twin used to
// arbitrate near-ties and size the f32 tolerance).
// =============================================================================
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

enum : int { LOSS_ORIGINAL = 0, LOSS_IALSPP = 1 };  // IALSLearningConfig.hpp:11
enum : int { STATUS_OK = 0, STATUS_INVALID = 1, STATUS_CG_SINGULAR = 2,
             STATUS_CHOL_DECOMP = 3, STATUS_CHOL_SOLVE = 4 };

template <typename F> void run_workers(int n_threads, F &&fn) {
  // The reference spawns n_threads std::async workers per call and joins them
  // (IALSTrainer.hpp:87-112, 180-270).  Same here, with std::thread.
  if (n_threads <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  th.reserve(n_threads);
  for (int t = 0; t < n_threads; t++) th.emplace_back([&fn, t]() { fn(t); });
  for (auto &t : th) t.join();
}

// ---------------------------------------------------------------------------
// Gram: P = alpha0 * Y^T Y          (IALSTrainer.hpp:78-115, Solver::prepare_p)
// 16-row blocks handed out by an atomic cursor, one local accumulator per
// worker, partials summed in worker order, alpha0 applied after the sum.
// ---------------------------------------------------------------------------
template <typename Real>
int gram(const Real *Y, int64_t n, int64_t K, Real alpha0, int n_threads, Real *P) {
  if (n_threads <= 0) return STATUS_INVALID;  // :81-83
  const int64_t mb = 16;                       // :84
  std::vector<std::vector<Real>> partial(n_threads, std::vector<Real>(K * K, Real(0)));
  std::atomic<int64_t> cursor{0};
  run_workers(n_threads, [&](int tid) {
    Real *Pl = partial[tid].data();
    while (true) {
      int64_t b = cursor.fetch_add(mb);
      if (b >= n) break;
      int64_t e = std::min(b + mb, n);
      for (int64_t r = b; r < e; r++) {
        const Real *y = Y + r * K;
        for (int64_t a = 0; a < K; a++) {
          const Real ya = y[a];
          Real *row = Pl + a * K;
          for (int64_t c = 0; c < K; c++) row[c] += ya * y[c];
        }
      }
    }
  });
  for (int64_t i = 0; i < K * K; i++) P[i] = Real(0);
  for (int t = 0; t < n_threads; t++)
    for (int64_t i = 0; i < K * K; i++) P[i] += partial[t][i];
  for (int64_t i = 0; i < K * K; i++) P[i] *= alpha0;  // :113
  return STATUS_OK;
}

// reg_u = reg * pow(alpha0 * n_other + nnz, nu)  in Real   (IALSTrainer.hpp:117-120)
template <typename Real>
inline Real compute_reg(int64_t nnz, int64_t other_size, Real alpha0, Real reg, Real nu) {
  return reg * std::pow(alpha0 * other_size + nnz, nu);
}

// Dot product with 16 independent partial sums (so the compiler can keep it in
// SIMD registers, as Eigen's packet reductions do) followed by a tree reduce.
template <typename Real>
inline Real dot(const Real *a, const Real *b, int64_t K) {
  constexpr int W = 16;
  Real acc[W];
  for (int w = 0; w < W; w++) acc[w] = 0;
  int64_t k = 0;
  for (; k + W <= K; k += W)
    for (int w = 0; w < W; w++) acc[w] += a[k + w] * b[k + w];
  Real tail = 0;
  for (; k < K; k++) tail += a[k] * b[k];
  for (int h = W / 2; h > 0; h /= 2)
    for (int w = 0; w < h; w++) acc[w] += acc[w + h];
  return acc[0] + tail;
}

// ---------------------------------------------------------------------------
// CG row solve              (IALSTrainer.hpp:170-271, Solver::step_cg)
// prior (n_rows x K, may be null): the feature-aware variant adds reg_u * prior_u to b and
// solves rows without interactions too (:207-215, called through step_with_prior :634-662).
// ---------------------------------------------------------------------------
template <typename Real>
int step_cg(Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,
            const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,
            Real alpha0, Real reg, Real nu, int loss_type, int max_cg_steps, int n_threads,
            const Real *prior = nullptr) {
  if (n_threads <= 0) return STATUS_INVALID;
  std::atomic<int64_t> cursor{0};
  std::atomic<int> status{STATUS_OK};
  run_workers(n_threads, [&](int) {
    std::vector<Real> b(K), x(K), r(K), p(K), Ap(K);
    const Real bias = loss_type == LOSS_IALSPP ? Real(0) : alpha0;  // :190-191
    while (true) {
      int64_t u = cursor.fetch_add(1);
      if (u >= n_rows) break;
      if (status.load(std::memory_order_relaxed) != STATUS_OK) break;
      Real *xu = target + u * K;
      for (int64_t k = 0; k < K; k++) x[k] = xu[k];  // warm start :199
      const int64_t s = indptr[u], e = indptr[u + 1], nnz = e - s;
      const Real reg_u = compute_reg<Real>(nnz, n_other, alpha0, reg, nu);  // :202-206
      if (!prior && nnz == 0) {  // :207-210
        for (int64_t k = 0; k < K; k++) xu[k] = 0;
        continue;
      }
      if (prior) {  // :212-214
        for (int64_t k = 0; k < K; k++) b[k] = reg_u * prior[u * K + k];
      } else {
        for (int64_t k = 0; k < K; k++) b[k] = 0;  // :216
      }
      for (int64_t j = s; j < e; j++) {           // :218-221
        const Real w = bias + data[j];
        const Real *v = other + (int64_t)indices[j] * K;
        for (int64_t k = 0; k < K; k++) b[k] += w * v[k];
      }
      for (int64_t a = 0; a < K; a++) r[a] = b[a] - dot(P + a * K, x.data(), K);  // :222
      for (int64_t k = 0; k < K; k++) r[k] -= reg_u * x[k];                        // :223
      for (int64_t j = s; j < e; j++) {                                            // :224-228
        const Real *v = other + (int64_t)indices[j] * K;
        const Real coef = data[j] * dot(v, x.data(), K);
        for (int64_t k = 0; k < K; k++) r[k] -= coef * v[k];
      }
      for (int64_t k = 0; k < K; k++) p[k] = r[k];                   // :230
      const int64_t iters = max_cg_steps == 0 ? K : max_cg_steps;    // :232-234
      for (int64_t it = 0; it < iters; it++) {                       // :236-263
        const Real r2 = dot(r.data(), r.data(), K);
        if (r2 <= Real(1e-20)) break;
        for (int64_t a = 0; a < K; a++) Ap[a] = dot(P + a * K, p.data(), K);
        for (int64_t k = 0; k < K; k++) Ap[k] += reg_u * p[k];
        for (int64_t j = s; j < e; j++) {
          const Real *v = other + (int64_t)indices[j] * K;
          const Real coef = data[j] * dot(v, p.data(), K);
          for (int64_t k = 0; k < K; k++) Ap[k] += coef * v[k];
        }
        const Real den = dot(p.data(), Ap.data(), K);
        if (!(den > Real(0)) || !std::isfinite(den)) {  // :249-254
          status.store(STATUS_CG_SINGULAR);
          break;
        }
        const Real alpha = r2 / den;
        for (int64_t k = 0; k < K; k++) x[k] += alpha * p[k];
        for (int64_t k = 0; k < K; k++) r[k] -= alpha * Ap[k];
        const Real r2n = dot(r.data(), r.data(), K);
        if (r2n <= Real(1e-20)) break;
        const Real beta = r2n / r2;
        for (int64_t k = 0; k < K; k++) p[k] = r[k] + beta * p[k];
      }
      if (status.load(std::memory_order_relaxed) != STATUS_OK) break;
      for (int64_t k = 0; k < K; k++) xu[k] = x[k];  // :264
    }
  });
  return status.load();
}

// ---------------------------------------------------------------------------
// Cholesky row solve   (IALSTrainer.hpp:273-331 step_cholesky; :37-58 updater)
// A_upper = P_upper + sum_batches B^T B with B rows = sqrt(c) v, 64 per batch;
// diag += reg_u; LLT<Upper> (A = U^T U, reads the upper triangle only);
// solve U^T y = b, U x = y.  Empty rows are not special-cased.
// ---------------------------------------------------------------------------
template <typename Real>
int step_cholesky(Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,
                  const Real *data, const Real *other, int64_t n_other, int64_t K,
                  const Real *P, Real alpha0, Real reg, Real nu, int loss_type,
                  int n_threads, const Real *prior = nullptr) {  // prior: step_cholesky_with_prior, :333-385
  if (n_threads <= 0) return STATUS_INVALID;
  std::atomic<int64_t> cursor{0};
  std::atomic<int> status{STATUS_OK};
  const int64_t NB = 64;  // BatchedRankUpdater<64>
  run_workers(n_threads, [&](int) {
    std::vector<Real> A(K * K), B(K), buf(NB * K), y(K);
    const Real bias = loss_type == LOSS_IALSPP ? Real(0) : alpha0;
    auto flush = [&](int64_t nb) {  // selfadjointView<Upper>().rankUpdate(buf^T, 1)
      for (int64_t a = 0; a < K; a++)
        for (int64_t t = 0; t < nb; t++) {
          const Real ba = buf[t * K + a];
          const Real *row = buf.data() + t * K;
          Real *Arow = A.data() + a * K;
          for (int64_t c = a; c < K; c++) Arow[c] += ba * row[c];
        }
    };
    while (true) {
      int64_t u = cursor.fetch_add(1);
      if (u >= n_rows) break;
      if (status.load(std::memory_order_relaxed) != STATUS_OK) break;
      std::memcpy(A.data(), P, sizeof(Real) * K * K);  // :296
      const int64_t s = indptr[u], e = indptr[u + 1];
      if (prior) {  // B = diagonal * prior_u, :363-365
        const Real reg_p = compute_reg<Real>(e - s, n_other, alpha0, reg, nu);
        for (int64_t k = 0; k < K; k++) B[k] = reg_p * prior[u * K + k];
      } else {
        for (int64_t k = 0; k < K; k++) B[k] = 0;
      }
      int64_t nb = 0, nnz = 0;
      for (int64_t j = s; j < e; j++) {  // :301-307
        const Real *v = other + (int64_t)indices[j] * K;
        const Real sc = std::sqrt(data[j]);
        for (int64_t k = 0; k < K; k++) buf[nb * K + k] = sc * v[k];
        nb++;
        if (nb >= NB) { flush(nb); nb = 0; }
        const Real w = bias + data[j];
        for (int64_t k = 0; k < K; k++) B[k] += w * v[k];
        nnz++;
      }
      if (nb > 0) flush(nb);  // :308
      const Real reg_u = compute_reg<Real>(nnz, n_other, alpha0, reg, nu);
      for (int64_t k = 0; k < K; k++) A[k * K + k] += reg_u;  // :312-314
      // In-place upper Cholesky A = U^T U, right-looking so every inner loop
      // runs along a contiguous row (only the upper triangle is read/written).
      bool ok = true;
      for (int64_t i = 0; i < K; i++) {
        Real d = A[i * K + i];
        if (d <= Real(0)) { ok = false; break; }  // Eigen LLT: pivot <= 0 -> NumericalIssue
        d = std::sqrt(d);
        A[i * K + i] = d;
        const Real inv = Real(1) / d;
        Real *Ui = A.data() + i * K;
        for (int64_t c = i + 1; c < K; c++) Ui[c] *= inv;
        for (int64_t r = i + 1; r < K; r++) {
          const Real f = Ui[r];
          Real *Ar = A.data() + r * K;
          for (int64_t c = r; c < K; c++) Ar[c] -= f * Ui[c];
        }
      }
      if (!ok) { status.store(STATUS_CHOL_DECOMP); break; }  // :317-319
      for (int64_t i = 0; i < K; i++) {  // U^T y = B
        Real v = B[i];
        for (int64_t k = 0; k < i; k++) v -= A[k * K + i] * y[k];
        y[i] = v / A[i * K + i];
      }
      for (int64_t i = K - 1; i >= 0; i--) {  // U x = y
        Real v = y[i];
        for (int64_t c = i + 1; c < K; c++) v -= A[i * K + c] * y[c];
        y[i] = v / A[i * K + i];
      }
      bool finite = true;
      for (int64_t k = 0; k < K; k++) finite = finite && std::isfinite(y[k]);
      if (!finite) { status.store(STATUS_CHOL_SOLVE); break; }  // :320-323
      for (int64_t k = 0; k < K; k++) target[u * K + k] = y[k];  // :324
    }
  });
  return status.load();
}

// ---------------------------------------------------------------------------
// iALS++ half-epoch    (IALSTrainer.hpp:387-424 _prediction, :426-518 _step_dimrange,
//                       :520-535 step_ialspp)
// Per iteration: pred_j = x_u . y_i for every stored (u, i); then, for each block of
// `subspace_dim` consecutive dimensions [d0, d1) and every row u:
//   A = P[d0:d1, d0:d1] + sum c y_S y_S^T + reg_u I          (upper triangle, LLT<Upper>)
//   B = P[d0:d1, :] x_u + reg_u x_u[S] + sum (c (pred - 1) - bias) y_S
//   delta = A^-1 B;  x_u[S] -= delta;  pred_j -= delta . y_S  for the row's entries.
// The reference works on a copy of target[:, d0:d1] and writes it back after the block
// (:441-442, :517); every row only reads its own row, so updating in place is identical.
// The reference does not check the LLT here (:503-505); a non-positive pivot is reported
// as STATUS_CHOL_DECOMP instead of propagating NaN.
// ---------------------------------------------------------------------------
template <typename Real>
int step_ialspp(Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,
                const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,
                Real alpha0, Real reg, Real nu, int loss_type, int64_t subspace_dim,
                int64_t iterations, int n_threads) {
  if (n_threads <= 0 || subspace_dim <= 0 || iterations < 0) return STATUS_INVALID;
  std::vector<Real> pred((size_t)indptr[n_rows]);
  std::atomic<int> status{STATUS_OK};
  const Real bias = loss_type == LOSS_IALSPP ? Real(0) : alpha0;
  const int64_t NB = 64;  // BatchedRankUpdater<64>
  for (int64_t iter = 0; iter < iterations; iter++) {
    {  // _prediction (:387-424)
      std::atomic<int64_t> cursor{0};
      run_workers(n_threads, [&](int) {
        while (true) {
          const int64_t u = cursor.fetch_add(1);
          if (u >= n_rows) break;
          for (int64_t j = indptr[u]; j < indptr[u + 1]; j++)
            pred[j] = dot<Real>(target + u * K, other + (int64_t)indices[j] * K, K);
        }
      });
    }
    for (int64_t d0 = 0; d0 < K; d0 += subspace_dim) {  // :526-533
      const int64_t S = std::min(d0 + subspace_dim, K) - d0;
      std::atomic<int64_t> cursor{0};
      run_workers(n_threads, [&](int) {  // _step_dimrange (:426-518)
        std::vector<Real> A(S * S), B(S), buf(NB * S), y(S);
        auto flush = [&](int64_t nb) {  // selfadjointView<Upper>().rankUpdate(buf^T, 1)
          for (int64_t a = 0; a < S; a++)
            for (int64_t t = 0; t < nb; t++) {
              const Real ba = buf[t * S + a];
              const Real *row = buf.data() + t * S;
              Real *Arow = A.data() + a * S;
              for (int64_t c = a; c < S; c++) Arow[c] += ba * row[c];
            }
        };
        while (true) {
          const int64_t u = cursor.fetch_add(1);
          if (u >= n_rows) break;
          if (status.load(std::memory_order_relaxed) != STATUS_OK) break;
          Real *x = target + u * K;
          for (int64_t a = 0; a < S; a++)  // P_local = P_quadratic (:463)
            for (int64_t c = 0; c < S; c++) A[a * S + c] = P[(d0 + a) * K + d0 + c];
          const int64_t s = indptr[u], e = indptr[u + 1];
          const Real reg_u = compute_reg<Real>(e - s, n_other, alpha0, reg, nu);  // :471-472
          for (int64_t a = 0; a < S; a++)  // B = P_subspaced x + reg x_S (:474-478)
            B[a] = dot<Real>(P + (d0 + a) * K, x, K) + reg_u * x[d0 + a];
          int64_t nb = 0;
          for (int64_t j = s; j < e; j++) {  // :482-491
            const Real *v = other + (int64_t)indices[j] * K + d0;
            const Real residual = data[j] * (pred[j] - Real(1)) - bias;
            const Real sc = std::sqrt(data[j]);
            for (int64_t k = 0; k < S; k++) buf[nb * S + k] = sc * v[k];
            nb++;
            if (nb >= NB) { flush(nb); nb = 0; }
            for (int64_t k = 0; k < S; k++) B[k] += residual * v[k];
          }
          if (nb > 0) flush(nb);
          for (int64_t k = 0; k < S; k++) A[k * S + k] += reg_u;  // :493-495
          bool ok = true;
          for (int64_t i = 0; i < S; i++) {  // LLT<Upper> (:497)
            Real d = A[i * S + i];
            if (!(d > Real(0))) { ok = false; break; }
            d = std::sqrt(d);
            A[i * S + i] = d;
            const Real inv = Real(1) / d;
            Real *Ui = A.data() + i * S;
            for (int64_t c = i + 1; c < S; c++) Ui[c] *= inv;
            for (int64_t r = i + 1; r < S; r++) {
              const Real f = Ui[r];
              Real *Ar = A.data() + r * S;
              for (int64_t c = r; c < S; c++) Ar[c] -= f * Ui[c];
            }
          }
          if (!ok) { status.store(STATUS_CHOL_DECOMP); break; }
          for (int64_t i = 0; i < S; i++) {  // U^T z = B
            Real v = B[i];
            for (int64_t k = 0; k < i; k++) v -= A[k * S + i] * y[k];
            y[i] = v / A[i * S + i];
          }
          for (int64_t i = S - 1; i >= 0; i--) {  // U delta = z
            Real v = y[i];
            for (int64_t c = i + 1; c < S; c++) v -= A[i * S + c] * y[c];
            y[i] = v / A[i * S + i];
          }
          for (int64_t k = 0; k < S; k++) x[d0 + k] -= y[k];  // :499-500
          for (int64_t j = s; j < e; j++)                     // :502-508
            pred[j] -= dot<Real>(y.data(), other + (int64_t)indices[j] * K + d0, S);
        }
      });
      if (status.load() != STATUS_OK) return status.load();
    }
  }
  return status.load();
}

// ---------------------------------------------------------------------------
// Score block S = user[b:e] * item^T        (IALSTrainer.hpp:942-984 user_scores)
// ---------------------------------------------------------------------------
template <typename Real>
int user_scores(const Real *user, const Real *item, int64_t n_users, int64_t n_items, int64_t K,
                int64_t begin, int64_t end, int n_threads, Real *out) {
  if (n_threads <= 0 || end < begin || end > n_users) return STATUS_INVALID;  // :944-951
  const int64_t rows = end - begin;
  const int64_t target_chunks = (int64_t)n_threads * 4;
  const int64_t ideal = (rows + target_chunks - 1) / target_chunks;
  const int64_t chunk = std::clamp<int64_t>(ideal, 16, 128);  // :957-968
  std::atomic<int64_t> cursor{0};
  run_workers(n_threads, [&](int) {
    while (true) {
      int64_t b = cursor.fetch_add(chunk);
      if (b >= rows) break;
      int64_t e = std::min(b + chunk, rows);
      for (int64_t r = b; r < e; r++) {
        const Real *u = user + (begin + r) * K;
        Real *o = out + r * n_items;
        for (int64_t j = 0; j < n_items; j++) o[j] = dot(u, item + j * K, K);
      }
    }
  });
  return STATUS_OK;
}

// ---------------------------------------------------------------------------
// Loss                           (IALSTrainer.hpp:836-940 compute_loss, no features)
// Accumulated in Real like the reference (Real loss_local).
// ---------------------------------------------------------------------------
template <typename Real>
int compute_loss(const Real *user, const Real *item, int64_t U, int64_t I, int64_t K,
                 const int64_t *indptr, const int32_t *indices, const Real *data,
                 const int64_t *indptr_t, Real alpha0, Real reg, Real nu, int loss_type,
                 int n_threads, Real *out) {
  if (n_threads <= 0) return STATUS_INVALID;
  std::vector<Real> Pu(K * K), Pi(K * K);
  gram<Real>(item, I, K, alpha0, n_threads, Pu.data());  // user_solver.P
  gram<Real>(user, U, K, alpha0, n_threads, Pi.data());  // item_solver.P
  Real loss = 0;
  if (alpha0 != Real(0)) {
    Real s = 0;
    for (int64_t i = 0; i < K * K; i++) s += Pu[i] * Pi[i];
    loss = s / alpha0;
  }
  const Real bias = loss_type == LOSS_IALSPP ? Real(0) : alpha0;
  {
    std::vector<Real> part(n_threads, Real(0));
    std::atomic<int64_t> cursor{0};
    run_workers(n_threads, [&](int tid) {
      Real l = 0;
      while (true) {
        int64_t u = cursor.fetch_add(1);
        if (u >= U) break;
        int64_t nnz = 0;
        for (int64_t j = indptr[u]; j < indptr[u + 1]; j++) {
          nnz++;
          const Real pred = dot(user + u * K, item + (int64_t)indices[j] * K, K);
          l += data[j] * pred * pred - 2 * (data[j] + bias) * pred + data[j] + bias;
        }
        const Real ru = compute_reg<Real>(nnz, I, alpha0, reg, nu);
        l += ru * dot(user + u * K, user + u * K, K);
      }
      part[tid] = l;
    });
    for (int t = 0; t < n_threads; t++) loss += part[t];
  }
  {
    std::vector<Real> part(n_threads, Real(0));
    std::atomic<int64_t> cursor{0};
    run_workers(n_threads, [&](int tid) {
      Real l = 0;
      while (true) {
        int64_t i = cursor.fetch_add(1);
        if (i >= I) break;
        const int64_t nnz = indptr_t[i + 1] - indptr_t[i];
        const Real ri = compute_reg<Real>(nnz, U, alpha0, reg, nu);
        l += ri * dot(item + i * K, item + i * K, K);
      }
      part[tid] = l;
    });
    for (int t = 0; t < n_threads; t++) loss += part[t];
  }
  *out = loss / 2;
  return STATUS_OK;
}

// One epoch                                     (IALSTrainer.hpp:784-788 step)
template <typename Real>
int epoch(Real *user, Real *item, int64_t U, int64_t I, int64_t K, const int64_t *indptr,
          const int32_t *indices, const Real *data, const int64_t *indptr_t,
          const int32_t *indices_t, const Real *data_t, Real alpha0, Real reg, Real nu,
          int loss_type, int solver_type, int max_cg_steps, int n_threads) {
  std::vector<Real> P(K * K);
  int st = gram<Real>(item, I, K, alpha0, n_threads, P.data());
  if (st) return st;
  st = solver_type == 1
           ? step_cg<Real>(user, U, indptr, indices, data, item, I, K, P.data(), alpha0, reg, nu,
                           loss_type, max_cg_steps, n_threads)
           : step_cholesky<Real>(user, U, indptr, indices, data, item, I, K, P.data(), alpha0,
                                 reg, nu, loss_type, n_threads);
  if (st) return st;
  st = gram<Real>(user, U, K, alpha0, n_threads, P.data());
  if (st) return st;
  st = solver_type == 1
           ? step_cg<Real>(item, I, indptr_t, indices_t, data_t, user, U, K, P.data(), alpha0, reg,
                           nu, loss_type, max_cg_steps, n_threads)
           : step_cholesky<Real>(item, I, indptr_t, indices_t, data_t, user, U, K, P.data(),
                                 alpha0, reg, nu, loss_type, n_threads);
  return st;
}

// ---------------------------------------------------------------------------
// Evaluator core       (cpp_source/evaluator.cpp:42-48 discount, :49-179 Metrics,
//                       :292-367 get_metrics_local; recommendable_items empty)
// acc layout: [0]=valid_user [1]=total_user [2]=hit [3]=recall [4]=ndcg
//             [5]=precision [6]=map            item_cnt: int64[n_items]
// rec_out (optional): int32[rows*cutoff] filled with -1 then the top lists;
// rec_cnt (optional): int32[rows] = n_recommendable per row (-1 if skipped).
// ---------------------------------------------------------------------------
template <typename Score>
int topk_metrics(const Score *scores, int64_t rows, int64_t n_items, const int64_t *gt_indptr,
                 const int32_t *gt_indices, int64_t offset, int64_t cutoff, int recall_with_cutoff,
                 double *acc, int64_t *item_cnt, int32_t *rec_out, int32_t *rec_cnt) {
  if (cutoff <= 0 || cutoff > n_items) return STATUS_INVALID;  // :265-266
  std::vector<double> discount(n_items);
  for (int64_t i = 0; i < n_items; i++) discount[i] = 1 / std::log2(2 + (double)i);  // :42-48
  std::vector<std::pair<Score, int32_t>> cand;
  cand.reserve(n_items);
  for (int64_t u = 0; u < rows; u++) {
    const int64_t uo = u + offset;
    acc[1] += 1;  // increment_total_user :317
    if (rec_cnt) rec_cnt[u] = -1;
    if (rec_out) for (int64_t c = 0; c < cutoff; c++) rec_out[u * cutoff + c] = -1;
    const int64_t gs = gt_indptr[uo], ge = gt_indptr[uo + 1];
    if (ge == gs) continue;  // :319-321
    std::unordered_set<int64_t> gt(gt_indices + gs, gt_indices + ge);
    cand.clear();
    const Score *row = scores + u * n_items;
    for (int64_t j = 0; j < n_items; j++)  // :324-331
      if (row[j] != -std::numeric_limits<Score>::infinity()) cand.emplace_back(-row[j], (int32_t)j);
    const int64_t n_rec = std::min<int64_t>(cutoff, (int64_t)cand.size());  // :350-351
    std::partial_sort(cand.begin(), cand.begin() + n_rec, cand.end());    // :353-355
    if (rec_cnt) rec_cnt[u] = (int32_t)n_rec;
    // Metrics::update :127-166
    const int64_t n_gt = (int64_t)gt.size();
    acc[0] += 1;
    if (n_rec == 0) continue;
    double dcg = 0, idcg = 0, ap = 0;
    for (int64_t i = 0; i < std::min(n_gt, n_rec); i++) idcg += discount[i];
    int64_t cum_hit = 0;
    for (int64_t i = 0; i < n_rec; i++) {
      const int32_t idx = cand[i].second;
      if (rec_out) rec_out[u * cutoff + i] = idx;
      item_cnt[idx] += 1;
      if (gt.count(idx)) {
        dcg += discount[i];
        cum_hit++;
        ap += (double)cum_hit / (double)(i + 1);
      }
    }
    if (cum_hit > 0) acc[2] += 1;
    acc[5] += cum_hit / (double)n_rec;
    acc[3] += cum_hit / (double)(recall_with_cutoff ? (n_gt > n_rec ? n_rec : n_gt) : n_gt);
    acc[4] += dcg / idcg;
    acc[6] += ap / n_gt;
  }
  return STATUS_OK;
}

}  // namespace

// ---- C ABI (ctypes) ---------------------------------------------------------
#define ORACLE_API extern "C" __attribute__((visibility("default")))

#define DEFINE_FOR(SFX, Real)                                                                     \
  ORACLE_API int oracle_gram_##SFX(const Real *Y, int64_t n, int64_t K, Real alpha0,              \
                                   int n_threads, Real *P) {                                      \
    return gram<Real>(Y, n, K, alpha0, n_threads, P);                                             \
  }                                                                                               \
  ORACLE_API int oracle_step_cg_##SFX(Real *target, int64_t n_rows, const int64_t *indptr,        \
                                      const int32_t *indices, const Real *data, const Real *other, \
                                      int64_t n_other, int64_t K, const Real *P, Real alpha0,     \
                                      Real reg, Real nu, int loss_type, int max_cg_steps,         \
                                      int n_threads) {                                            \
    return step_cg<Real>(target, n_rows, indptr, indices, data, other, n_other, K, P, alpha0,     \
                         reg, nu, loss_type, max_cg_steps, n_threads);                            \
  }                                                                                               \
  ORACLE_API int oracle_step_cholesky_##SFX(                                                      \
      Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,                \
      const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,             \
      Real alpha0, Real reg, Real nu, int loss_type, int n_threads) {                             \
    return step_cholesky<Real>(target, n_rows, indptr, indices, data, other, n_other, K, P,       \
                               alpha0, reg, nu, loss_type, n_threads);                            \
  }                                                                                               \
  ORACLE_API int oracle_step_cg_prior_##SFX(                                                      \
      Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,                \
      const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,             \
      Real alpha0, Real reg, Real nu, int loss_type, int max_cg_steps, int n_threads,             \
      const Real *prior) {                                                                        \
    return step_cg<Real>(target, n_rows, indptr, indices, data, other, n_other, K, P, alpha0,     \
                         reg, nu, loss_type, max_cg_steps, n_threads, prior);                     \
  }                                                                                               \
  ORACLE_API int oracle_step_cholesky_prior_##SFX(                                                \
      Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,                \
      const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,             \
      Real alpha0, Real reg, Real nu, int loss_type, int n_threads, const Real *prior) {          \
    return step_cholesky<Real>(target, n_rows, indptr, indices, data, other, n_other, K, P,       \
                               alpha0, reg, nu, loss_type, n_threads, prior);                     \
  }                                                                                               \
  ORACLE_API int oracle_step_ialspp_##SFX(                                                        \
      Real *target, int64_t n_rows, const int64_t *indptr, const int32_t *indices,                \
      const Real *data, const Real *other, int64_t n_other, int64_t K, const Real *P,             \
      Real alpha0, Real reg, Real nu, int loss_type, int64_t subspace_dim, int64_t iterations,    \
      int n_threads) {                                                                            \
    return step_ialspp<Real>(target, n_rows, indptr, indices, data, other, n_other, K, P,         \
                             alpha0, reg, nu, loss_type, subspace_dim, iterations, n_threads);    \
  }                                                                                               \
  ORACLE_API int oracle_user_scores_##SFX(const Real *user, const Real *item, int64_t n_users,    \
                                          int64_t n_items, int64_t K, int64_t begin, int64_t end, \
                                          int n_threads, Real *out) {                             \
    return user_scores<Real>(user, item, n_users, n_items, K, begin, end, n_threads, out);        \
  }                                                                                               \
  ORACLE_API int oracle_compute_loss_##SFX(                                                       \
      const Real *user, const Real *item, int64_t U, int64_t I, int64_t K, const int64_t *indptr, \
      const int32_t *indices, const Real *data, const int64_t *indptr_t, Real alpha0, Real reg,   \
      Real nu, int loss_type, int n_threads, Real *out) {                                         \
    return compute_loss<Real>(user, item, U, I, K, indptr, indices, data, indptr_t, alpha0, reg,  \
                              nu, loss_type, n_threads, out);                                     \
  }                                                                                               \
  ORACLE_API int oracle_epoch_##SFX(Real *user, Real *item, int64_t U, int64_t I, int64_t K,      \
                                    const int64_t *indptr, const int32_t *indices,                \
                                    const Real *data, const int64_t *indptr_t,                    \
                                    const int32_t *indices_t, const Real *data_t, Real alpha0,    \
                                    Real reg, Real nu, int loss_type, int solver_type,            \
                                    int max_cg_steps, int n_threads) {                            \
    return epoch<Real>(user, item, U, I, K, indptr, indices, data, indptr_t, indices_t, data_t,   \
                       alpha0, reg, nu, loss_type, solver_type, max_cg_steps, n_threads);         \
  }                                                                                               \
  ORACLE_API int oracle_topk_metrics_##SFX(                                                       \
      const Real *scores, int64_t rows, int64_t n_items, const int64_t *gt_indptr,                \
      const int32_t *gt_indices, int64_t offset, int64_t cutoff, int recall_with_cutoff,          \
      double *acc, int64_t *item_cnt, int32_t *rec_out, int32_t *rec_cnt) {                       \
    return topk_metrics<Real>(scores, rows, n_items, gt_indptr, gt_indices, offset, cutoff,       \
                              recall_with_cutoff, acc, item_cnt, rec_out, rec_cnt);               \
  }

DEFINE_FOR(f32, float)
DEFINE_FOR(f64, double)

ORACLE_API int oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
