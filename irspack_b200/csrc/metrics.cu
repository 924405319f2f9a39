// Ranking metrics of a block of users on the device: the bookkeeping of
// EvaluatorCore::get_metrics_local + Metrics::update (/root/reference/cpp_source/evaluator.cpp:127-166,
// 308-361) for recommendation lists that the fused top-k kernel (score_tc.cu) has already produced.
// The reference walks the users on host threads; a numpy restatement of that loop (np.isin over
// rows * cutoff keys) took ~3 s for the 138 493 users of configs[1] next to 8 ms of scoring.
//
// One warp per user.  Lane j takes position j of the list (32 positions per step, in order):
//   hit_j = rec_j in the user's ground-truth row (binary search; the row is sorted),
//   dcg  += hit_j * discount_j,   ap += hit_j * (hits up to j) / (j + 1),   item_cnt[rec_j] += 1,
// then  hit += (hits > 0), precision += hits / n_rec, recall += hits / n_gt (or min(n_gt, n_rec)),
// ndcg += dcg / idcg(min(n_gt, n_rec)), map += ap / n_gt   -- users without ground truth are
// skipped (:319-321), users without recommendable items only count as valid (:131-133).
// All sums are float64 (the reference's); discount_j = 1 / log2(2 + j) and its prefix sums come from
// the host so that both sides use the very same table.
#include "common.cuh"

namespace ials {
namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void metrics_rows_kernel(const int32_t *__restrict__ rec, const int32_t *__restrict__ cnt,
                                    int64_t rows, int k, const int64_t *__restrict__ gt_indptr,
                                    const int32_t *__restrict__ gt_indices,
                                    const double *__restrict__ discount, const double *__restrict__ cum_discount,
                                    int recall_with_cutoff, double *acc /*hit recall ndcg map precision*/,
                                    unsigned long long *valid_user, unsigned long long *item_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double a_hit = 0, a_recall = 0, a_ndcg = 0, a_map = 0, a_prec = 0;
  unsigned long long a_valid = 0;
  for (int64_t u = warp; u < rows; u += n_warps) {
    const int64_t g0 = gt_indptr[u], g1 = gt_indptr[u + 1];
    const int64_t n_gt = g1 - g0;
    if (n_gt == 0) continue;
    a_valid++;
    const int n_rec = min(cnt[u], k);
    if (n_rec <= 0) continue;
    double dcg = 0, ap = 0;
    int hits = 0;
    for (int base = 0; base < n_rec; base += 32) {
      const int j = base + lane;
      const int item = j < n_rec ? rec[u * k + j] : -1;
      bool hit = false;
      if (item >= 0) {
        int64_t lo = g0, hi = g1;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (gt_indices[mid] < item) lo = mid + 1; else hi = mid;
        }
        hit = lo < g1 && gt_indices[lo] == item;
        atomicAdd(&item_cnt[item], 1ull);
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int upto = hits + __popc(ballot & (0xffffffffu >> (31 - lane)));
        dcg += discount[j];
        ap += (double)upto / (double)(j + 1);
      }
      hits += __popc(ballot);
    }
    dcg = warp_sum_d(dcg);
    ap = warp_sum_d(ap);
    const int64_t lim = min(n_gt, (int64_t)n_rec);
    a_hit += hits > 0 ? 1.0 : 0.0;
    a_prec += (double)hits / (double)n_rec;
    a_recall += (double)hits / (double)(recall_with_cutoff ? lim : n_gt);
    a_ndcg += dcg / cum_discount[lim - 1];
    a_map += ap / (double)n_gt;
  }
  if (lane == 0) {  // every lane holds the same per-user values: one lane adds the warp's share
    if (a_valid) atomicAdd(valid_user, a_valid);
    if (a_hit != 0) atomicAdd(acc + 0, a_hit);
    if (a_recall != 0) atomicAdd(acc + 1, a_recall);
    if (a_ndcg != 0) atomicAdd(acc + 2, a_ndcg);
    if (a_map != 0) atomicAdd(acc + 3, a_map);
    if (a_prec != 0) atomicAdd(acc + 4, a_prec);
  }
}

}  // namespace

void launch_metrics_rows(const int32_t *rec, const int32_t *cnt, int64_t rows, int k, const int64_t *gt_indptr,
                         const int32_t *gt_indices, const double *discount, const double *cum_discount,
                         int recall_with_cutoff, double *acc, unsigned long long *valid_user,
                         unsigned long long *item_cnt, cudaStream_t s) {
  if (rows == 0) return;
  const int T = 256;
  const int64_t want = ceil_div(rows * kWarp, T);
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)kNumSMsB200 * 16);
  metrics_rows_kernel<<<grid, T, 0, s>>>(rec, cnt, rows, k, gt_indptr, gt_indices, discount, cum_discount,
                                         recall_with_cutoff, acc, valid_user, item_cnt);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
