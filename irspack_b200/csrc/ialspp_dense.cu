// iALS++ (Solver::step_ialspp / _prediction / _step_dimrange,
// /root/reference/cpp_source/als/IALSTrainer.hpp:387-535) as block Gauss-Seidel on the explicitly
// formed normal equations of a row.
//
// The reference keeps one prediction per stored entry, pred_i = x . y_i, and for the block
// D = [d0, d0 + S) of the factor dimensions solves
//     (P_DD + reg I + sum_i c_i y_iD y_iD^T) delta = P_D: x + reg x_D + sum_i (c_i (pred_i - 1) - bias) y_iD
// then x_D -= delta and pred_i -= delta . y_iD (:455-508).  With G = sum_i c_i y_i y_i^T (K x K),
// b = sum_i (c_i + bias) y_i and A = P + G + reg I the right-hand side is
//     P_D: x + reg x_D + G_D: x - b_D = (A x - b)_D
// (sum_i c_i pred_i y_iD = G_D: x, and the prediction update keeps pred_i = x . y_i), and the
// matrix is A_DD: one sweep of block Gauss-Seidel on A x = b, rows independent of each other.
// So the whole half-epoch of a row needs G and b ONCE -- the tensor-core Gram of wgram.cu, one
// pass over the row's neighbours instead of (1 + 3 K / S) -- and then only K x K work:
//     for every sweep, for every block D:  r = (A x - b)_D;  A_DD delta = r;  x_D -= delta.
// Same linear systems, same update order; the rounding differs (a dense row dot instead of a sum
// over the stored entries), which the parity tests bound against the f32 and f64 oracles.
//
// One 128-thread CTA per row (three per SM): A in shared memory ([128][129] floats), thread t owns
// row t.  The matrix-vector products read the lower triangle only, so the upper triangle of a
// diagonal block is free for its factor.  Elimination of a block (S <= 64): thread (column c, row
// parity) keeps its half of column c in REGISTERS (r02ak: the first version updated the block in
// shared memory and ran 131 us per row -- every load of the inner loop waited for the store before
// it); right-looking, unscaled pivot rows (LDL^T order of operations = Cholesky without the square
// roots); the pivot row travels through a double-buffered 64-float row in shared memory, one
// barrier per pivot; r rides along as an extra column (the forward substitution); warp 0 does the
// backward substitution with the unknowns in registers and one shuffle per pivot.  A non-positive pivot / non-finite solution
// raise the flags of the Cholesky solver (the reference's LLT has no other failure mode here).
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;
constexpr int LDA = KP + 1;
constexpr int kThreads = KP;

__global__ void __launch_bounds__(kThreads) ialspp_dense_kernel(DenseSolveArgs d, int S, int iters) {
  extern __shared__ __align__(16) float smem[];
  float *A = smem;               // [128][129]
  float *xs = A + KP * LDA;      // [128] current x
  float *z = xs + KP;            // [128] riding column / delta
  float *diag = z + KP;          // [128] diagonal of A
  float *prow = diag + KP;       // [2][64] pivot row of the elimination, double buffered
  float *zpiv = prow + 128;      // [2] its riding-column entry
  __shared__ int s_fail;
  const SolveArgs &a = d.base;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int K = a.K;

  for (int64_t h = blockIdx.x; h < d.n_heavy; h += gridDim.x) {
    const int64_t u = a.order[h];
    const int64_t gu = a.row_base + u;
    const int j0 = d.heavy_first_job ? d.heavy_first_job[h] - d.job0 : 0;
    const int j1 = d.heavy_first_job ? d.heavy_first_job[h + 1] - d.job0 : 0;
    const int64_t nnz = a.indptr[u + 1] - a.indptr[u];
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)nnz, a.nu);  // :117-120
    __syncthreads();  // the previous row's readers of A / xs / z are done
    if (t == 0) s_fail = 0;
    // T = sum_j W_j + P / 2  (A = T + T^T + reg I), 32 rows per step, 8 loads in flight per thread
    {
      const int rsub = t >> 5, c4 = (t & 31) * 4;
      for (int r0 = 0; r0 < KP; r0 += 32) {
        float4 acc[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 pq = *reinterpret_cast<const float4 *>(a.P + (size_t)(r0 + q * 4 + rsub) * KP + c4);
          acc[q] = make_float4(0.5f * pq.x, 0.5f * pq.y, 0.5f * pq.z, 0.5f * pq.w);
        }
        for (int j = j0; j < j1; j++) {
          const float *Wj = d.W + (size_t)j * KP * KP;
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 wq = *reinterpret_cast<const float4 *>(Wj + (size_t)(r0 + q * 4 + rsub) * KP + c4);
            acc[q].x += wq.x; acc[q].y += wq.y; acc[q].z += wq.z; acc[q].w += wq.w;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float *dst = A + (r0 + q * 4 + rsub) * LDA + c4;
          dst[0] = acc[q].x; dst[1] = acc[q].y; dst[2] = acc[q].z; dst[3] = acc[q].w;
        }
      }
    }
    float b = 0.f;
    for (int j = j0; j < j1; j++)
#pragma unroll
      for (int q = 0; q < kWGramBParts; q++) b += d.bpart[((size_t)j * kWGramBParts + q) * KP + t];
    if (a.ready_flags != nullptr && t == 0) wait_row_ready(a, gu);  // warm start still arriving?
    __syncthreads();
    float x = a.target[gu * KP + t];
    // lower triangle of A (row t, columns < t) and the diagonal; the upper triangle becomes scratch.
    // Loads of a batch are issued before its stores (the compiler cannot reorder them itself).
    for (int i0 = 0; i0 < t; i0 += 8) {
      float lo[8], up[8];
#pragma unroll
      for (int e = 0; e < 8; e++)
        if (i0 + e < t) {
          lo[e] = A[t * LDA + i0 + e];
          up[e] = A[(i0 + e) * LDA + t];
        }
#pragma unroll
      for (int e = 0; e < 8; e++)
        if (i0 + e < t) A[t * LDA + i0 + e] = lo[e] + up[e];
    }
    const float adiag = 2.f * A[t * LDA + t] + reg_u;
    diag[t] = adiag;
    xs[t] = x;
    __syncthreads();

    bool failed = false;
    const int c = t & 63, half = t >> 6;  // elimination role: column of the block, row parity
    for (int it = 0; it < iters && !failed; it++) {
      for (int d0 = 0; d0 < K; d0 += S) {
        const int Sd = min(S, K - d0);
        // r_t = (A x - b)_t from the lower triangle: row t left of the diagonal, column t below it
        float r0 = adiag * xs[t], r1 = 0.f;
        {
          const float *row = A + t * LDA;
          int j = 0;
          for (; j + 1 < t; j += 2) {
            r0 = fmaf(row[j], xs[j], r0);
            r1 = fmaf(row[j + 1], xs[j + 1], r1);
          }
          if (j < t) r0 = fmaf(row[j], xs[j], r0);
          const float *col = A + t;
          j = t + 1;
          for (; j + 1 < KP; j += 2) {
            r0 = fmaf(col[j * LDA], xs[j], r0);
            r1 = fmaf(col[(j + 1) * LDA], xs[j + 1], r1);
          }
          if (j < KP) r0 = fmaf(col[j * LDA], xs[j], r0);
        }
        const bool mine = t >= d0 && t < d0 + Sd;
        if (mine) z[t] = (r0 + r1) - b;
        __syncthreads();
        // this thread's half of column cj: local rows li = 2 u + half <= c, from the lower triangle
        const int cj = d0 + c;
        const bool col_ok = c < Sd;
        float colv[32];
        {
          // branch-free: every load stays inside A (rows of a column that does not exist are
          // read from column d0 and dropped)
          const float *src = A + (col_ok ? cj : d0) * LDA + d0;
          const float dg = diag[col_ok ? cj : d0];
#pragma unroll
          for (int uu = 0; uu < 32; uu++) {
            const int li = 2 * uu + half;
            const float v = src[li];
            colv[uu] = (col_ok && li < c) ? v : ((col_ok && li == c) ? dg : 0.f);
          }
        }
        float zr = (col_ok && half == 0) ? z[cj] : 0.f;
        if (col_ok && half == 0) {  // pivot row 0
          prow[c] = colv[0];
          A[d0 * LDA + cj] = colv[0];
          if (c == 0) zpiv[0] = zr;
        }
        __syncthreads();
        // The pivot loop is unrolled completely: with k a compile-time constant the rows of the
        // column are registers with fixed names, the pivot row is read at immediate offsets and
        // nothing is predicated or selected (r02am ncu: 259 instructions per pivot and warp in the
        // rolled, predicated form -- 133 k warp instructions per row, issue-bound).  Rows from
        // 2 (k >> 1) on are updated: for one parity that includes a row <= k, already final and
        // published, whose register is never read again.
        const float *prh = prow + half;
#pragma unroll
        for (int k = 0; k < 64; k++) {
          if (k >= Sd) break;  // warp-uniform
          const float *pr = prh + 64 * (k & 1);
          float *pn = prow + 64 * ((k + 1) & 1);
          const float piv = prow[64 * (k & 1) + k];
          if (!(piv > 0.f)) {  // every thread reads the same word
            failed = true;
            break;
          }
          const bool act = col_ok && c > k;
          const float akj = act ? prow[64 * (k & 1) + c] * (1.0f / piv) : 0.f;
#pragma unroll
          for (int uu = k >> 1; uu < 32; uu++) colv[uu] = fmaf(-pr[2 * uu], akj, colv[uu]);
          if (act && ((k + 1) & 1) == half) {  // row k + 1 is final: the next pivot row, a row of the factor
            pn[c] = colv[(k + 1) >> 1];
            A[(d0 + k + 1) * LDA + cj] = colv[(k + 1) >> 1];
          }
          if (act && half == 0) {
            zr = fmaf(-akj, zpiv[k & 1], zr);
            if (c == k + 1) {
              zpiv[(k + 1) & 1] = zr;
              z[cj] = zr;
            }
          }
          __syncthreads();
        }
        if (failed) break;
        // U delta = z from the bottom: lane l of warp 0 keeps unknowns l and l + 32 of the block
        if (warp == 0) {
          const int i0 = d0 + lane, i1 = d0 + lane + 32;
          float z0 = lane < Sd ? z[i0] : 0.f, z1 = lane + 32 < Sd ? z[i1] : 0.f;
          const float inv0 = lane < Sd ? 1.0f / A[i0 * LDA + i0] : 0.f;
          const float inv1 = lane + 32 < Sd ? 1.0f / A[i1 * LDA + i1] : 0.f;
          float dl0 = 0.f, dl1 = 0.f;
          for (int j = Sd - 1; j >= 0; j--) {
            const int owner = j & 31;
            const float mine_d = j >= 32 ? z1 * inv1 : z0 * inv0;
            const float dj = __shfl_sync(0xffffffffu, mine_d, owner);
            if (lane == owner) {
              if (j >= 32) dl1 = dj; else dl0 = dj;
            }
            const int jj = d0 + j;
            if (lane < j && lane < Sd) z0 = fmaf(-A[i0 * LDA + jj], dj, z0);
            if (lane + 32 < j) z1 = fmaf(-A[i1 * LDA + jj], dj, z1);
          }
          if (lane < Sd) z[i0] = dl0;
          if (lane + 32 < Sd) z[i1] = dl1;
        }
        __syncthreads();
        if (mine) {
          x -= z[t];  // :499-500
          xs[t] = x;
        }
        __syncthreads();
      }
    }
    if (failed) {
      if (t == 0) atomicExch(&a.err_flags[kErrCholDecomp], 1);
      continue;
    }
    const bool finite = isfinite(x);
    if (!finite) atomicExch(&s_fail, 1);
    __syncthreads();
    if (s_fail) {
      if (t == 0) atomicExch(&a.err_flags[kErrCholSolve], 1);
      continue;
    }
    if (t < K) {
      a.target[gu * KP + t] = x;
      for (int pi = 0; pi < a.n_peers; pi++) a.peers[pi][gu * KP + t] = x;
    }
  }
}

}  // namespace

bool ialspp_dense_supported(const SolveArgs &a, int S) { return a.ld == KP && S >= 1 && S <= 64; }

// Rows order[0 .. n_heavy) with their Gram jobs (heavy_first_job == nullptr: rows without
// interactions, A = P + reg I, b = 0).
void launch_ialspp_dense(const DenseSolveArgs &d, int S, int iters, cudaStream_t s) {
  if (d.n_heavy <= 0) return;
  const size_t smem = sizeof(float) * (KP * LDA + 4 * KP + 8);
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(ialspp_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  const unsigned grid = (unsigned)std::min<int64_t>(d.n_heavy, (int64_t)kNumSMsB200 * 3);
  ialspp_dense_kernel<<<grid, kThreads, smem, s>>>(d, S, iters);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
