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
// diagonal block is free for its factor (unscaled pivot rows: LDL^T order of operations = Cholesky
// without the square roots).  The blocks' matrices do not depend on x: they are all factored
// first, one WARP per block with the block's columns in registers and no block barrier, then the
// sweeps only need a residual (all threads) and two substitutions (warp 0, one shuffle per step)
// per block.  History (ML-20M shape, ms per epoch incl. the Gram): elimination in shared memory
// 58.5 (r02aj; every load behind the store before it); columns in registers, rolled and predicated
// 46.5 (r02am; 259 instructions per pivot and warp); unrolled with compile-time pivots 32.1 (r02an;
// 520 cycles per pivot between two block barriers, 57 % of the kernel); this form: see DESIGN.md.
// A non-positive pivot / non-finite solution raise the flags of the Cholesky solver (the
// reference's LLT has no other failure mode here).
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
    // the next row of this CTA: its first W (64 KB, written by the Gram kernel a launch ago, mostly
    // in HBM by now) starts towards the L2 while this row is solved (r02ao ncu: 18 % of the
    // kernel waited for these loads with 12 warps per SM)
    if (d.heavy_first_job != nullptr && h + gridDim.x < d.n_heavy) {
      const char *nxt = reinterpret_cast<const char *>(
          d.W + (size_t)(d.heavy_first_job[h + gridDim.x] - d.job0) * KP * KP);
#pragma unroll
      for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)(q * kThreads + t) * 128));
    }
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

    // ---- the diagonal blocks are factored first, warp w the blocks w, w + 4, ... (their matrices
    // do not depend on x), each by ONE warp without block barriers: lane l keeps columns l and
    // l + 32 of the block's upper triangle in registers, the finished (unscaled) pivot row goes to
    // the block's upper triangle in shared memory -- where the substitutions read it later -- and
    // comes back to every lane by broadcast loads.  (r02an ncu of the CTA-wide elimination with a
    // block barrier per pivot: 520 cycles per pivot, 57 % of the kernel.)
    const int n_blocks = (K + S - 1) / S;
    for (int blk = warp; blk < n_blocks; blk += kThreads / 32) {
      const int d0 = blk * S, Sd = min(S, K - d0);
      const bool ok0 = lane < Sd, ok1 = lane + 32 < Sd;
      float c0[32], c1[64];
      {
        const float *s0 = A + (d0 + (ok0 ? lane : 0)) * LDA + d0;
        const float *s1 = A + (d0 + (ok1 ? lane + 32 : 0)) * LDA + d0;
#pragma unroll
        for (int i = 0; i < 32; i++) c0[i] = (ok0 && i < lane) ? s0[i] : 0.f;
#pragma unroll
        for (int i = 0; i < 64; i++) c1[i] = (ok1 && i < lane + 32) ? s1[i] : 0.f;
        const float dg0 = diag[d0 + (ok0 ? lane : 0)], dg1 = diag[d0 + (ok1 ? lane + 32 : 0)];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          if (i == lane) c0[i] = ok0 ? dg0 : 0.f;
          if (i == lane) c1[i + 32] = ok1 ? dg1 : 0.f;
        }
      }
      __syncwarp();  // every lane has read its columns from the lower triangle
      bool bad = false;
      // pivots 0..31: rows of both columns; pivots 32..63: rows of column l + 32 only.  Two loops,
      // each unrolled completely (k is a compile-time constant: registers with fixed names).
#pragma unroll
      for (int k = 0; k < 32; k++) {
        if (k >= Sd || bad) break;  // warp-uniform
        float *rowk = A + (d0 + k) * LDA + d0;
        if (ok0 && lane >= k) rowk[lane] = c0[k];
        if (ok1) rowk[lane + 32] = c1[k];
        __syncwarp();
        const float piv = rowk[k];
        if (!(piv > 0.f)) {  // every lane reads the same word
          bad = true;
          break;
        }
        const float inv = 1.0f / piv;
        const float m0 = c0[k] * inv, m1 = c1[k] * inv;
#pragma unroll
        for (int i = k + 1; i < 32; i++) {
          const float aki = rowk[i];
          c0[i] = fmaf(-aki, m0, c0[i]);
          c1[i] = fmaf(-aki, m1, c1[i]);
        }
        if (Sd > 32) {
#pragma unroll
          for (int i = 32; i < 64; i++) c1[i] = fmaf(-rowk[i], m1, c1[i]);
        }
      }
#pragma unroll
      for (int k = 32; k < 64; k++) {
        if (k >= Sd || bad) break;  // warp-uniform
        float *rowk = A + (d0 + k) * LDA + d0;
        if (ok1 && lane + 32 >= k) rowk[lane + 32] = c1[k];
        __syncwarp();
        const float piv = rowk[k];
        if (!(piv > 0.f)) {
          bad = true;
          break;
        }
        const float m1 = c1[k] * (1.0f / piv);
#pragma unroll
        for (int i = k + 1; i < 64; i++) c1[i] = fmaf(-rowk[i], m1, c1[i]);
      }
      if (bad && lane == 0) s_fail = 1;
    }
    __syncthreads();
    bool failed = s_fail != 0;

    // ---- the sweeps: residual of the block, forward and backward substitution by warp 0
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
        if (warp == 0) {
          // lane l keeps entries l and l + 32 of the block's vector; U (unscaled rows) is the
          // block's upper triangle:  z_i -= U_ki z_k / U_kk  (k ascending),  then
          // delta_j = z_j / U_jj,  z_i -= U_ij delta_j  (j descending)
          const bool ok0 = lane < Sd, ok1 = lane + 32 < Sd;
          const int i0 = d0 + (ok0 ? lane : 0), i1 = d0 + (ok1 ? lane + 32 : 0);
          float z0 = ok0 ? z[i0] : 0.f, z1 = ok1 ? z[i1] : 0.f;
          const float inv0 = ok0 ? 1.0f / A[i0 * LDA + i0] : 0.f;
          const float inv1 = ok1 ? 1.0f / A[i1 * LDA + i1] : 0.f;
          const float *up0 = A + d0 * LDA + i0, *up1 = A + d0 * LDA + i1;  // column i of U, row k at k * LDA
          for (int k = 0; k < Sd; k++) {
            const float tk = __shfl_sync(0xffffffffu, k >= 32 ? z1 * inv1 : z0 * inv0, k & 31);
            const float u0 = up0[k * LDA], u1 = up1[k * LDA];
            if (ok0 && lane > k) z0 = fmaf(-u0, tk, z0);
            if (ok1 && lane + 32 > k) z1 = fmaf(-u1, tk, z1);
          }
          float dl0 = 0.f, dl1 = 0.f;
          const float *rw0 = A + i0 * LDA + d0, *rw1 = A + i1 * LDA + d0;  // row i of U
          for (int j = Sd - 1; j >= 0; j--) {
            const int owner = j & 31;
            const float dj = __shfl_sync(0xffffffffu, j >= 32 ? z1 * inv1 : z0 * inv0, owner);
            if (lane == owner) {
              if (j >= 32) dl1 = dj; else dl0 = dj;
            }
            const float u0 = rw0[j], u1 = rw1[j];
            if (ok0 && lane < j) z0 = fmaf(-u0, dj, z0);
            if (ok1 && lane + 32 < j) z1 = fmaf(-u1, dj, z1);
          }
          if (ok0) z[i0] = dl0;
          if (ok1) z[i1] = dl1;
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
  const size_t smem = sizeof(float) * (KP * LDA + 3 * KP + 64);  // the tail: row loads past a short block stay inside
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
