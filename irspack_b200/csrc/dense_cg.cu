// K2 for heavy rows: conjugate gradient on EXPLICITLY formed normal equations.
//
// Solver::step_cg (/root/reference/cpp_source/als/IALSTrainer.hpp:170-271) applies
//   A = P + reg_u I + sum_i c_i y_i y_i^T
// to a vector by walking the row's neighbours (1 + max_cg_steps passes over n_u vectors).
// For rows with n_u >> K that is the wrong shape for a GPU: here the K x K matrix
// sum c y y^T comes from the tensor-core Gram (wgram.cu: one pass over the neighbours, no
// per-pass synchronisation) and the very same CG recurrences -- warm start, the
// ||r||^2 <= 1e-20 exits, the !(p.Ap > 0) failure test (:236-263) -- run on the dense
// matrix held in shared memory.  Same linear system, same iteration, different rounding
// of A p (a dense row dot instead of a sum over neighbours).
//
// One 128-thread CTA per heavy row; thread t owns row t of A (129-float stride: both the
// row walk of one thread and the column access of a warp are bank-conflict free).
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;
constexpr int LDA = KP + 1;
constexpr int kThreads = KP;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the 128 threads of the CTA, result in every thread (scratch: 2 x 4 floats,
// alternating so that back-to-back reductions need one barrier each)
__device__ __forceinline__ float block_sum(float v, float *scratch, int &phase) {
  v = warp_sum(v);
  float *s = scratch + 4 * (phase & 1);
  phase++;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  return (s[0] + s[1]) + (s[2] + s[3]);
}

__global__ void __launch_bounds__(kThreads) dense_cg_kernel(DenseSolveArgs d) {
  extern __shared__ __align__(16) float smem[];
  float *A = smem;                 // [128][129]
  float *pv = A + KP * LDA;        // [128] current search direction (16-byte aligned: 16512 floats)
  float *red = pv + KP;            // [8]
  const SolveArgs &a = d.base;
  const int t = threadIdx.x;
  int phase = 0;

  for (int64_t h = blockIdx.x; h < d.n_heavy; h += gridDim.x) {
    const int64_t u = a.order[h];             // CSR row (heavy rows lead the degree-sorted order)
    const int64_t gu = a.row_base + u;        // factor row
    const int j0 = d.heavy_first_job[h], j1 = d.heavy_first_job[h + 1];
    const int64_t nnz = a.indptr[u + 1] - a.indptr[u];
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)nnz, a.nu);  // :117-120
    __syncthreads();  // previous row's readers of A / pv are done
    // S = sum_j W_j + P / 2   (then A = S + S^T = P + sum c y y^T).  32 rows per step: every
    // thread keeps 8 independent 128-bit loads in flight (the partials come from HBM / L2).
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
    __syncthreads();
    // symmetrise in place: the pair (i, t), i < t, belongs to thread t alone
    for (int i = 0; i < t; i++) {
      const float v = A[i * LDA + t] + A[t * LDA + i];
      A[i * LDA + t] = v;
      A[t * LDA + i] = v;
    }
    A[t * LDA + t] = 2.f * A[t * LDA + t] + reg_u;
    pv[t] = x;
    __syncthreads();

    auto matvec = [&]() {  // (A pv)[t]
      const float *row = A + t * LDA;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
      for (int j = 0; j < KP; j += 4) {
        const float4 p4 = *reinterpret_cast<const float4 *>(pv + j);
        acc0 = fmaf(row[j + 0], p4.x, acc0);
        acc1 = fmaf(row[j + 1], p4.y, acc1);
        acc2 = fmaf(row[j + 2], p4.z, acc2);
        acc3 = fmaf(row[j + 3], p4.w, acc3);
      }
      return (acc0 + acc1) + (acc2 + acc3);
    };

    float r = b - matvec();  // r = b - A x   (:216-228)
    float p = r;
    bool failed = false;
    for (int it = 0; it < a.max_cg_steps; it++) {
      const float r2 = block_sum(r * r, red, phase);
      if (r2 <= 1e-20f) break;  // :237-240
      __syncthreads();          // everyone has consumed the previous pv
      pv[t] = p;
      __syncthreads();
      const float Ap = matvec();
      const float den = block_sum(p * Ap, red, phase);
      if (!(den > 0.f) || !isfinite(den)) { failed = true; break; }  // :249-254
      const float alpha = r2 / den;
      x = fmaf(alpha, p, x);
      r = fmaf(-alpha, Ap, r);
      const float r2n = block_sum(r * r, red, phase);
      if (r2n <= 1e-20f) break;  // :258-260
      p = fmaf(r2n / r2, p, r);
    }
    if (failed) {
      if (t == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
      continue;  // the reference throws before writing the row back
    }
    a.target[gu * KP + t] = x;
    for (int pi = 0; pi < a.n_peers; pi++) a.peers[pi][gu * KP + t] = x;
  }
}

}  // namespace

void launch_dense_cg(const DenseSolveArgs &d, cudaStream_t s) {
  if (d.n_heavy <= 0) return;
  const size_t smem = sizeof(float) * (KP * LDA + KP + 8);
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(dense_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  const unsigned grid = (unsigned)std::min<int64_t>(d.n_heavy, (int64_t)kNumSMsB200 * 3);
  dense_cg_kernel<<<grid, kThreads, smem, s>>>(d);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
