// K3 Cholesky row solve (replaces Solver::step_cholesky + BatchedRankUpdater,
// /root/reference/cpp_source/als/IALSTrainer.hpp:273-331, 37-58).
//
// Per row:  A = P + sum c y y^T + reg_u I (upper triangle only, like the
// reference's selfadjointView<Upper>), b = sum (bias + c) y, A = U^T U,
// solve U^T z = b, U x = z.  Empty rows are not special-cased (x = 0 falls out).
//
// v0: one CTA per row.  The upper triangle lives packed in shared memory
// (ld(ld+1)/2 floats: 131.6 KB at K=256, the only form that fits), neighbour
// vectors are staged 32 at a time, the rank update runs on 8x8 register tiles,
// the factorisation is right-looking in shared memory.
#include "common.cuh"

namespace ials {
namespace {

constexpr int kThreads = 256;
constexpr int kStage = 32;  // neighbours staged per rank-update step

__device__ __forceinline__ int packed_index(int i, int j, int ld) {  // j >= i
  return i * ld - (i * (i - 1)) / 2 + (j - i);
}

__global__ void __launch_bounds__(kThreads) cholesky_row_kernel(SolveArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int ld = a.ld, K = a.K;
  const int kd = min(ld, (K + 7) & ~7);   // columns that can be non-zero (ld may pad K to 128)
  const int nt = kd / 8;                  // 8x8 tiles per edge
  const int n_tiles = nt * (nt + 1) / 2;  // upper triangle incl. diagonal
  const int n_packed = ld * (ld + 1) / 2;
  float *A = smem;                             // n_packed (rounded up to 4)
  float *V = A + ((n_packed + 3) & ~3);        // kStage * ld
  float *B = V + kStage * ld;                  // ld
  float *cw = B + ld;                          // kStage confidences
  unsigned short *tile_i = reinterpret_cast<unsigned short *>(cw + kStage);  // n_tiles
  unsigned short *tile_j = tile_i + n_tiles;
  __shared__ long long s_slot;
  __shared__ int s_fail;
  const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
  constexpr int n_warps = kThreads / kWarp;

  for (int t = tid; t < n_tiles; t += kThreads) {  // tile table, once per CTA
    int ti = 0, rem = t;
    while (rem >= nt - ti) { rem -= nt - ti; ti++; }
    tile_i[t] = (unsigned short)ti;
    tile_j[t] = (unsigned short)(ti + rem);
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_slot = (long long)atomicAdd(a.work_counter, 1ull); s_fail = 0; }
    __syncthreads();
    const int64_t slot = s_slot;
    if (slot >= a.n_sched) break;
    const int64_t u = a.order ? (int64_t)a.order[slot] : slot;  // CSR row
    const int64_t gu = a.row_base + u;                          // factor row

    // A <- upper(P), B <- 0                                       (:296-299)
    for (int i = warp; i < kd; i += n_warps)
      for (int j = i + lane; j < kd; j += kWarp) A[packed_index(i, j, ld)] = a.P[i * ld + j];
    for (int k = tid; k < ld; k += kThreads) B[k] = 0.f;
    const int64_t s = a.indptr[u], e = a.indptr[u + 1];
    const int64_t nnz = e - s;

    for (int64_t base = s; base < e; base += kStage) {  // (:301-308)
      const int m = (int)min((int64_t)kStage, e - base);
      __syncthreads();
      for (int t = warp; t < m; t += n_warps) {
        const float *v = a.other + (int64_t)a.indices[base + t] * ld;
        for (int k = lane * 4; k < kd; k += kWarp * 4)
          *reinterpret_cast<float4 *>(&V[t * ld + k]) = *reinterpret_cast<const float4 *>(v + k);
        if (lane == 0) cw[t] = a.data[base + t];
      }
      __syncthreads();
      for (int k = tid; k < kd; k += kThreads) {
        float acc = B[k];
        for (int t = 0; t < m; t++) acc = fmaf(a.bias + cw[t], V[t * ld + k], acc);
        B[k] = acc;
      }
      for (int t = tid; t < n_tiles; t += kThreads) {
        const int i0 = tile_i[t] * 8, j0 = tile_j[t] * 8;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
        for (int q = 0; q < m; q++) {
          float av[8], bv[8];
          const float c = cw[q];
          *reinterpret_cast<float4 *>(&av[0]) = *reinterpret_cast<const float4 *>(&V[q * ld + i0]);
          *reinterpret_cast<float4 *>(&av[4]) = *reinterpret_cast<const float4 *>(&V[q * ld + i0 + 4]);
          *reinterpret_cast<float4 *>(&bv[0]) = *reinterpret_cast<const float4 *>(&V[q * ld + j0]);
          *reinterpret_cast<float4 *>(&bv[4]) = *reinterpret_cast<const float4 *>(&V[q * ld + j0 + 4]);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float ca = c * av[i];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = fmaf(ca, bv[j], acc[i][j]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (i0 + i <= j0 + j) A[packed_index(i0 + i, j0 + j, ld)] += acc[i][j];
      }
    }
    __syncthreads();
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)nnz, a.nu);  // :309-310
    for (int k = tid; k < K; k += kThreads) A[packed_index(k, k, ld)] += reg_u;        // :312-314
    __syncthreads();

    // right-looking upper Cholesky on the leading K x K block      (:316-319)
    bool failed = false;
    for (int i = 0; i < K; i++) {
      const float d2 = A[packed_index(i, i, ld)];
      if (!(d2 > 0.f)) { failed = true; break; }  // uniform: every thread reads the same value
      const float d = sqrtf(d2);
      const float inv = 1.0f / d;
      float *Ui = A + packed_index(i, i, ld);  // row i, element (i, i + t) at Ui[t]
      __syncthreads();
      for (int c = i + 1 + tid; c < K; c += kThreads) Ui[c - i] *= inv;
      if (tid == 0) Ui[0] = d;
      __syncthreads();
      for (int r = i + 1 + warp; r < K; r += n_warps) {
        const float f = Ui[r - i];
        float *Ar = A + packed_index(r, r, ld);
        for (int c = r + lane; c < K; c += kWarp) Ar[c - r] = fmaf(-f, Ui[c - i], Ar[c - r]);
      }
      __syncthreads();
    }
    if (failed) {
      if (tid == 0) atomicExch(&a.err_flags[kErrCholDecomp], 1);
      continue;
    }
    // forward  U^T z = B  (column-oriented: z_i fixed, then eliminate it from later rows)
    for (int i = 0; i < K; i++) {
      __syncthreads();
      const float zi = B[i] / A[packed_index(i, i, ld)];
      __syncthreads();
      if (tid == 0) B[i] = zi;
      const float *Ui = A + packed_index(i, i, ld);
      for (int c = i + 1 + tid; c < K; c += kThreads) B[c] = fmaf(-Ui[c - i], zi, B[c]);
    }
    // backward  U x = z
    for (int i = K - 1; i >= 0; i--) {
      __syncthreads();
      const float xi = B[i] / A[packed_index(i, i, ld)];
      __syncthreads();
      if (tid == 0) B[i] = xi;
      for (int r = tid; r < i; r += kThreads) B[r] = fmaf(-A[packed_index(r, i, ld)], xi, B[r]);
    }
    __syncthreads();
    bool finite = true;
    for (int k = tid; k < K; k += kThreads) finite = finite && isfinite(B[k]);
    if (!finite) s_fail = 1;
    __syncthreads();
    if (s_fail) {  // :320-323
      if (tid == 0) atomicExch(&a.err_flags[kErrCholSolve], 1);
      continue;
    }
    for (int k = tid; k < ld; k += kThreads) {
      const float v = k < K ? B[k] : 0.f;
      a.target[gu * ld + k] = v;
      for (int pi = 0; pi < a.n_peers; pi++) a.peers[pi][gu * ld + k] = v;
    }
  }
}

}  // namespace

size_t cholesky_smem_bytes(int ld) {
  const int nt = ld / 8;
  const int n_tiles = nt * (nt + 1) / 2;
  const int n_packed = ld * (ld + 1) / 2;
  size_t floats = ((n_packed + 3) & ~3) + (size_t)kStage * ld + ld + kStage;
  return floats * sizeof(float) + 2 * sizeof(unsigned short) * n_tiles + 16;
}

void launch_solve_cholesky(const SolveArgs &a, cudaStream_t s) {
  const size_t smem = cholesky_smem_bytes(a.ld);
  if (smem > 227 * 1024) throw NotImplemented("Cholesky solver: n_components > 256 not supported");
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  CUDA_CHECK(cudaFuncSetAttribute(cholesky_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const int per_sm = std::max<int>(1, std::min<int>(8, (int)((227 * 1024) / (smem + 1024))));
  const unsigned grid =
      (unsigned)std::min<int64_t>(std::max<int64_t>(a.n_sched, 1), (int64_t)kNumSMsB200 * per_sm);
  cholesky_row_kernel<<<grid, kThreads, smem, s>>>(a); count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
