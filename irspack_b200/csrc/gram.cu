// K1 Gram:  P = alpha0 * Y^T Y     (replaces Solver::prepare_p,
// /root/reference/cpp_source/als/IALSTrainer.hpp:78-115).
//
// v0: FP32 SIMT register-tiled kernel.  Each CTA walks a strided set of 16-row
// chunks of Y, keeps a 128x128 tile of the K x K result in registers (8x8 per
// thread), and adds it to a global scratch with atomics; a finalize kernel
// scales by alpha0 and mirrors the upper triangle so P is exactly symmetric.
#include "common.cuh"

namespace ials {
namespace {

constexpr int kTile = 128;   // output tile edge
constexpr int kChunk = 16;   // Y rows staged per step
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
gram_tile_kernel(const float *__restrict__ Y, int64_t row_begin, int64_t row_end, int ld,
                 float *__restrict__ S) {
  const int ta = blockIdx.y, tb = blockIdx.z;
  if (ta > tb) return;  // lower tiles are mirrored by the finalize kernel
  __shared__ __align__(16) float Ya[kChunk][kTile];
  __shared__ __align__(16) float Yb[kChunk][kTile];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const int64_t n_chunks = ceil_div(row_end - row_begin, (int64_t)kChunk);
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int64_t r0 = row_begin + c * kChunk;
    // stage 16 x 128 floats of each operand: 512 float4, 2 per thread
#pragma unroll
    for (int it = 0; it < 2; it++) {
      const int f = tid + it * kThreads;  // float4 index
      const int rr = f / (kTile / 4), cc = (f % (kTile / 4)) * 4;
      const int64_t r = r0 + rr;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (r < row_end) {
        const int ca = ta * kTile + cc, cb = tb * kTile + cc;
        if (ca < ld) va = *reinterpret_cast<const float4 *>(Y + r * ld + ca);
        if (cb < ld) vb = *reinterpret_cast<const float4 *>(Y + r * ld + cb);
      }
      *reinterpret_cast<float4 *>(&Ya[rr][cc]) = va;
      *reinterpret_cast<float4 *>(&Yb[rr][cc]) = vb;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kChunk; rr++) {
      float a[8], b[8];
      *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&Ya[rr][ty * 4]);
      *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&Ya[rr][64 + ty * 4]);
      *reinterpret_cast<float4 *>(&b[0]) = *reinterpret_cast<const float4 *>(&Yb[rr][tx * 4]);
      *reinterpret_cast<float4 *>(&b[4]) = *reinterpret_cast<const float4 *>(&Yb[rr][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int a = ta * kTile + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (a >= ld) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int b = tb * kTile + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (b < ld && a <= b) atomicAdd(&S[a * ld + b], acc[i][j]);
    }
  }
}

__global__ void gram_finalize_kernel(const float *__restrict__ S, int ld, float alpha0,
                                     float *__restrict__ P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ld * ld) {
    int a = i / ld, b = i % ld;
    int lo = a < b ? a : b, hi = a < b ? b : a;
    P[i] = alpha0 * S[lo * ld + hi];  // "P *= alpha0" after the sum, :113
  }
}

}  // namespace

void launch_gram(const float *Y, int64_t row_begin, int64_t row_end, int ld, float alpha0,
                 float *scratch, float *P, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(float) * ld * ld, s));
  const int64_t n = row_end - row_begin;
  if (n > 0) {
    const int nt = (int)ceil_div(ld, kTile);
    const int64_t n_chunks = ceil_div(n, (int64_t)kChunk);
    // two resident CTAs per SM per tile pair is plenty; small inputs get fewer
    const unsigned gx = (unsigned)std::min<int64_t>(n_chunks, 2 * kNumSMsB200);
    gram_tile_kernel<<<dim3(gx, nt, nt), kThreads, 0, s>>>(Y, row_begin, row_end, ld, scratch); count_launch();
    CUDA_CHECK(cudaGetLastError());
  }
  gram_finalize_kernel<<<(unsigned)ceil_div((int64_t)ld * ld, 256), 256, 0, s>>>(scratch, ld,
                                                                                 alpha0, P); count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
