// K2 conjugate-gradient row solve (replaces Solver::step_cg,
// /root/reference/cpp_source/als/IALSTrainer.hpp:170-271).
//
// Per row u with CSR neighbours (i, c), Y = other factors, P = alpha0 Y^T Y:
//   A = P + reg_u I + sum c y_i y_i^T,   b = sum (bias + c) y_i
//   x <- warm start;  r = b - A x;  p = r;  <= max_cg_steps CG iterations with
//   the reference's exits (||r||^2 <= 1e-20) and failure test (!(p.Ap > 0)).
// The b pass and the r-init pass are fused:  r = sum (bias + c - c (y.x)) y - P x - reg x
// so a row makes 1 + steps passes over its neighbour vectors instead of 2 + steps.
//
// This file holds the simple reference kernel: one warp per row, neighbour
// vectors re-read from global/L2 on every pass.  It works for every ld (multiple
// of 32, <= 512) and is the fallback / cross-check for the staged kernel in
// cg_staged.cu.
#include "common.cuh"

namespace ials {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__global__ void __launch_bounds__(128) cg_warp_kernel(SolveArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x % kWarp;
  const int warp_in_block = threadIdx.x / kWarp;
  float *ps = smem + warp_in_block * a.ld;  // per-warp broadcast buffer for P*v products
  const int ld = a.ld;

  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.work_counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((int64_t)slot >= a.n_sched) break;
    const int64_t u = a.order ? (int64_t)a.order[slot] : (int64_t)slot;  // CSR row
    const int64_t gu = a.row_base + u;                                   // factor row

    float *xrow = a.target + gu * ld;
    float x[NV], r[NV], p[NV], Ap[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) x[j] = xrow[lane + 32 * j];
    const int64_t s = a.indptr[u], e = a.indptr[u + 1];
    const int64_t nnz = e - s;
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)nnz, a.nu);
    if (nnz == 0) {
#pragma unroll
      for (int j = 0; j < NV; j++) x[j] = 0.f;
    } else {
      // fused b / r-init pass
#pragma unroll
      for (int j = 0; j < NV; j++) r[j] = 0.f;
      for (int64_t jn = s; jn < e; jn++) {
        const float *v = a.other + (int64_t)a.indices[jn] * ld;
        const float c = a.data[jn];
        float vv[NV], d = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) { vv[j] = v[lane + 32 * j]; d = fmaf(vv[j], x[j], d); }
        d = warp_sum(d);
        const float coef = (a.bias + c) - c * d;
#pragma unroll
        for (int j = 0; j < NV; j++) r[j] = fmaf(coef, vv[j], r[j]);
      }
      // r -= P x + reg_u x   (P symmetric: column access is coalesced)
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NV; j++) ps[lane + 32 * j] = x[j];
      __syncwarp();
      {
        float acc[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) acc[j] = 0.f;
        for (int k = 0; k < a.K; k++) {
          const float xk = ps[k];
          const float *Prow = a.P + (int64_t)k * ld;
#pragma unroll
          for (int j = 0; j < NV; j++) acc[j] = fmaf(Prow[lane + 32 * j], xk, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < NV; j++) { r[j] -= acc[j]; r[j] = fmaf(-reg_u, x[j], r[j]); p[j] = r[j]; }
      }
      bool failed = false;
      for (int it = 0; it < a.max_cg_steps; it++) {
        float r2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) r2 = fmaf(r[j], r[j], r2);
        r2 = warp_sum(r2);
        if (r2 <= 1e-20f) break;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NV; j++) ps[lane + 32 * j] = p[j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NV; j++) Ap[j] = 0.f;
        for (int k = 0; k < a.K; k++) {
          const float pk = ps[k];
          const float *Prow = a.P + (int64_t)k * ld;
#pragma unroll
          for (int j = 0; j < NV; j++) Ap[j] = fmaf(Prow[lane + 32 * j], pk, Ap[j]);
        }
#pragma unroll
        for (int j = 0; j < NV; j++) Ap[j] = fmaf(reg_u, p[j], Ap[j]);
        for (int64_t jn = s; jn < e; jn++) {
          const float *v = a.other + (int64_t)a.indices[jn] * ld;
          const float c = a.data[jn];
          float vv[NV], d = 0.f;
#pragma unroll
          for (int j = 0; j < NV; j++) { vv[j] = v[lane + 32 * j]; d = fmaf(vv[j], p[j], d); }
          d = warp_sum(d);
          const float coef = c * d;
#pragma unroll
          for (int j = 0; j < NV; j++) Ap[j] = fmaf(coef, vv[j], Ap[j]);
        }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) den = fmaf(p[j], Ap[j], den);
        den = warp_sum(den);
        if (!(den > 0.f) || !isfinite(den)) { failed = true; break; }
        const float alpha = r2 / den;
        float r2n = 0.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
          x[j] = fmaf(alpha, p[j], x[j]);
          r[j] = fmaf(-alpha, Ap[j], r[j]);
          r2n = fmaf(r[j], r[j], r2n);
        }
        r2n = warp_sum(r2n);
        if (r2n <= 1e-20f) break;
        const float beta = r2n / r2;
#pragma unroll
        for (int j = 0; j < NV; j++) p[j] = fmaf(beta, p[j], r[j]);
      }
      if (failed) {
        if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
        continue;  // the reference throws before writing the row back
      }
    }
#pragma unroll
    for (int j = 0; j < NV; j++) xrow[lane + 32 * j] = x[j];
    for (int pi = 0; pi < a.n_peers; pi++) {
      float *prow = a.peers[pi] + gu * ld;
#pragma unroll
      for (int j = 0; j < NV; j++) prow[lane + 32 * j] = x[j];
    }
  }
}


// ---------------------------------------------------------------------------------------
// Light rows at K padded to 128: one warp per row, P staged ONCE per CTA in shared memory
// (64 KB, 3 CTAs per SM), 128-bit loads everywhere (lane l owns elements 4l .. 4l+3).
// The neighbour vectors are re-read from L2 on each of the 1 + max_cg_steps passes: this
// kernel only gets rows with few neighbours (the heavy ones go to wgram.cu + dense_cg.cu),
// so that traffic is small, and a warp needs no block-level synchronisation at all.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void axpy4(float w, const float4 &v, float4 &acc) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}

constexpr int kLightThreads = 256;

__global__ void __launch_bounds__(kLightThreads, 3) cg_light128_kernel(SolveArgs a) {
  constexpr int KP = 128;
  extern __shared__ __align__(16) float smem[];
  float *Ps = smem;                                           // [128][128]
  float *ps = Ps + KP * KP + (threadIdx.x / kWarp) * KP;      // per-warp vector for P * v
  const int lane = threadIdx.x % kWarp;
  for (int i = threadIdx.x * 4; i < KP * KP; i += kLightThreads * 4)
    *reinterpret_cast<float4 *>(Ps + i) = *reinterpret_cast<const float4 *>(a.P + i);
  __syncthreads();

  // y = P * (vector in ps), elements 4l .. 4l+3 (P symmetric: row k contributes ps[k] * P[k][:])
  auto p_times = [&]() {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < KP; k += 4) {
      const float4 xk = *reinterpret_cast<const float4 *>(ps + k);
      axpy4(xk.x, *reinterpret_cast<const float4 *>(Ps + (k + 0) * KP + 4 * lane), acc);
      axpy4(xk.y, *reinterpret_cast<const float4 *>(Ps + (k + 1) * KP + 4 * lane), acc);
      axpy4(xk.z, *reinterpret_cast<const float4 *>(Ps + (k + 2) * KP + 4 * lane), acc);
      axpy4(xk.w, *reinterpret_cast<const float4 *>(Ps + (k + 3) * KP + 4 * lane), acc);
    }
    return acc;
  };
  // acc += sum_i coef_i(v_i . q) v_i over the row's neighbours; first == true builds the fused
  // b / r-init coefficients (bias + c - c (v . x)), else c (v . p)
  auto neighbour_pass = [&](int64_t s, int64_t e, const float4 &q, bool first, float4 &acc) {
    for (int64_t base = s; base < e; base += kWarp) {
      const int m = (int)min((int64_t)kWarp, e - base);
      int my_i = 0;
      float my_c = 0.f;
      if (lane < m) {
        my_i = a.indices[base + lane];
        my_c = a.data[base + lane];
      }
      int t = 0;
      for (; t + 1 < m; t += 2) {  // two neighbours in flight
        const int i0 = __shfl_sync(0xffffffffu, my_i, t), i1 = __shfl_sync(0xffffffffu, my_i, t + 1);
        const float c0 = __shfl_sync(0xffffffffu, my_c, t), c1 = __shfl_sync(0xffffffffu, my_c, t + 1);
        const float4 v0 = *reinterpret_cast<const float4 *>(a.other + (size_t)i0 * KP + 4 * lane);
        const float4 v1 = *reinterpret_cast<const float4 *>(a.other + (size_t)i1 * KP + 4 * lane);
        float d0 = dot4(v0, q), d1 = dot4(v1, q);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d0 += __shfl_xor_sync(0xffffffffu, d0, o);
          d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        }
        axpy4(first ? (a.bias + c0) - c0 * d0 : c0 * d0, v0, acc);
        axpy4(first ? (a.bias + c1) - c1 * d1 : c1 * d1, v1, acc);
      }
      if (t < m) {
        const int i0 = __shfl_sync(0xffffffffu, my_i, t);
        const float c0 = __shfl_sync(0xffffffffu, my_c, t);
        const float4 v0 = *reinterpret_cast<const float4 *>(a.other + (size_t)i0 * KP + 4 * lane);
        const float d0 = warp_sum(dot4(v0, q));
        axpy4(first ? (a.bias + c0) - c0 * d0 : c0 * d0, v0, acc);
      }
    }
  };

  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.work_counter, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((int64_t)slot >= a.n_sched) break;
    const int64_t u = a.order ? (int64_t)a.order[slot] : (int64_t)slot;  // CSR row
    const int64_t gu = a.row_base + u;                                   // factor row
    float4 *xrow = reinterpret_cast<float4 *>(a.target + gu * KP) + lane;
    const int64_t s = a.indptr[u], e = a.indptr[u + 1];
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    bool failed = false;
    if (e > s) {  // rows without interactions become zero (IALSTrainer.hpp:207-210)
      x = *xrow;
      const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)(e - s), a.nu);
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      neighbour_pass(s, e, x, true, r);
      __syncwarp();
      *reinterpret_cast<float4 *>(ps + 4 * lane) = x;
      __syncwarp();
      {
        const float4 px = p_times();
        r.x -= px.x; r.y -= px.y; r.z -= px.z; r.w -= px.w;
        axpy4(-reg_u, x, r);
      }
      float4 p = r;
      for (int it = 0; it < a.max_cg_steps; it++) {
        const float r2 = warp_sum(dot4(r, r));
        if (r2 <= 1e-20f) break;
        __syncwarp();
        *reinterpret_cast<float4 *>(ps + 4 * lane) = p;
        __syncwarp();
        float4 Ap = p_times();
        axpy4(reg_u, p, Ap);
        neighbour_pass(s, e, p, false, Ap);
        const float den = warp_sum(dot4(p, Ap));
        if (!(den > 0.f) || !isfinite(den)) { failed = true; break; }
        const float alpha = r2 / den;
        axpy4(alpha, p, x);
        axpy4(-alpha, Ap, r);
        const float r2n = warp_sum(dot4(r, r));
        if (r2n <= 1e-20f) break;
        const float beta = r2n / r2;
        p = make_float4(fmaf(beta, p.x, r.x), fmaf(beta, p.y, r.y), fmaf(beta, p.z, r.z), fmaf(beta, p.w, r.w));
      }
    }
    if (failed) {
      if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
      continue;  // the reference throws before writing the row back
    }
    *xrow = x;
    for (int pi = 0; pi < a.n_peers; pi++) reinterpret_cast<float4 *>(a.peers[pi] + gu * KP)[lane] = x;
  }
}

template <int NV>
void launch_nv(const SolveArgs &a, cudaStream_t s) {
  const int threads = 128;
  const size_t smem = sizeof(float) * (threads / kWarp) * a.ld;
  const int64_t warps_needed = std::max<int64_t>(a.n_sched, 1);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(warps_needed, threads / kWarp),
                                                    (int64_t)kNumSMsB200 * 16);
  cg_warp_kernel<NV><<<grid, threads, smem, s>>>(a); count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

// One warp per row with P in shared memory (ld == 128): the light-row kernel.
void launch_solve_cg_light128(const SolveArgs &a, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  const size_t smem = sizeof(float) * (128 * 128 + (kLightThreads / kWarp) * 128);
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(cg_light128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  const int64_t ctas_needed = ceil_div(a.n_sched, (int64_t)(kLightThreads / kWarp));
  const unsigned grid = (unsigned)std::min<int64_t>(ctas_needed, (int64_t)kNumSMsB200 * 3);
  cg_light128_kernel<<<grid, kLightThreads, smem, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

void launch_solve_cg_simple(const SolveArgs &a, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  switch (a.ld / 32) {
    case 1: launch_nv<1>(a, s); break;
    case 2: launch_nv<2>(a, s); break;
    case 3: launch_nv<3>(a, s); break;
    case 4: launch_nv<4>(a, s); break;
    case 5: launch_nv<5>(a, s); break;
    case 6: launch_nv<6>(a, s); break;
    case 7: launch_nv<7>(a, s); break;
    case 8: launch_nv<8>(a, s); break;
    case 10: launch_nv<10>(a, s); break;
    case 12: launch_nv<12>(a, s); break;
    case 16: launch_nv<16>(a, s); break;
    default:
      throw NotImplemented("CG solver: n_components must pad to 32..256, 320, 384 or 512");
  }
}

}  // namespace ials
