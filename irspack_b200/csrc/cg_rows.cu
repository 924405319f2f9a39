// K2 conjugate-gradient row solve for the light rows (K padded to 128): warp-per-row batches.
// Replaces Solver::step_cg, /root/reference/cpp_source/als/IALSTrainer.hpp:170-271;
// same arithmetic as cg.cu (fused b / r-init pass, the reference's exits and failure test).
//
// ncu on cg_light128_kernel (profiles/r01e_*) shows the warp-per-row solve bound by the
// LSU / L1TEX data pipe (81 % busy, 1 wavefront per clock and SM), which serves the gathered
// vectors (4 wavefronts each), every shuffle (1 each) and the reads of P in shared memory
// (544 per row and pass).  This kernel spends far fewer wavefronts on the same arithmetic:
//   * a neighbour vector is owned by an 8-lane group (16 floats per lane): one warp-wide
//     LDG.128 fetches a quarter of FOUR neighbours, the dot product needs 3 shuffles per four
//     neighbours instead of 5 per neighbour, and each group reads its own (index, confidence)
//     instead of receiving it by shuffle: 5.25 wavefronts per neighbour and pass instead of 11;
//   * a warp owns R consecutive rows of the degree-sorted schedule; the two warps of a PAIR
//     multiply P with their 2R search directions in one sweep over P, half of P's rows each, and
//     exchange the partial sums through shared memory (two named 64-thread barriers per pass):
//     256 / R + 40 wavefronts per row and pass where one warp sweeping all of P for its own R rows
//     needed 512 / R + 32;
//   * x, r and p of the R rows live in a per-warp slab of shared memory between the phases
//     (lane-private words for x and r), so the register budget holds two neighbour batches
//     in flight (8 vectors per warp) on top of the accumulators.
// No block-level synchronisation after the prologue (only the pairs' barriers); one 512-thread
// CTA per SM.
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;
constexpr int kRowsWarps = 16;
constexpr int kRowsThreads = kRowsWarps * kWarp;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                     __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
// Packed FP32 (sm_100: fma.rn.f32x2 -> SASS FFMA2): two fused multiply-adds per issue slot.  The
// kernel is co-limited by the L1TEX wavefront pipe and by instruction issue (ncu, r01e: issue
// active 56 %, half of the instructions FFMA), so the gather loop and the sweep over P use it.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long *>(&d))
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
  return d;
}
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
// acc (two partial sums) += a . b, element pairs (x, y) and (z, w)
__device__ __forceinline__ float2 dot4p(float4 a, float4 b, float2 acc) {
  acc = fma2(lo2(a), lo2(b), acc);
  return fma2(hi2(a), hi2(b), acc);
}
__device__ __forceinline__ void axpy4(float w, float4 v, float4 &acc) {
  const float2 ww = make_float2(w, w);
  const float2 l = fma2(ww, lo2(v), lo2(acc)), h = fma2(ww, hi2(v), hi2(acc));
  acc = make_float4(l.x, l.y, h.x, h.y);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// shared memory: P, then per PAIR of warps the slab  X | R | V  of its 2R rows (the two warps' rows
// side by side: both read all 2R search directions in the sweep), the partial products each warp
// computes for its partner's rows, and the pair's hand-over words
template <int R>
constexpr size_t rows_pair_floats() {
  return (size_t)3 * 2 * R * KP + (size_t)2 * R * KP + 8;
}
template <int R>
constexpr size_t rows_smem_bytes() {
  return sizeof(float) * ((size_t)KP * KP + (size_t)(kRowsWarps / 2) * rows_pair_floats<R>());
}
// the two warps of a pair meet at a named barrier (ids 1 .. 8; 0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}
template <int R>
__global__ void __launch_bounds__(kRowsThreads, 1) cg_rows_kernel(SolveArgs a) {
  extern __shared__ __align__(16) float smem[];
  float *Ps = smem;  // [128][128]
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int pair = warp >> 1, half = warp & 1;  // half: which 64 rows of P this warp sweeps
  const int g = lane >> 3, l8 = lane & 7;       // 8-lane group, lane within the group
  float *pslab = Ps + KP * KP + (size_t)pair * rows_pair_floats<R>();
  float *Xp = pslab;                   // [2R][128] x      (lane-private words 4*lane .. 4*lane+3)
  float *Rp = pslab + 2 * R * KP;      // [2R][128] r      (lane-private)
  float *Vp = pslab + 4 * R * KP;      // [2R][128] vector to multiply: x in pass 0, then p
  float *PPp = pslab + 6 * R * KP;     // [2][R][128] PPp[w]: warp w's partial P.v for its PARTNER's rows
  volatile int *s_any = reinterpret_cast<volatile int *>(pslab + 8 * R * KP);             // [2]
  volatile unsigned long long *s_slot =
      reinterpret_cast<volatile unsigned long long *>(pslab + 8 * R * KP + 2);            // [2] (by trip parity)
  float *Xs = Xp + half * R * KP, *Rs = Rp + half * R * KP, *Vs = Vp + half * R * KP;   // this warp's own rows
  for (int i = threadIdx.x * 4; i < KP * KP; i += kRowsThreads * 4) st4(Ps + i, ld4(a.P + i));
  __syncthreads();

  for (unsigned trip = 0;; trip++) {
    // one cursor step per pair: 2R consecutive rows of the schedule, R for each warp
    if (half == 0 && lane == 0) s_slot[trip & 1] = atomicAdd(a.work_counter, (unsigned long long)(2 * R));
    pair_sync(pair);
    const unsigned long long slot00 = s_slot[trip & 1];
    if ((int64_t)slot00 >= a.n_sched) break;
    const unsigned long long slot0 = slot00 + (unsigned long long)(half * R);

    int64_t gu[R], s[R];
    int n[R];
    float reg_u[R], r2[R];
    bool active[R], failed[R], exists[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int64_t slot = (int64_t)slot0 + r;
      exists[r] = slot < a.n_sched;
      const int64_t u = exists[r] ? (a.order ? (int64_t)a.order[slot] : slot) : 0;
      gu[r] = a.row_base + u;
      s[r] = exists[r] ? a.indptr[u] : 0;
      n[r] = exists[r] ? (int)(a.indptr[u + 1] - s[r]) : 0;
      reg_u[r] = a.reg * powf(a.alpha0 * (float)a.n_other + (float)n[r], a.nu);
      r2[r] = 0.f;
      active[r] = n[r] > 0;
      failed[r] = false;
      if (a.ready_flags != nullptr && exists[r]) {  // the row may still be on its way (an empty one too:
                                                    // its zero must not be overwritten by a late chunk)
        if (lane == 0) wait_row_ready(a, gu[r]);
        __syncwarp();
      }
      // rows without interactions become zero (IALSTrainer.hpp:207-210)
      const float4 x0 = active[r] ? ld4(a.target + gu[r] * KP + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      st4(Xs + r * KP + 4 * lane, x0);
      st4(Vs + r * KP + 4 * lane, x0);
    }
    {
      bool mine = false;
#pragma unroll
      for (int r = 0; r < R; r++) mine |= active[r];
      if (lane == 0) s_any[half] = mine ? 1 : 0;
    }
    __syncwarp();

    for (int pass = 0; pass <= a.max_cg_steps; pass++) {
      pair_sync(pair);  // the 2R vectors of this pass and both warps' flags are in shared memory
      if ((s_any[0] | s_any[1]) == 0) break;  // the same for both warps of the pair

      // ---- P * V for the pair's 2R rows in ONE sweep over P, half of P's rows per warp (the sweep
      // was a fifth of this kernel's shared-memory wavefronts with a full sweep per warp and two rows:
      // r02x, 3.08 -> 2.81 ms per user half-epoch with half of it skipped) ----
      float4 Pp[R];
      {
        float4 acc[2 * R];
#pragma unroll
        for (int r = 0; r < 2 * R; r++) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int kb = half * (KP / 2);
#pragma unroll 2
        for (int k = kb; k < kb + KP / 2; k += 4) {
          const float4 p0 = ld4(Ps + (k + 0) * KP + 4 * lane), p1 = ld4(Ps + (k + 1) * KP + 4 * lane);
          const float4 p2 = ld4(Ps + (k + 2) * KP + 4 * lane), p3 = ld4(Ps + (k + 3) * KP + 4 * lane);
#pragma unroll
          for (int r = 0; r < 2 * R; r++) {
            const float4 vk = ld4(Vp + r * KP + k);
            axpy4(vk.x, p0, acc[r]);
            axpy4(vk.y, p1, acc[r]);
            axpy4(vk.z, p2, acc[r]);
            axpy4(vk.w, p3, acc[r]);
          }
        }
        // hand the partner its rows' partial sums, take mine from it
#pragma unroll
        for (int r = 0; r < R; r++)  // (selects, not a runtime index: the accumulators stay in registers)
          st4(PPp + (half * R + r) * KP + 4 * lane, half ? acc[r] : acc[R + r]);
        pair_sync(pair);
#pragma unroll
        for (int r = 0; r < R; r++)
          Pp[r] = add4(half ? acc[R + r] : acc[r], ld4(PPp + ((1 - half) * R + r) * KP + 4 * lane));
      }

#pragma unroll
      for (int r = 0; r < R; r++) {
        if (!active[r]) continue;  // warp-uniform
        // ---- neighbour pass: acc = sum_t coef_t v_t,  coef = bias + c - c (v.x) in pass 0,
        //      c (v.p) afterwards; four neighbours per load instruction, eight in flight ----
        float4 q[4], acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          q[i] = ld4(Vs + r * KP + i * 32 + l8 * 4);
          acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int32_t *idxp = a.indices + s[r];
        const float *cp = a.data + s[r];
        const float *ybase = a.other + l8 * 4;
        const int nr = n[r];
        // (index, confidence) of the next batch are fetched one batch ahead
        int ia = idxp[g < nr ? g : 0], ib = idxp[4 + g < nr ? 4 + g : 0];
        float ca = g < nr ? cp[g] : 0.f, cb = 4 + g < nr ? cp[4 + g] : 0.f;
        for (int tb = 0; tb < nr; tb += 8) {
          const float *ya = ybase + (size_t)ia * KP, *yb = ybase + (size_t)ib * KP;
          const bool va = tb + g < nr, vb = tb + 4 + g < nr;
          const float c0 = ca, c1 = cb;
          float4 v0[4], v1[4];
#pragma unroll
          for (int i = 0; i < 4; i++) v0[i] = ldg4(ya + i * 32);
#pragma unroll
          for (int i = 0; i < 4; i++) v1[i] = ldg4(yb + i * 32);
          {
            const int ta = tb + 8 + g, tb2 = tb + 12 + g;
            ia = idxp[ta < nr ? ta : 0];
            ib = idxp[tb2 < nr ? tb2 : 0];
            ca = ta < nr ? cp[ta] : 0.f;
            cb = tb2 < nr ? cp[tb2] : 0.f;
          }
          const float2 z2 = make_float2(0.f, 0.f);
          float2 D0 = dot4p(v0[0], q[0], z2), E0 = dot4p(v0[1], q[1], z2);
          float2 D1 = dot4p(v1[0], q[0], z2), E1 = dot4p(v1[1], q[1], z2);
          D0 = dot4p(v0[2], q[2], D0);
          E0 = dot4p(v0[3], q[3], E0);
          D1 = dot4p(v1[2], q[2], D1);
          E1 = dot4p(v1[3], q[3], E1);
          float d0 = (D0.x + D0.y) + (E0.x + E0.y);
          float d1 = (D1.x + D1.y) + (E1.x + E1.y);
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o);
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
          }
          float w0 = pass == 0 ? (a.bias + c0) - c0 * d0 : c0 * d0;
          float w1 = pass == 0 ? (a.bias + c1) - c1 * d1 : c1 * d1;
          w0 = va ? w0 : 0.f;
          w1 = vb ? w1 : 0.f;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            axpy4(w0, v0[i], acc[i]);
            axpy4(w1, v1[i], acc[i]);
          }
        }
        // reduce-scatter over the 4 groups: lane ends up with elements [4*lane, 4*lane+4)
        const bool hi = (g & 2) != 0, odd = (g & 1) != 0;
        float4 k0 = hi ? acc[2] : acc[0], k1 = hi ? acc[3] : acc[1];
        const float4 s0 = hi ? acc[0] : acc[2], s1 = hi ? acc[1] : acc[3];
        k0 = add4(k0, shfl_xor4(s0, 16));
        k1 = add4(k1, shfl_xor4(s1, 16));
        float4 mine = odd ? k1 : k0;
        mine = add4(mine, shfl_xor4(odd ? k0 : k1, 8));

        // ---- CG algebra of row r (flat layout) ----
        float4 x = ld4(Xs + r * KP + 4 * lane);
        float4 p;
        if (pass == 0) {
          float4 rv = make_float4(mine.x - Pp[r].x, mine.y - Pp[r].y, mine.z - Pp[r].z, mine.w - Pp[r].w);
          axpy4(-reg_u[r], x, rv);
          p = rv;
          r2[r] = warp_sum(dot4(rv, rv, 0.f));
          st4(Rs + r * KP + 4 * lane, rv);
          if (r2[r] <= 1e-20f) active[r] = false;  // IALSTrainer.hpp:237-240
        } else {
          p = ld4(Vs + r * KP + 4 * lane);
          float4 rv = ld4(Rs + r * KP + 4 * lane);
          float4 Ap = add4(mine, Pp[r]);
          axpy4(reg_u[r], p, Ap);
          const float den = warp_sum(dot4(p, Ap, 0.f));
          if (!(den > 0.f) || !isfinite(den)) {  // :249-254
            failed[r] = true;
            active[r] = false;
            continue;
          }
          const float alpha = r2[r] / den;
          axpy4(alpha, p, x);
          axpy4(-alpha, Ap, rv);
          st4(Xs + r * KP + 4 * lane, x);
          st4(Rs + r * KP + 4 * lane, rv);
          const float r2n = warp_sum(dot4(rv, rv, 0.f));
          if (r2n <= 1e-20f) {  // :258-260
            active[r] = false;
            continue;
          }
          const float beta = r2n / r2[r];
          p = make_float4(fmaf(beta, p.x, rv.x), fmaf(beta, p.y, rv.y), fmaf(beta, p.z, rv.z),
                          fmaf(beta, p.w, rv.w));
          r2[r] = r2n;
        }
        __syncwarp();  // every lane is done reading V[r] (group layout, flat)
        st4(Vs + r * KP + 4 * lane, p);
      }
      {  // this warp's flag for the pair's next pass (the pair barrier at its top publishes it and V)
        bool mine = false;
#pragma unroll
        for (int r = 0; r < R; r++) mine |= active[r];
        if (lane == 0) s_any[half] = mine ? 1 : 0;
      }
      __syncwarp();
    }

#pragma unroll
    for (int r = 0; r < R; r++) {
      if (!exists[r]) continue;
      if (failed[r]) {  // the reference throws before writing the row back
        if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
        continue;
      }
      const float4 x = ld4(Xs + r * KP + 4 * lane);
      st4(a.target + gu[r] * KP + 4 * lane, x);
      for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + gu[r] * KP + 4 * lane, x);
    }
    __syncwarp();
  }
}

constexpr int kRowsPerWarp = 2;  // R = 1 and R = 4 measured slower (profiles/r01i_ab_pipe.md, r02a)

}  // namespace

// Light rows, ld == 128: two rows per warp and sweep over P.
void launch_solve_cg_rows(const SolveArgs &a, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  if (a.ld != KP) throw NotImplemented("cg_rows kernel: ld must be 128");
  constexpr int R = kRowsPerWarp;
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(cg_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)rows_smem_bytes<R>()));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ctas = ceil_div(a.n_sched, (int64_t)kRowsWarps * R);
  const unsigned grid = (unsigned)std::min<int64_t>(ctas, sms);
  cg_rows_kernel<R><<<grid, kRowsThreads, rows_smem_bytes<R>(), s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
