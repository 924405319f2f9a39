// K2 conjugate-gradient row solve, staged kernel (K padded to 128).
// Replaces Solver::step_cg, /root/reference/cpp_source/als/IALSTrainer.hpp:170-271;
// same arithmetic as cg.cu (fused b / r-init pass, the reference's exits and
// failure test), re-cut for the SM:
//
//   * persistent grid, one 544-thread CTA per SM: 16 consumer warps + 1 producer warp;
//   * the producer warp takes rows off the degree-sorted schedule, reads their CSR
//     slice and issues one TMA bulk copy (cp.async.bulk, 512 B) per neighbour
//     vector into a two-buffer shared-memory ring (2 x 192 vectors = 192 KB),
//     completion tracked by mbarriers; it also stages the confidences and the
//     row's warm-start vector, and it runs ahead of the consumers, so the next
//     row's gather overlaps the current row's arithmetic;
//   * rows of <= 384 neighbours stay resident for all 1 + max_cg_steps passes
//     (one HBM/L2 read per neighbour vector); longer rows stream through the
//     ring once per pass;
//   * each neighbour vector is owned by an 8-lane group (16 floats per lane, read
//     with conflict-free LDS.128): dot product = 16 FMA + 3 shuffles, update =
//     16 FMA; partial sums are reduce-scattered across the 4 groups of a warp
//     (12 shuffles) and across warps through shared memory;
//   * P (K x K) lives in REGISTERS, 32 per consumer thread (warp w owns rows
//     8w..8w+7, lane l columns 4l..4l+3), so the P*p product costs no shared-memory
//     or L2 traffic at all;
//   * consumer warp 0 keeps x, r, p in registers and does the scalar CG algebra.
#include "common.cuh"

namespace ials {
namespace {

constexpr int kConsumerWarps = 16;
constexpr int kThreads = (kConsumerWarps + 1) * kWarp;  // 544
constexpr int kRingBytesPerBuffer = 98304;              // 96 KB
constexpr int kDescDepth = 4;
constexpr bool kGatherTma = false;  // true: one 512-byte TMA bulk copy per vector (slow: ~10 copies/us/SM)

struct RowDesc {
  long long u;
  int n;
  int pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  // try_wait suspends the thread for a bounded time per call; the counter only turns a
  // protocol bug into a trapped launch (error) instead of a hung GPU.
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Ampere-style 16-byte asynchronous copy global -> shared (SASS: LDGSTS), L1 bypassed.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
// The mbarrier gets one arrival from this thread once all its prior cp.async copies landed.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * kWarp) : "memory");
}
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
__device__ __forceinline__ void axpy4(float w, float4 v, float4 &acc) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// KP = 128 only (NV4 = 4 float4 per lane and neighbour).
__global__ void __launch_bounds__(640, 1) cg_staged_kernel_k128(SolveArgs a) {
  constexpr int KP = 128;
  constexpr int NV4 = 4;
  constexpr int CHUNK = kRingBytesPerBuffer / (KP * 4);  // 192 neighbour vectors per buffer
  constexpr int KR = KP / kConsumerWarps;                // 8 rows of P per consumer warp
  constexpr int SLOTS = kConsumerWarps * 4;              // 64 neighbours in flight per round

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *vec = reinterpret_cast<float *>(smem_raw);        // [2][CHUNK][KP]
  float *coef = vec + 2 * CHUNK * KP;                      // [2][CHUNK]
  float *partial = coef + 2 * CHUNK;                       // [2][16][KP]
  float *pvec = partial + 2 * kConsumerWarps * KP;         // [KP]   current p
  float *xrow = pvec + KP;                                 // [4][KP] warm-start rows
  RowDesc *desc = reinterpret_cast<RowDesc *>(xrow + kDescDepth * KP);  // [4]
  int *flags = reinterpret_cast<int *>(desc + kDescDepth);              // [4]: 0 = done
  uint64_t *bars = reinterpret_cast<uint64_t *>(flags + 4);
  uint64_t *full = bars, *empty = bars + 2, *rowfull = bars + 4;  // [2], [2], [4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = a.ld;  // == KP

  if (tid == 0) {
    for (int b = 0; b < 2; b++) {
      mbar_init(&full[b], kGatherTma ? 1 : kWarp + 1);
      mbar_init(&empty[b], kConsumerWarps);
    }
    for (int i = 0; i < kDescDepth; i++) mbar_init(&rowfull[i], 1);
    flags[0] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int passes = 1 + a.max_cg_steps;

  if (warp == kConsumerWarps) {
    // ============================ PRODUCER WARP ============================
    unsigned long long loads = 0;  // chunk loads issued so far; load L uses buffer L & 1
    unsigned long long q = 0;      // rows published so far
    for (;;) {
      unsigned long long slot = 0;
      if (lane == 0) slot = atomicAdd(a.work_counter, 1ull);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if ((long long)slot >= a.n_sched) {
        if (lane == 0) {
          desc[q % kDescDepth].n = -1;
          mbar_arrive(&rowfull[q % kDescDepth]);
        }
        break;
      }
      const long long u = a.order ? (long long)a.order[slot] : (long long)slot;  // CSR row
      const long long s = a.indptr[u];
      const int n = (int)(a.indptr[u + 1] - s);
      if (n == 0) continue;  // empty rows are zero-filled by zero_rows_kernel
      if (lane == 0) {
        desc[q % kDescDepth].u = a.row_base + u;  // factor row
        desc[q % kDescDepth].n = n;
        mbar_arrive(&rowfull[q % kDescDepth]);  // release: the descriptor is visible
      }
      const int nchunks = (n + CHUNK - 1) / CHUNK;
      const bool resident = nchunks <= 2;
      const int total = resident ? nchunks : passes * nchunks;
      for (int li = 0; li < total; li++) {
        const int j = li % nchunks;
        const unsigned long long id = loads + li;
        const int b = (int)(id & 1);
        if (id >= 2) mbar_wait(&empty[b], (uint32_t)(((id >> 1) - 1) & 1));
        const int m = min(CHUNK, n - j * CHUNK);
        const long long base = s + (long long)j * CHUNK;
        // stage confidences first, so that they are visible when the barrier completes
        int idx[CHUNK / kWarp];
#pragma unroll
        for (int i = 0; i < CHUNK / kWarp; i++) {
          const int t = lane + i * kWarp;
          idx[i] = 0;
          if (t < m) {
            idx[i] = a.indices[base + t];
            coef[b * CHUNK + t] = a.data[base + t];
          }
        }
        __syncwarp();
        if (kGatherTma) {
          if (lane == 0) {
            const uint32_t bytes = (uint32_t)m * KP * 4 + (li == 0 ? KP * 4 : 0);
            mbar_arrive_expect_tx(&full[b], bytes);
            if (li == 0)
              bulk_copy_g2s(xrow + (q % kDescDepth) * KP, a.target + (a.row_base + u) * ld, KP * 4, &full[b]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < CHUNK / kWarp; i++) {
            const int t = lane + i * kWarp;
            if (t < m)
              bulk_copy_g2s(vec + ((size_t)b * CHUNK + t) * KP, a.other + (long long)idx[i] * ld,
                            KP * 4, &full[b]);
          }
        } else {
          // one warp-wide LDGSTS (32 x 16 B) per neighbour vector; the column index is
          // broadcast from the lane that read it
          if (li == 0) cp_async16(xrow + (q % kDescDepth) * KP + 4 * lane, a.target + (a.row_base + u) * ld + 4 * lane);
          float *dst = vec + (size_t)b * CHUNK * KP + 4 * lane;
          const float *src = a.other + 4 * lane;
#pragma unroll
          for (int i = 0; i < CHUNK / kWarp; i++) {
            if (i * kWarp < m) {
              const int cnt = min(kWarp, m - i * kWarp);
#pragma unroll 8
              for (int tt = 0; tt < kWarp; tt++) {
                const int col = __shfl_sync(0xffffffffu, idx[i], tt);
                if (tt < cnt) cp_async16(dst + (size_t)(i * kWarp + tt) * KP, src + (long long)col * ld);
              }
            }
          }
          cp_async_mbar_arrive_noinc(&full[b]);  // 32 arrivals, each when its lane's copies landed
          __syncwarp();                          // orders every lane's coef stores before ...
          if (lane == 0) mbar_arrive(&full[b]);  // ... the releasing arrival number 33
        }
      }
      loads += total;
      q++;
    }
    return;
  }

  // ============================ CONSUMER WARPS ============================
  const int g = lane >> 3, l8 = lane & 7;  // 8-lane group, lane within the group
  // P rows [warp*8, warp*8+8), columns [4*lane, 4*lane+4): 32 registers
  float4 Preg[KR];
#pragma unroll
  for (int kk = 0; kk < KR; kk++)
    Preg[kk] = *reinterpret_cast<const float4 *>(a.P + (size_t)(warp * KR + kk) * ld + 4 * lane);

  unsigned long long loads = 0, q = 0, gpass = 0;
  for (;;) {
    mbar_wait(&rowfull[q % kDescDepth], (uint32_t)((q / kDescDepth) & 1));
    const int n = desc[q % kDescDepth].n;
    if (n < 0) break;
    const long long u = desc[q % kDescDepth].u;
    const float *x0 = xrow + (q % kDescDepth) * KP;
    const int nchunks = (n + CHUNK - 1) / CHUNK;
    const bool resident = nchunks <= 2;
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)n, a.nu);

    // master state (consumer warp 0): elements [4*lane, 4*lane+4)
    float4 mx = make_float4(0.f, 0.f, 0.f, 0.f), mr = mx, mp = mx;
    float r2 = 0.f;
    bool failed = false;
    int pass = 0;
    if (n > 0) {
      for (; pass < passes; pass++) {
        bool active = true;
        if (pass > 0) {
          consumer_sync();  // new p (and the done flag) published by the master
          active = flags[0] == 0;
          if (!active && resident) break;
        }
        const float *pv_src = pass == 0 ? x0 : pvec;
        float4 acc[NV4];
#pragma unroll
        for (int i = 0; i < NV4; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pv[NV4];
        for (int j = 0; j < nchunks; j++) {
          const unsigned long long id = resident ? loads + j : loads + (unsigned long long)pass * nchunks + j;
          const int b = (int)(id & 1);
          if (!resident || pass == 0) mbar_wait(&full[b], (uint32_t)((id >> 1) & 1));
          if (j == 0) {  // x0 arrives with the row's first chunk
#pragma unroll
            for (int i = 0; i < NV4; i++)
              pv[i] = *reinterpret_cast<const float4 *>(pv_src + i * 32 + l8 * 4);
          }
          if (active) {
            const int m = min(CHUNK, n - j * CHUNK);
            const float4 *vb = reinterpret_cast<const float4 *>(vec + (size_t)b * CHUNK * KP);
            const float *cb = coef + b * CHUNK;
            // the trip count must be warp-uniform (full-mask shuffles inside): iterate on
            // the warp's first slot and predicate the groups that fall off the end
            for (int tb = warp * 4; tb < m; tb += SLOTS) {
              const int t = tb + g;
              const bool valid = t < m;
              float4 v[NV4];
#pragma unroll
              for (int i = 0; i < NV4; i++)
                v[i] = valid ? vb[t * (KP / 4) + i * 8 + l8] : make_float4(0.f, 0.f, 0.f, 0.f);
              float d = 0.f;
#pragma unroll
              for (int i = 0; i < NV4; i++) d = dot4(v[i], pv[i], d);
              d += __shfl_xor_sync(0xffffffffu, d, 4);
              d += __shfl_xor_sync(0xffffffffu, d, 2);
              d += __shfl_xor_sync(0xffffffffu, d, 1);
              const float c = valid ? cb[t] : 0.f;
              const float w = pass == 0 ? (a.bias + c) - c * d : c * d;
#pragma unroll
              for (int i = 0; i < NV4; i++) axpy4(w, v[i], acc[i]);
            }
          }
          if (!resident) {  // streamed chunk: hand the buffer back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);
          }
        }
        if (active) {
          // reduce-scatter over the 4 groups: lane ends up with elements [4*lane, 4*lane+4)
          const bool hi = (g & 2) != 0, odd = (g & 1) != 0;
          float4 k0 = hi ? acc[2] : acc[0], k1 = hi ? acc[3] : acc[1];
          const float4 s0 = hi ? acc[0] : acc[2], s1 = hi ? acc[1] : acc[3];
          k0 = add4(k0, shfl_xor4(s0, 16));
          k1 = add4(k1, shfl_xor4(s1, 16));
          float4 mine = odd ? k1 : k0;
          mine = add4(mine, shfl_xor4(odd ? k0 : k1, 8));
          // P * pv for this warp's 8 rows of P (P symmetric), columns [4*lane, 4*lane+4)
          const float4 pa = *reinterpret_cast<const float4 *>(pv_src + warp * KR);
          const float4 pb = *reinterpret_cast<const float4 *>(pv_src + warp * KR + 4);
          float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
          axpy4(pa.x, Preg[0], pp); axpy4(pa.y, Preg[1], pp);
          axpy4(pa.z, Preg[2], pp); axpy4(pa.w, Preg[3], pp);
          axpy4(pb.x, Preg[4], pp); axpy4(pb.y, Preg[5], pp);
          axpy4(pb.z, Preg[6], pp); axpy4(pb.w, Preg[7], pp);
          // pass 0 builds r = sum(...) - P x ; later passes build Ap = sum(...) + P p
          if (pass == 0) { mine.x -= pp.x; mine.y -= pp.y; mine.z -= pp.z; mine.w -= pp.w; }
          else mine = add4(mine, pp);
          *reinterpret_cast<float4 *>(partial + ((gpass & 1) * kConsumerWarps + warp) * KP + 4 * lane) = mine;
        }
        consumer_sync();  // partials complete
        if (warp == 0 && active) {
          float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
          const float *pbuf = partial + (gpass & 1) * kConsumerWarps * KP + 4 * lane;
#pragma unroll
          for (int w = 0; w < kConsumerWarps; w++)
            tot = add4(tot, *reinterpret_cast<const float4 *>(pbuf + w * KP));
          int done = 0;
          if (pass == 0) {
            mx = *reinterpret_cast<const float4 *>(x0 + 4 * lane);
            mr = make_float4(fmaf(-reg_u, mx.x, tot.x), fmaf(-reg_u, mx.y, tot.y),
                             fmaf(-reg_u, mx.z, tot.z), fmaf(-reg_u, mx.w, tot.w));
            mp = mr;
            r2 = warp_sum(dot4(mr, mr, 0.f));
            if (r2 <= 1e-20f) done = 1;  // IALSTrainer.hpp:237-240
          } else {
            const float4 Ap = make_float4(fmaf(reg_u, mp.x, tot.x), fmaf(reg_u, mp.y, tot.y),
                                          fmaf(reg_u, mp.z, tot.z), fmaf(reg_u, mp.w, tot.w));
            const float den = warp_sum(dot4(mp, Ap, 0.f));
            if (!(den > 0.f) || !isfinite(den)) {  // :249-254
              failed = true;
              done = 1;
            } else {
              const float alpha = r2 / den;
              axpy4(alpha, mp, mx);
              axpy4(-alpha, Ap, mr);
              const float r2n = warp_sum(dot4(mr, mr, 0.f));
              if (r2n <= 1e-20f) done = 1;  // :258-260
              const float beta = r2n / r2;
              mp = make_float4(fmaf(beta, mp.x, mr.x), fmaf(beta, mp.y, mr.y),
                               fmaf(beta, mp.z, mr.z), fmaf(beta, mp.w, mr.w));
              r2 = r2n;
            }
          }
          *reinterpret_cast<float4 *>(pvec + 4 * lane) = mp;
          if (lane == 0) flags[0] = done;
        } else if (warp == 0 && !active) {
          // converged earlier in a streamed row: keep the flag as it is
        }
        gpass++;
      }
      if (resident) {  // release the row's buffers (all reads are done for this warp)
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < nchunks; j++) mbar_arrive(&empty[(loads + j) & 1]);
      }
      loads += resident ? (unsigned long long)nchunks : (unsigned long long)passes * nchunks;
    }
    if (warp == 0) {
      if (failed) {
        if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
      } else {
        // n == 0: zero row (IALSTrainer.hpp:207-210); otherwise the solved x
        *reinterpret_cast<float4 *>(a.target + u * ld + 4 * lane) = mx;
        for (int pi = 0; pi < a.n_peers; pi++)
          *reinterpret_cast<float4 *>(a.peers[pi] + u * ld + 4 * lane) = mx;
      }
      if (lane == 0) flags[0] = 0;
    }
    q++;
  }
}

// Rows without interactions get x = 0 (IALSTrainer.hpp:207-210); one warp per row.
__global__ void zero_rows_kernel(SolveArgs a) {
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= a.n_rows) return;
  if (a.indptr[warp + 1] != a.indptr[warp]) return;
  const int64_t gu = a.row_base + warp;
  for (int k = lane; k < a.ld; k += kWarp) {
    a.target[gu * a.ld + k] = 0.f;
    for (int pi = 0; pi < a.n_peers; pi++) a.peers[pi][gu * a.ld + k] = 0.f;
  }
}

size_t staged_smem_bytes() {
  constexpr int KP = 128, CHUNK = kRingBytesPerBuffer / (KP * 4);
  size_t floats = (size_t)2 * CHUNK * KP + 2 * CHUNK + 2 * kConsumerWarps * KP + KP + kDescDepth * KP;
  return floats * 4 + kDescDepth * sizeof(RowDesc) + 16 + 8 * 8;
}

}  // namespace

bool cg_staged_supported(const SolveArgs &a) { return a.ld == 128; }

void launch_solve_cg_staged(const SolveArgs &a, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  const size_t smem = staged_smem_bytes();
  CUDA_CHECK(cudaFuncSetAttribute(cg_staged_kernel_k128,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(a.n_sched, 1), sms);
  const int64_t n_rows = a.n_rows;
  if (n_rows > 0) {
    zero_rows_kernel<<<(unsigned)ceil_div(n_rows * kWarp, 256), 256, 0, s>>>(a);
    count_launch();
  }
  cg_staged_kernel_k128<<<grid, kThreads, smem, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
