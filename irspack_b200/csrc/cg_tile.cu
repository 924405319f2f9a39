// K2 conjugate-gradient row solve, ONE-TOUCH kernel for the light rows (K padded to 128).
// Replaces Solver::step_cg, /root/reference/cpp_source/als/IALSTrainer.hpp:170-271; same
// arithmetic as cg_rows.cu (fused b / r-init pass, the reference's exits and failure test).
//
// Why (profiles/r01e_rows_prof_rows.md): cg_rows_kernel walks a row's
// neighbours once per pass (1 + max_cg_steps = 4 passes) straight from L2: 33.5 GB of L2 -> SM
// reads per user half-epoch against 10.5 GB algorithmic, i.e. the ~6300 B/clk LTS cap of the
// chip -- and 4x the algorithmic HBM bytes as soon as the gathered matrix does not fit the
// 126 MB L2 (configs[3]).  Here a neighbour vector crosses L2 -> SM ONCE per row:
//   * one CTA (a team of T warps) owns a row.  T = 8: 256 threads, two CTAs per SM, so that one
//     team's loads overlap the other's arithmetic, CAP = 204 resident neighbours; T = 16: 512
//     threads, one CTA per SM, CAP = 412 (A/B);
//   * warp w owns the neighbours t with (t / 4) % T == w.  Pass 0 gathers them with LDG.128
//     (8-lane groups, 16 floats per lane, as cg_rows), uses them from registers and parks the
//     first CAP of the row in shared memory; passes 1 .. max_cg_steps read them back with
//     conflict-free LDS.128 (4 wavefronts per vector = the shared-memory floor).  Only the
//     tail of a row longer than CAP streams from L2 again;
//   * warp w is the only writer and reader of its slots: no barrier between the passes' tile
//     accesses, and the next row's pass 0 needs none either;
//   * P (K x K) lives in REGISTERS (128 / T rows per warp, columns 4*lane .. 4*lane+3): P * p
//     costs no shared-memory or L2 traffic; its partial sum joins the warp's partial of the
//     neighbour sum.  With 64 registers of P per thread (T = 8, 128 registers at two CTAs per
//     SM) the cold CG state must not be in registers during the neighbour loop, or ptxas spills
//     P to local memory (= L2 traffic): x lives in a CTA-wide shared buffer kept by warp 0, the
//     search direction in the warp's private buffer, row ids are 32-bit;
//   * two barriers per pass: warp partials -> 128 threads add them in a fixed order -> every
//     warp reads the total and runs the scalar CG algebra redundantly (bit-identical in all
//     warps: exits and the failure test need no further communication);
//   * rows are dealt statically off the degree-sorted schedule, boustrophedon over the CTAs, so
//     the next row's (row id, extent, warm start, first neighbour ids) are prefetched while the
//     current row is solved: no dependent-load chain at a row boundary.
// Algorithmic bytes per row of n neighbours: n (4K + 8) + 8K + 8.
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;

template <int T>
struct TileCfg {
  static constexpr int kThreads = T * kWarp;
  static constexpr int kCtasPerSm = 16 / T;
  static constexpr int SLOTS = 4 * T;  // neighbours a team touches per step
  static constexpr int KR = KP / T;    // rows of P per warp
  // shared memory: vec [CAP][KP] | coef [CAP] | part [T][KP] | tot [KP] | pbuf [T][KP] | xbuf [2][KP]
  static constexpr size_t kFixedFloats = (size_t)T * KP + KP + (size_t)T * KP + 2 * KP;
  static constexpr size_t kBudget = (size_t)(228 * 1024) / kCtasPerSm - 1024;
  static constexpr int CAP = (int)((kBudget - kFixedFloats * 4) / ((KP + 1) * 4)) / 4 * 4;
  static constexpr size_t kSmemBytes = ((size_t)CAP * (KP + 1) + kFixedFloats) * sizeof(float);
  static_assert(kCtasPerSm * (kSmemBytes + 1024) <= 228 * 1024, "CTAs per SM must fit shared memory");
  static_assert(kSmemBytes <= 232448, "opt-in shared memory limit per CTA");
};

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
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), no register staging
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Row `k`-th of this CTA: boustrophedon over the grid (rows are sorted by descending degree, so
// consecutive rounds hand a CTA alternately the longest and the shortest row of the round).
__device__ __forceinline__ int64_t slot_of(int k) {
  const int64_t G = gridDim.x, b = blockIdx.x;
  return (int64_t)k * G + ((k & 1) ? G - 1 - b : b);
}

struct RowMeta {  // 32-bit: rows and nnz fit int32 (api.cu new_trainer / upload_csr)
  int gu, s, n;   // factor row, first CSR entry, neighbours
};

template <int T>
__global__ void __launch_bounds__(TileCfg<T>::kThreads, TileCfg<T>::kCtasPerSm) cg_tile_kernel(SolveArgs a) {
  using C = TileCfg<T>;
  constexpr int CAP = C::CAP, SLOTS = C::SLOTS, KR = C::KR;
  extern __shared__ __align__(16) float smem[];
  float *vec = smem;                        // [CAP][KP] resident neighbour vectors
  float *coef = vec + (size_t)CAP * KP;     // [CAP]     their confidences
  float *part = coef + CAP;                 // [T][KP]   per-warp partial sums
  float *tot = part + T * KP;               // [KP]      their total
  float *pbuf = tot + KP;                   // [T][KP]   per-warp copy of the search direction
  float *xbuf = pbuf + T * KP;              // [2][KP]   x of the current row (warp 0 keeps it) / next row

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int g = lane >> 3, l8 = lane & 7;  // 8-lane group, lane within the group
  float *mypb = pbuf + w * KP;
  const float *ybase = a.other + l8 * 4;

  // P rows [w*KR, w*KR + KR), columns [4*lane, 4*lane + 4)
  float4 Preg[KR];
#pragma unroll
  for (int kk = 0; kk < KR; kk++) Preg[kk] = ld4(a.P + (size_t)(w * KR + kk) * KP + 4 * lane);

  auto load_meta = [&](int k, RowMeta &m) -> bool {
    const int64_t slot = slot_of(k);
    if (slot >= a.n_sched) return false;
    const int64_t u = a.order ? (int64_t)a.order[slot] : slot;
    const int64_t s64 = a.indptr[u];
    m.gu = (int)(a.row_base + u);
    m.s = (int)s64;
    m.n = (int)(a.indptr[u + 1] - s64);
    return true;
  };

  RowMeta cur, nxt;
  bool have = load_meta(0, cur);
  // first neighbour (id, confidence) of this lane group in the current row, prefetched
  int idx0 = 0;
  float c0 = 0.f;
  if (have) {
    if (w == 0) st4(xbuf + 4 * lane, ld4(a.target + (size_t)cur.gu * KP + 4 * lane));
    const int t0 = w * 4 + g;
    if (t0 < cur.n) {
      idx0 = a.indices[(size_t)cur.s + t0];
      c0 = a.data[(size_t)cur.s + t0];
    }
  }
  __syncthreads();

  for (int k = 0; have; k++) {
    const bool have_next = load_meta(k + 1, nxt);
    const int n = cur.n;
    float *xs = xbuf + (k & 1) * KP;        // this row's x (flat)
    float *xn = xbuf + ((k + 1) & 1) * KP;  // the next row's warm start lands here
    int nidx0 = 0;
    float nc0 = 0.f;
    auto prefetch_next = [&]() {  // the next row's warm start and first neighbour of this lane group
      if (have_next) {
        if (w == 0) cp_async16(xn + 4 * lane, a.target + (size_t)nxt.gu * KP + 4 * lane);
        const int t0 = w * 4 + g;
        if (t0 < nxt.n) {
          nidx0 = a.indices[(size_t)nxt.s + t0];
          nc0 = a.data[(size_t)nxt.s + t0];
        }
      }
    };

    if (n > 0) {
      const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)n, a.nu);  // :117-120
      const int32_t *idxp = a.indices + cur.s;
      const float *datp = a.data + cur.s;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      float r2 = 0.f;
      bool failed = false;

      for (int pass = 0; pass <= a.max_cg_steps; pass++) {
        // the vector this pass multiplies: x (CTA-wide buffer) in pass 0, then the search
        // direction (this warp's buffer); group layout: elements i*32 + l8*4 .. +3
        const float *pv = pass == 0 ? xs : mypb;
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = ld4(pv + i * 32 + l8 * 4);
        // sum_t coef_t v_t over this warp's neighbours: coef = bias + c - c (v . x) in pass 0
        // (fused b / r-init), c (v . p) afterwards
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int idx = idx0;  // (id, confidence) of the entry this group handles next, one step ahead
        float cc = c0;
        for (int tb = w * 4; tb < n; tb += SLOTS) {  // warp-uniform trip count
          const int t = tb + g;
          const bool valid = t < n;
          const bool resident = tb < CAP;  // CAP % 4 == 0: uniform over the four groups
          float4 v[4];
          float c;
          if (pass > 0 && resident) {
            const float *vp = vec + (size_t)(valid ? t : tb) * KP + l8 * 4;
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = ld4(vp + i * 32);
            c = coef[valid ? t : tb];
          } else {
            const float *vp = ybase + (size_t)idx * KP;
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = ldg4(vp + i * 32);
            c = cc;
          }
          // (id, confidence) of this group's entry of the next step, if that step gathers from
          // L2 (always in pass 0, the tail beyond CAP afterwards)
          const int tn = t + SLOTS;
          if (tn < n && (pass == 0 || tb + SLOTS >= CAP)) {
            idx = idxp[tn];
            cc = datp[tn];
          }
          if (pass == 0 && resident && valid) {
            float *dp = vec + (size_t)t * KP + l8 * 4;
#pragma unroll
            for (int i = 0; i < 4; i++) st4(dp + i * 32, v[i]);
            if (l8 == 0) coef[t] = c;
          }
          float d0 = dot4(v[0], q[0], 0.f), d1 = dot4(v[1], q[1], 0.f);
          d0 = dot4(v[2], q[2], d0);
          d1 = dot4(v[3], q[3], d1);
          float d = d0 + d1;
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          float wgt = pass == 0 ? (a.bias + c) - c * d : c * d;
          wgt = valid ? wgt : 0.f;
#pragma unroll
          for (int i = 0; i < 4; i++) axpy4(wgt, v[i], acc[i]);
        }
        if (pass == 0) prefetch_next();  // lands while the CG passes run
        // reduce-scatter over the 4 groups: lane ends up with elements [4*lane, 4*lane+4)
        const bool hi = (g & 2) != 0, odd = (g & 1) != 0;
        float4 k0 = hi ? acc[2] : acc[0], k1 = hi ? acc[3] : acc[1];
        const float4 s0 = hi ? acc[0] : acc[2], s1 = hi ? acc[1] : acc[3];
        k0 = add4(k0, shfl_xor4(s0, 16));
        k1 = add4(k1, shfl_xor4(s1, 16));
        float4 mine = odd ? k1 : k0;
        mine = add4(mine, shfl_xor4(odd ? k0 : k1, 8));
        // P * (x or p) for this warp's rows of P (P symmetric)
        float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k4 = 0; k4 < KR / 4; k4++) {
          const float4 pk = ld4(pv + w * KR + k4 * 4);
          axpy4(pk.x, Preg[k4 * 4 + 0], pp);
          axpy4(pk.y, Preg[k4 * 4 + 1], pp);
          axpy4(pk.z, Preg[k4 * 4 + 2], pp);
          axpy4(pk.w, Preg[k4 * 4 + 3], pp);
        }
        if (pass == 0) { mine.x -= pp.x; mine.y -= pp.y; mine.z -= pp.z; mine.w -= pp.w; }
        else mine = add4(mine, pp);
        st4(part + w * KP + 4 * lane, mine);
        __syncthreads();
        if (w < 4) {  // 128 threads, one element each, fixed summation order
          const int j = w * kWarp + lane;
          float sum = part[j];
#pragma unroll
          for (int ww = 1; ww < T; ww++) sum += part[ww * KP + j];
          tot[j] = sum;
        }
        __syncthreads();
        const float4 tv = ld4(tot + 4 * lane);

        float4 p;
        if (pass == 0) {
          const float4 x = ld4(xs + 4 * lane);
          r = make_float4(fmaf(-reg_u, x.x, tv.x), fmaf(-reg_u, x.y, tv.y),
                          fmaf(-reg_u, x.z, tv.z), fmaf(-reg_u, x.w, tv.w));
          p = r;
          r2 = warp_sum(dot4(r, r, 0.f));
          if (r2 <= 1e-20f) break;  // IALSTrainer.hpp:237-240
        } else {
          p = ld4(mypb + 4 * lane);
          const float4 Ap = make_float4(fmaf(reg_u, p.x, tv.x), fmaf(reg_u, p.y, tv.y),
                                        fmaf(reg_u, p.z, tv.z), fmaf(reg_u, p.w, tv.w));
          const float den = warp_sum(dot4(p, Ap, 0.f));
          if (!(den > 0.f) || !isfinite(den)) {  // :249-254
            failed = true;
            break;
          }
          const float alpha = r2 / den;
          if (w == 0) {  // x is only needed again for the write-back: warp 0 keeps it
            float4 x = ld4(xs + 4 * lane);
            axpy4(alpha, p, x);
            st4(xs + 4 * lane, x);
          }
          axpy4(-alpha, Ap, r);
          const float r2n = warp_sum(dot4(r, r, 0.f));
          if (r2n <= 1e-20f) break;  // :258-260
          const float beta = r2n / r2;
          p = make_float4(fmaf(beta, p.x, r.x), fmaf(beta, p.y, r.y), fmaf(beta, p.z, r.z),
                          fmaf(beta, p.w, r.w));
          r2 = r2n;
        }
        if (pass == a.max_cg_steps) break;
        // publish the new direction in this warp's private buffer (flat layout)
        __syncwarp();  // every lane has read the old one
        st4(mypb + 4 * lane, p);
        __syncwarp();
      }
      if (w == 0) {
        if (failed) {
          if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);  // the reference throws before the write-back
        } else {
          const float4 x = ld4(xs + 4 * lane);
          st4(a.target + (size_t)cur.gu * KP + 4 * lane, x);
          for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + (size_t)cur.gu * KP + 4 * lane, x);
        }
      }
    } else {  // rows without interactions become zero (IALSTrainer.hpp:207-210)
      if (w == 0) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        st4(a.target + (size_t)cur.gu * KP + 4 * lane, z);
        for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + (size_t)cur.gu * KP + 4 * lane, z);
      }
      prefetch_next();
    }
    // the next row's warm start has landed in xn, and every warp is done with xs / its buffers
    if (w == 0) cp_async_wait_all();
    __syncthreads();
    cur = nxt;
    have = have_next;
    idx0 = nidx0;
    c0 = nc0;
  }
}

template <int T>
void launch_tile(const SolveArgs &a, cudaStream_t s) {
  using C = TileCfg<T>;
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(cg_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)C::kSmemBytes));
    CUDA_CHECK(cudaFuncSetAttribute(cg_tile_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)std::min<int64_t>(a.n_sched, (int64_t)sms * C::kCtasPerSm);
  cg_tile_kernel<T><<<grid, C::kThreads, C::kSmemBytes, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

int cg_tile_capacity(int team_warps) { return team_warps == 16 ? TileCfg<16>::CAP : TileCfg<8>::CAP; }

// Light rows, ld == 128: every scheduled row, whatever its length (rows longer than the
// resident capacity stream their tail from L2 on the later passes).  team_warps in {8, 16}.
void launch_solve_cg_tile(const SolveArgs &a, int team_warps, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  if (a.ld != KP) throw NotImplemented("cg_tile kernel: ld must be 128");
  if (team_warps == 16) launch_tile<16>(a, s);
  else launch_tile<8>(a, s);
}

}  // namespace ials
