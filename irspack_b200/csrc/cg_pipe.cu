// K2 conjugate-gradient row solve for the light rows (K padded to 128): warp-per-row batches
// with a software-pipelined neighbour gather.
// Replaces Solver::step_cg, /root/reference/cpp_source/als/IALSTrainer.hpp:170-271; same
// arithmetic, in the same order, as cg_rows.cu (fused b / r-init pass, the reference's exits
// and failure test), so the two kernels agree bit for bit.
//
// ncu on cg_rows_kernel (profiles/r01e_rows_prof_rows.md): 16 warps per SM, each loads a batch
// of 8 neighbour vectors and consumes it in the SAME loop iteration, so every batch exposes one
// L2 round trip (34 % long-scoreboard stalls, issue slots 56 % busy, L1TEX 77 %).  Here
//   * the gather runs through a ring of three 4-neighbour register stages: while one stage is
//     consumed the next two (8 vectors per warp) are already in flight;
//   * neighbour ids and confidences are read 32 at a time, one block ahead, with one coalesced
//     load per lane and handed to the 8-lane groups by shuffle; block 0 of every row is kept in
//     shared memory between the passes, so a pass starts without a dependent DRAM access;
//   * everything that is per row and not per neighbour (row id, CSR offset, degree, reg, |r|^2,
//     P p) lives in a per-warp slab of shared memory: the register budget goes to the ring, and
//     a warp can own R = 4 rows per sweep over P (160 LSU wavefronts per row and pass for P p
//     instead of 288 at R = 2);
//   * the leading (longest) rows of the schedule are handed out one at a time, the rest R at a
//     time, so that the longest-first dynamic queue stays balanced.
// No block-level synchronisation after the prologue; one 512-thread CTA per SM.
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;
constexpr int kPipeWarps = 16;
constexpr int kPipeThreads = kPipeWarps * kWarp;
constexpr unsigned kFull = 0xffffffffu;

struct RowState {  // warp-uniform, one per owned row
  long long gu;    // global target row
  long long s;     // first CSR entry
  int n;           // degree
  float reg_u;     // reg * (alpha0 * n_other + n)^nu, float32 like Solver::compute_reg (:117-120)
  float r2;        // |r|^2 of the last pass
  int pad;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
  return make_float4(__shfl_xor_sync(kFull, v.x, m), __shfl_xor_sync(kFull, v.y, m),
                     __shfl_xor_sync(kFull, v.z, m), __shfl_xor_sync(kFull, v.w, m));
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

template <int R>
constexpr size_t pipe_smem_bytes() {
  return sizeof(float) * ((size_t)KP * KP + (size_t)kPipeWarps * 4 * R * KP)  // P; x, r, p, P p
         + (size_t)kPipeWarps * R * (sizeof(RowState) + 32 * (sizeof(int) + sizeof(float)));
}

// Number of leading schedule slots whose row has more than `d` neighbours (the schedule is in
// descending degree order): a 32-ary search, one probe per lane and round.
__device__ long long count_rows_above(const SolveArgs &a, int d, int lane) {
  if (a.order == nullptr) return 0;  // unsorted schedule: no leading tier
  long long lo = 0, hi = a.n_sched;  // slots < lo are above, slots >= hi are not
  while (lo < hi) {
    const long long step = (hi - lo + 31) / 32;
    const long long p = lo + (long long)lane * step;
    bool above = false;
    if (p < hi) {
      const long long u = a.order[p];
      above = (a.indptr[u + 1] - a.indptr[u]) > (long long)d;
    }
    const int c = __popc(__ballot_sync(kFull, above));
    if (c == 0) {
      hi = lo;
    } else {
      const long long nhi = lo + (long long)c * step;
      lo = lo + (long long)(c - 1) * step + 1;
      hi = nhi < hi ? nhi : hi;
    }
  }
  return lo;
}

// Q[r] = P * V[r] for RR rows in one sweep over P (P symmetric); lane-private result words.
template <int RR>
__device__ __forceinline__ void sweep_P(const float *Ps, const float *Vs, float *Qs, int lane) {
  float4 Pp[RR];
#pragma unroll
  for (int r = 0; r < RR; r++) Pp[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int k = 0; k < KP; k += 4) {
    const float4 p0 = ld4(Ps + (k + 0) * KP + 4 * lane), p1 = ld4(Ps + (k + 1) * KP + 4 * lane);
    const float4 p2 = ld4(Ps + (k + 2) * KP + 4 * lane), p3 = ld4(Ps + (k + 3) * KP + 4 * lane);
#pragma unroll
    for (int r = 0; r < RR; r++) {
      const float4 vk = ld4(Vs + r * KP + k);
      axpy4(vk.x, p0, Pp[r]);
      axpy4(vk.y, p1, Pp[r]);
      axpy4(vk.z, p2, Pp[r]);
      axpy4(vk.w, p3, Pp[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RR; r++) st4(Qs + r * KP + 4 * lane, Pp[r]);
}

template <int R>
__global__ void __launch_bounds__(kPipeThreads, 1) cg_pipe_kernel(SolveArgs a, int single_degree) {
  extern __shared__ __align__(16) float smem[];
  __shared__ long long n_single_s;
  float *Ps = smem;  // [128][128]
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int g = lane >> 3, l8 = lane & 7;  // 8-lane group, lane within the group
  float *slab = Ps + KP * KP + (size_t)warp * 4 * R * KP;
  float *Xs = slab;               // [R][128] x      (lane-private words 4*lane .. 4*lane+3)
  float *Rs = slab + R * KP;      // [R][128] r      (lane-private)
  float *Vs = slab + 2 * R * KP;  // [R][128] vector to multiply: x in pass 0, then p
  float *Qs = slab + 3 * R * KP;  // [R][128] P * V  (lane-private)
  unsigned char *tail = reinterpret_cast<unsigned char *>(Ps + KP * KP + (size_t)kPipeWarps * 4 * R * KP);
  RowState *st = reinterpret_cast<RowState *>(tail) + warp * R;
  int *idx0 = reinterpret_cast<int *>(tail + sizeof(RowState) * kPipeWarps * R) + warp * R * 32;
  float *c0 = reinterpret_cast<float *>(tail + (sizeof(RowState) + 32 * sizeof(int)) * kPipeWarps * R) +
              warp * R * 32;

  if (warp == 0) {
    const long long ns = count_rows_above(a, single_degree, lane);
    if (lane == 0) n_single_s = ns;
  }
  for (int i = threadIdx.x * 4; i < KP * KP; i += kPipeThreads * 4) st4(Ps + i, ld4(a.P + i));
  __syncthreads();
  const long long n_single = n_single_s;

  for (;;) {
    unsigned long long ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.work_counter, (unsigned long long)R);
    ticket = __shfl_sync(kFull, ticket, 0);
    // the n_single longest rows go out one per grab, the others R per grab
    long long first;
    int cnt;
    if ((long long)ticket < n_single * R) {
      first = (long long)ticket / R;
      cnt = 1;
    } else {
      first = n_single + ((long long)ticket - n_single * R);
      const long long left = a.n_sched - first;
      cnt = left < (long long)R ? (int)left : R;
    }
    if (first >= a.n_sched) break;

    unsigned act = 0, fail = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      long long u = 0, s = 0;
      int n = 0;
      if (r < cnt) {
        const long long slot = first + r;
        u = a.order ? (long long)a.order[slot] : slot;
        s = a.indptr[u];
        n = (int)(a.indptr[u + 1] - s);
      }
      const long long gu = a.row_base + u;
      // rows without interactions become zero (IALSTrainer.hpp:207-210)
      const float4 x0 = n > 0 ? ld4(a.target + gu * KP + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      st4(Xs + r * KP + 4 * lane, x0);
      st4(Vs + r * KP + 4 * lane, x0);
      idx0[r * 32 + lane] = lane < n ? a.indices[s + lane] : 0;
      c0[r * 32 + lane] = lane < n ? a.data[s + lane] : 0.f;
      if (lane == 0) {
        RowState rs;
        rs.gu = gu;
        rs.s = s;
        rs.n = n;
        rs.reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)n, a.nu);
        rs.r2 = 0.f;
        rs.pad = 0;
        st[r] = rs;
      }
      if (n > 0) act |= 1u << r;
    }
    __syncwarp();

    for (int pass = 0; pass <= a.max_cg_steps && act != 0; pass++) {
      if (cnt == 1)
        sweep_P<1>(Ps, Vs, Qs, lane);
      else
        sweep_P<R>(Ps, Vs, Qs, lane);

#pragma unroll 1
      for (int r = 0; r < cnt; r++) {
        if (!((act >> r) & 1u)) continue;  // warp-uniform
        // ---- neighbour pass: acc = sum_t coef_t v_t,  coef = bias + c - c (v.x) in pass 0,
        //      c (v.p) afterwards; one neighbour per 8-lane group and stage, three stages ----
        float4 q[4], acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          q[i] = ld4(Vs + r * KP + i * 32 + l8 * 4);
          acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const long long s = st[r].s;
        const int nr = st[r].n;
        const int nh = (nr + 3) >> 2;  // stages of 4 neighbours
        const int32_t *idxp = a.indices + s;
        const float *cp = a.data + s;
        const float *ybase = a.other + l8 * 4;
        // 32 (id, confidence) pairs per lane-block: the block being issued and the next one
        int ci = idx0[r * 32 + lane], ni = 0;
        float cc = c0[r * 32 + lane], nc = 0.f;
        float4 vA[4], vB[4], vC[4];
        float cA = 0.f, cB = 0.f, cC = 0.f;

        // start the loads of stage j (j < nh)
        auto issue = [&](float4(&v)[4], float &cst, int j) {
          const int t0 = j << 2;
          if ((t0 & 31) == 0) {  // first stage of a block: rotate, fetch the following block
            if (j > 0) {
              ci = ni;
              cc = nc;
            }
            const int t = t0 + 32 + lane;
            if (t < nr) {
              ni = idxp[t];
              nc = cp[t];
            }
          }
          const int src = (t0 & 31) + g;
          const int idx = __shfl_sync(kFull, ci, src);
          cst = __shfl_sync(kFull, cc, src);
          const float *y = ybase + (size_t)idx * KP;
#pragma unroll
          for (int i = 0; i < 4; i++) v[i] = ldg4(y + i * 32);
        };
        // fold a landed stage into acc; `valid` masks the neighbours past the end of the row
        auto consume = [&](const float4(&v)[4], float cst, bool valid) {
          float d = dot4(v[0], q[0], 0.f), e = dot4(v[1], q[1], 0.f);
          d = dot4(v[2], q[2], d);
          e = dot4(v[3], q[3], e);
          d += e;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
          float w = pass == 0 ? (a.bias + cst) - cst * d : cst * d;
          w = valid ? w : 0.f;
#pragma unroll
          for (int i = 0; i < 4; i++) axpy4(w, v[i], acc[i]);
        };

        issue(vA, cA, 0);
        if (1 < nh) issue(vB, cB, 1);
        int j = 0;
        // steady state: the three stages issued are inside the row and none of the three
        // consumed is the last one, so nothing is conditional
        for (; j + 5 <= nh; j += 3) {
          issue(vC, cC, j + 2);
          consume(vA, cA, true);
          issue(vA, cA, j + 3);
          consume(vB, cB, true);
          issue(vB, cB, j + 4);
          consume(vC, cC, true);
        }
        for (; j < nh; j += 3) {  // the last (at most four) stages
          if (j + 2 < nh) issue(vC, cC, j + 2);
          consume(vA, cA, (j << 2) + g < nr);
          if (j + 1 < nh) {
            if (j + 3 < nh) issue(vA, cA, j + 3);
            consume(vB, cB, ((j + 1) << 2) + g < nr);
          }
          if (j + 2 < nh) {
            if (j + 4 < nh) issue(vB, cB, j + 4);
            consume(vC, cC, ((j + 2) << 2) + g < nr);
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
        const float reg_u = st[r].reg_u;
        const float r2 = st[r].r2;
        const float4 Pp = ld4(Qs + r * KP + 4 * lane);
        float4 x = ld4(Xs + r * KP + 4 * lane);
        float4 p;
        float r2_new = r2;
        bool done = false;
        if (pass == 0) {
          float4 rv = make_float4(mine.x - Pp.x, mine.y - Pp.y, mine.z - Pp.z, mine.w - Pp.w);
          axpy4(-reg_u, x, rv);
          p = rv;
          r2_new = warp_sum(dot4(rv, rv, 0.f));
          st4(Rs + r * KP + 4 * lane, rv);
          if (r2_new <= 1e-20f) done = true;  // IALSTrainer.hpp:237-240
        } else {
          p = ld4(Vs + r * KP + 4 * lane);
          float4 rv = ld4(Rs + r * KP + 4 * lane);
          float4 Ap = add4(mine, Pp);
          axpy4(reg_u, p, Ap);
          const float den = warp_sum(dot4(p, Ap, 0.f));
          if (!(den > 0.f) || !isfinite(den)) {  // :249-254
            fail |= 1u << r;
            done = true;
          } else {
            const float alpha = r2 / den;
            axpy4(alpha, p, x);
            axpy4(-alpha, Ap, rv);
            st4(Xs + r * KP + 4 * lane, x);
            st4(Rs + r * KP + 4 * lane, rv);
            r2_new = warp_sum(dot4(rv, rv, 0.f));
            if (r2_new <= 1e-20f) {  // :258-260
              done = true;
            } else {
              const float beta = r2_new / r2;
              p = make_float4(fmaf(beta, p.x, rv.x), fmaf(beta, p.y, rv.y), fmaf(beta, p.z, rv.z),
                              fmaf(beta, p.w, rv.w));
            }
          }
        }
        __syncwarp();  // every lane is done reading V[r] (group layout, flat) and st[r]
        if (done) {
          act &= ~(1u << r);
        } else {
          st4(Vs + r * KP + 4 * lane, p);
          if (lane == 0) st[r].r2 = r2_new;
        }
      }
      __syncwarp();  // the new directions (and |r|^2) are visible to the next pass
    }

#pragma unroll 1
    for (int r = 0; r < cnt; r++) {
      if ((fail >> r) & 1u) {  // the reference throws before writing the row back
        if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);
        continue;
      }
      const long long gu = st[r].gu;
      const float4 x = ld4(Xs + r * KP + 4 * lane);
      st4(a.target + gu * KP + 4 * lane, x);
      for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + gu * KP + 4 * lane, x);
    }
    __syncwarp();  // the slab is about to be overwritten by the next grab
  }
}

template <int R>
void launch_pipe(const SolveArgs &a, int single_degree, cudaStream_t s) {
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  constexpr size_t smem = pipe_smem_bytes<R>();
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(cg_pipe_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ctas = ceil_div(a.n_sched, (int64_t)kPipeWarps);
  const unsigned grid = (unsigned)std::min<int64_t>(ctas, sms);
  cg_pipe_kernel<R><<<grid, kPipeThreads, smem, s>>>(a, single_degree);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

// Light rows, ld == 128.  rows_per_warp in {1, 2, 4}; rows with more than single_degree
// neighbours are handed out one per grab.
void launch_solve_cg_pipe(const SolveArgs &a, int rows_per_warp, int single_degree, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  if (a.ld != KP) throw NotImplemented("cg_pipe kernel: ld must be 128");
  switch (rows_per_warp) {
    case 1: launch_pipe<1>(a, single_degree, s); break;
    case 2: launch_pipe<2>(a, single_degree, s); break;
    default: launch_pipe<4>(a, single_degree, s); break;
  }
}

}  // namespace ials
