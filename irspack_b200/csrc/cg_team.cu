// K2 conjugate-gradient row solve, shared-memory-resident team kernel (K padded to 128).
// Replaces Solver::step_cg, /root/reference/cpp_source/als/IALSTrainer.hpp:170-271;
// same arithmetic as cg.cu (fused b / r-init pass, the reference's exits and failure test).
//
// Why: a row makes 1 + max_cg_steps passes over its neighbours' factor vectors.  Re-reading
// them from L2 on every pass (cg_light128_kernel) is bound by the L2 -> SM fabric
// (~42 B/clk/SM); here every vector crosses that fabric ONCE (cp.async, L1 bypassed) into
// shared memory and the later passes run at shared-memory bandwidth (128 B/clk/SM).
//
//   * one 512-thread CTA per SM, cut into teams of T warps (T = 8: two teams, rows of
//     <= 208 neighbours; T = 16: one team, rows of <= 416); a team owns one row at a time,
//     rows are dealt round-robin off the degree-sorted schedule (neighbouring rows have
//     near-equal lengths, so the static deal is balanced and needs no atomics);
//   * warp w of a team owns the neighbours t with (t / 4) % T == w: it stages them itself and
//     is the only reader of their shared-memory copy, so staging needs no team barrier and a
//     team whose rows are being fetched overlaps with the team that computes;
//   * a neighbour vector is owned by an 8-lane group (16 floats per lane, conflict-free
//     LDS.128): dot = 16 FMA + 3 shuffles, update = 16 FMA; the four groups of a warp are
//     reduce-scattered with 12 shuffles;
//   * P (K x K) lives in REGISTERS (128 / T rows per warp, columns 4*lane .. 4*lane+3), so
//     P * p costs no shared-memory or L2 traffic;
//   * two named barriers per pass: warp partials -> 128 threads sum them in a fixed order ->
//     every warp reads the total and runs the scalar CG algebra REDUNDANTLY (bit-identical
//     in all warps, so exits and the failure test need no further communication).
#include "common.cuh"

namespace ials {
namespace {

constexpr int kCtaWarps = 16;
constexpr int kCtaThreads = kCtaWarps * kWarp;  // 512
constexpr int kCapTotal = 416;                  // neighbour vectors resident per CTA
constexpr int KP = 128;

template <int T>
struct TeamCfg {
  static constexpr int NT = kCtaWarps / T;    // teams per CTA
  static constexpr int CAP = kCapTotal / NT;  // neighbours a team keeps resident
  static constexpr int KR = KP / T;           // rows of P per warp
  static constexpr int SLOTS = 4 * T;         // neighbours a team touches per step
  static constexpr int MAX_IT = (CAP + SLOTS - 1) / SLOTS;
};

constexpr size_t kTeamSmemBytes =
    sizeof(float) * ((size_t)kCapTotal * KP + kCapTotal + kCtaWarps * KP + 2 * KP + kCtaWarps * KP);

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), L1 bypassed.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void team_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

template <int T>
__global__ void __launch_bounds__(kCtaThreads, 1) cg_team_kernel(SolveArgs a) {
  using C = TeamCfg<T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *vec = reinterpret_cast<float *>(smem_raw);  // [kCapTotal][KP] staged neighbour vectors
  float *coef = vec + (size_t)kCapTotal * KP;        // [kCapTotal]     their confidences
  float *partial = coef + kCapTotal;                 // [16][KP]        per-warp partial sums
  float *totb = partial + kCtaWarps * KP;            // [2][KP]         per-team totals
  float *pbuf = totb + 2 * KP;                       // [16][KP]        per-warp copy of x / p

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int team = warp / T, w = warp % T;
  const int g = lane >> 3, l8 = lane & 7;  // 8-lane group, lane within the group
  float *tvec = vec + (size_t)team * C::CAP * KP;
  float *tcoef = coef + team * C::CAP;
  float *tpart = partial + team * T * KP;
  float *ttot = totb + team * KP;
  float *mypb = pbuf + warp * KP;
  const int bar_id = 1 + team;

  // P rows [w*KR, w*KR+KR), columns [4*lane, 4*lane+4)
  float4 Preg[C::KR];
#pragma unroll
  for (int kk = 0; kk < C::KR; kk++) Preg[kk] = ld4(a.P + (size_t)(w * C::KR + kk) * KP + 4 * lane);

  const int64_t n_teams = (int64_t)gridDim.x * C::NT;
  for (int64_t slot = (int64_t)blockIdx.x * C::NT + team; slot < a.n_sched; slot += n_teams) {
    const int64_t u = a.order ? (int64_t)a.order[slot] : slot;  // CSR row
    const int64_t gu = a.row_base + u;                          // factor row
    const int64_t s = a.indptr[u];
    const int64_t n64 = a.indptr[u + 1] - s;
    float *xdst = a.target + gu * KP + 4 * lane;
    if (n64 == 0) {  // rows without interactions become zero (IALSTrainer.hpp:207-210)
      if (w == 0) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        st4(xdst, z);
        for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + gu * KP + 4 * lane, z);
      }
      continue;
    }
    if (n64 > C::CAP) {  // the host schedules only rows that fit; never solve a row partially
      if (w == 0 && lane == 0) atomicExch(&a.err_flags[kErrInternal], 1);
      continue;
    }
    const int n = (int)n64;

    // ---- stage this warp's neighbours (and the warm start) ----
    __syncwarp();  // the previous row's reads of mypb / tcoef are done
#pragma unroll
    for (int it = 0; it < C::MAX_IT; it++) {
      const int t = it * C::SLOTS + w * 4 + g;
      if (t < n) {
        const int col = a.indices[s + t];
        if (l8 == 0) tcoef[t] = a.data[s + t];
        const float *src = a.other + (size_t)col * KP + l8 * 4;
        float *dst = tvec + (size_t)t * KP + l8 * 4;
        cp_async16(dst, src);
        cp_async16(dst + 32, src + 32);
        cp_async16(dst + 64, src + 64);
        cp_async16(dst + 96, src + 96);
      }
    }
    cp_async16(mypb + 4 * lane, xdst);
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)n, a.nu);
    cp_async_wait_all();
    __syncwarp();

    float4 x = ld4(mypb + 4 * lane);  // flat layout: elements [4*lane, 4*lane+4)
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f), p = r;
    float4 q[4];                      // group layout: elements i*32 + l8*4 .. +3
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = ld4(mypb + i * 32 + l8 * 4);
    float r2 = 0.f;
    bool failed = false;

    for (int pass = 0; pass <= a.max_cg_steps; pass++) {
      // sum_t coef_t(v_t . q) v_t over this warp's neighbours; pass 0 builds the fused
      // b / r-init coefficients (bias + c - c (v . x)), later passes c (v . p)
      float4 acc[4];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int tb = w * 4; tb < n; tb += C::SLOTS) {  // warp-uniform trip count
        const int t = tb + g;
        const bool valid = t < n;
        const float *vp = tvec + (size_t)(valid ? t : tb) * KP + l8 * 4;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = ld4(vp + i * 32);
        float d0 = dot4(v[0], q[0], 0.f), d1 = dot4(v[1], q[1], 0.f);
        d0 = dot4(v[2], q[2], d0);
        d1 = dot4(v[3], q[3], d1);
        float d = d0 + d1;
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        const float c = tcoef[valid ? t : tb];
        float wgt = pass == 0 ? (a.bias + c) - c * d : c * d;
        wgt = valid ? wgt : 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) axpy4(wgt, v[i], acc[i]);
      }
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
      for (int k4 = 0; k4 < C::KR / 4; k4++) {
        const float4 pk = ld4(mypb + w * C::KR + k4 * 4);
        axpy4(pk.x, Preg[k4 * 4 + 0], pp);
        axpy4(pk.y, Preg[k4 * 4 + 1], pp);
        axpy4(pk.z, Preg[k4 * 4 + 2], pp);
        axpy4(pk.w, Preg[k4 * 4 + 3], pp);
      }
      if (pass == 0) { mine.x -= pp.x; mine.y -= pp.y; mine.z -= pp.z; mine.w -= pp.w; }
      else mine = add4(mine, pp);
      st4(tpart + w * KP + 4 * lane, mine);
      team_bar(bar_id, T * kWarp);
      if (w < 4) {  // 128 threads, one element each, fixed summation order
        const int j = w * kWarp + lane;
        float sum = tpart[j];
#pragma unroll
        for (int ww = 1; ww < T; ww++) sum += tpart[ww * KP + j];
        ttot[j] = sum;
      }
      team_bar(bar_id, T * kWarp);
      const float4 tot = ld4(ttot + 4 * lane);

      if (pass == 0) {
        r = make_float4(fmaf(-reg_u, x.x, tot.x), fmaf(-reg_u, x.y, tot.y),
                        fmaf(-reg_u, x.z, tot.z), fmaf(-reg_u, x.w, tot.w));
        p = r;
        r2 = warp_sum(dot4(r, r, 0.f));
        if (r2 <= 1e-20f) break;  // IALSTrainer.hpp:237-240
      } else {
        const float4 Ap = make_float4(fmaf(reg_u, p.x, tot.x), fmaf(reg_u, p.y, tot.y),
                                      fmaf(reg_u, p.z, tot.z), fmaf(reg_u, p.w, tot.w));
        const float den = warp_sum(dot4(p, Ap, 0.f));
        if (!(den > 0.f) || !isfinite(den)) {  // :249-254
          failed = true;
          break;
        }
        const float alpha = r2 / den;
        axpy4(alpha, p, x);
        axpy4(-alpha, Ap, r);
        const float r2n = warp_sum(dot4(r, r, 0.f));
        if (r2n <= 1e-20f) break;  // :258-260
        const float beta = r2n / r2;
        p = make_float4(fmaf(beta, p.x, r.x), fmaf(beta, p.y, r.y), fmaf(beta, p.z, r.z),
                        fmaf(beta, p.w, r.w));
        r2 = r2n;
      }
      if (pass == a.max_cg_steps) break;
      // republish p in this warp's private buffer, reload it in the group layout
      __syncwarp();
      st4(mypb + 4 * lane, p);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; i++) q[i] = ld4(mypb + i * 32 + l8 * 4);
    }
    if (w == 0) {
      if (failed) {
        if (lane == 0) atomicExch(&a.err_flags[kErrCgSingular], 1);  // the reference throws before the write-back
      } else {
        st4(xdst, x);
        for (int pi = 0; pi < a.n_peers; pi++) st4(a.peers[pi] + gu * KP + 4 * lane, x);
      }
    }
  }
}

template <int T>
void launch_team(const SolveArgs &a, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(cg_team_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kTeamSmemBytes));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ctas = ceil_div(a.n_sched, (int64_t)TeamCfg<T>::NT);
  const unsigned grid = (unsigned)std::min<int64_t>(ctas, sms);
  cg_team_kernel<T><<<grid, kCtaThreads, kTeamSmemBytes, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace

int cg_team_capacity(int team_warps) { return team_warps == 8 ? TeamCfg<8>::CAP : TeamCfg<16>::CAP; }

// Rows of <= 208 neighbours: two 8-warp teams per SM.
void launch_solve_cg_team8(const SolveArgs &a, cudaStream_t s) {
  if (a.ld != KP) throw NotImplemented("cg_team kernel: ld must be 128");
  launch_team<8>(a, s);
}
// Rows of <= 416 neighbours: one 16-warp team per SM.
void launch_solve_cg_team16(const SolveArgs &a, cudaStream_t s) {
  if (a.ld != KP) throw NotImplemented("cg_team kernel: ld must be 128");
  launch_team<16>(a, s);
}

}  // namespace ials
