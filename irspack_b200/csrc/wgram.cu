// Weighted, gathered Gram on the 5th-generation tensor cores (tcgen05 + TMEM):
//
//     G_job = sum_{t in job} w_t * y_{i_t} y_{i_t}^T          (K x K, K padded to 128)
//
// One kernel, three callers:
//   * K1  Solver::prepare_p                 P = alpha0 * Y^T Y   (all rows, w = 1)
//         /root/reference/cpp_source/als/IALSTrainer.hpp:78-115
//   * K2  heavy rows of Solver::step_cg     A_u = P + reg_u I + sum c y y^T formed explicitly
//         (:216-247 evaluate the same operator neighbour by neighbour)
//   * K3  Solver::step_cholesky's rank update (BatchedRankUpdater, :37-58, 301-308)
//
// float32 parity on TF32 tensor cores: u = sqrt(w) * y is split into hi = tf32(u) and
// lo = u - hi (exact), and  u u^T = hi hi^T + hi lo^T + lo hi^T + O(2^-22).  The
// kernel accumulates  HH = sum hi hi^T  and  HL = sum hi lo^T  in two TMEM
// accumulators (fp32) and emits  W = HH / 2 + HL;  consumers use  G = W + W^T,
// which is exactly symmetric.  One N = 256 MMA per k-step ([hi | lo] as the B operand).
//
// Structure (one persistent CTA per SM, 672 threads):
//   warps 0-15 producers (four groups of four warps, round-robin over the stages): gather the neighbour rows with 128-bit loads, scale, split,
//              and store both operand tiles into shared memory in the UMMA canonical
//              MN-major SWIZZLE_128B_BASE32B layout (4 panels of 32 features, 128-byte rows,
//              32-byte chunks XOR-swizzled by row mod 4); also b = sum (bias + w) y.
//   warp  20   MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M = N = 128,
//              K = 8 per instruction) and tcgen05.commit to the stage / accumulator barriers.
//   warps 16-19 epilogue: tcgen05.ld the two accumulators (double-buffered: 2 x 256 TMEM
//              columns), combine, and write W to global memory.
// Stages: 4 x (32 neighbours x (hi + lo) x 512 B) = 128 KB of shared memory.
#include "common.cuh"

namespace ials {
namespace {

constexpr int KP = 128;       // padded feature dimension = UMMA M = UMMA N
constexpr int KT = 32;        // neighbours per pipeline stage
constexpr int STAGES = 4;
constexpr int kProducerWarps = 4;  // per producer group (one stage = 32 neighbours = 4 x 8)
constexpr int kGroups = 4;         // producer groups; group g fills stages g, g + 4, ...
constexpr int kAllProducerWarps = kGroups * kProducerWarps;
static_assert(kAllProducerWarps == kWGramBParts, "bpart layout");
static_assert(kAllProducerWarps % 4 == 0, "epilogue warps must start on a TMEM lane quadrant");
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kAllProducerWarps + kEpilogueWarps + 1) * kWarp;  // 416
constexpr int kPanelBytes = KT * 128;          // one 32-feature panel of a tile
constexpr int kTileBytes = 4 * kPanelBytes;    // 16 KB: hi or lo operand of one stage
constexpr int kStageBytes = 2 * kTileBytes;    // 32 KB
constexpr int kTmemCols = 512;                 // 2 buffers x (HH 128 + HL 128)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  // a protocol bug becomes a trapped launch (an error), never a hung GPU
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far has completed
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor of one [128 features x 8 neighbours] MN-major operand
// slice.  MN-major tf32 operands exist in ONE canonical layout only, SWIZZLE_128B_BASE32B
// ("128-byte swizzle with 32-byte atomicity"; cute/atom/mma_traits_sm100.hpp,
// make_umma_desc<Major::MN>):  ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in tf32 elements under
// Swizzle<2,5,2>, i.e. 32 consecutive features are 128 contiguous bytes, consecutive
// neighbours are 128 bytes apart, 4 neighbours form a 512-byte swizzle atom, the next
// 32-feature panel is LBO bytes away and the next group of 4 neighbours SBO bytes away.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);               // start address
  d |= (uint64_t)((kPanelBytes >> 4) & 0x3FFF) << 16;       // leading byte offset (panel stride)
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;               // stride byte offset (4-neighbour group)
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                                   // SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor: D fp32, A = B = tf32, both MN-major, M = 128, N = 256: the B
// operand is the stage's hi tile followed by its lo tile (8 panels, same panel stride), so
// ONE instruction per k-step yields HH in accumulator columns 0-127 and HL in 128-255 and
// the A operand is fetched from shared memory once.
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)((2 * KP) >> 3) << 17) | ((uint32_t)(KP >> 4) << 24);

__device__ __forceinline__ void tmem_st_probe(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t accumulate, uint32_t idesc = kInstrDesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Position in this CTA's sequence of pipeline stages (jobs blockIdx.x, + grid, ...; KT
// entries per stage).  `it` numbers the stages: the ring slot is it % STAGES.
struct StageCursor {
  int j, base, je;  // job, first entry of the stage, end of the job (entry offsets fit int32)
  unsigned it;
  __device__ __forceinline__ bool valid(const WGramArgs &a) const { return j < (int)a.n_jobs; }
  __device__ __forceinline__ void seek(const WGramArgs &a, int grid) {  // skip empty jobs
    base = je = 0;
    while (j < (int)a.n_jobs) {
      base = (int)a.job_begin[j];
      je = (int)a.job_end[j];
      if (je > base) return;
      j += grid;
    }
  }
  __device__ __forceinline__ void advance(const WGramArgs &a, int grid) {
    base += KT;
    it++;
    if (base >= je) {
      j += grid;
      seek(a, grid);
    }
  }
};

__global__ void __launch_bounds__(kThreads, 1) wgram_kernel(WGramArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to the swizzle atom
  unsigned char *tiles = reinterpret_cast<unsigned char *>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + STAGES * kStageBytes);
  uint64_t *full = bars;                 // [STAGES]  producers -> MMA
  uint64_t *empty = bars + STAGES;       // [STAGES]  MMA (commit) -> producers
  uint64_t *accfull = bars + 2 * STAGES; // [2]       MMA (commit) -> epilogue
  uint64_t *accempty = accfull + 2;      // [2]       epilogue -> MMA
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], kProducerWarps);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&accfull[b], 1);
      mbar_init(&accempty[b], kEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAllProducerWarps + kEpilogueWarps) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kAllProducerWarps) {
    // ================================ PRODUCERS ================================
    // kGroups groups of kProducerWarps warps; group g fills the stages whose number is
    // congruent to g modulo kGroups, so that kGroups (index -> gather -> convert) latency
    // chains overlap; the neighbour ids are prefetched two own stages ahead.  Every warp
    // walks the whole stage sequence (cheap cursor arithmetic) so that it can write its
    // part of b for every job of the CTA.
    const int group = warp / kProducerWarps, pw = warp % kProducerWarps;
    const int panel = lane >> 3, chunk = lane & 7;  // this lane's 16 bytes of every row
    const int grid = (int)gridDim.x;
    int flushed = (int)blockIdx.x - grid;  // last job whose b slot this warp wrote
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush_until = [&](int j_stop) {  // write the b slots of this CTA's jobs < j_stop
      for (int jj = flushed + grid; jj < j_stop && jj < (int)a.n_jobs; jj += grid) {
        if (a.bpart)
          *reinterpret_cast<float4 *>(a.bpart + ((size_t)jj * kAllProducerWarps + warp) * KP + 4 * lane) = bacc;
        bacc = make_float4(0.f, 0.f, 0.f, 0.f);
        flushed = jj;
      }
    };
    auto next_own = [&](StageCursor c) {
      for (int g = 0; g < kGroups && c.valid(a); g++) c.advance(a, grid);
      return c;
    };
    auto load_ids = [&](const StageCursor &c, int &row, float &w) {  // one neighbour per lane
      row = 0;
      w = 0.f;
      if (c.valid(a) && c.base + lane < c.je) {
        row = a.indices ? a.indices[c.base + lane] : c.base + lane;
        w = a.weights ? a.weights[c.base + lane] : 1.f;
      }
    };
    StageCursor cur;
    cur.j = (int)blockIdx.x;
    cur.it = 0;
    cur.seek(a, grid);
    for (int g = 0; g < group && cur.valid(a); g++) cur.advance(a, grid);
    StageCursor n1 = next_own(cur), n2 = next_own(n1);
    int row0, row1, row2;
    float w0, w1, w2;
    load_ids(cur, row0, w0);
    load_ids(n1, row1, w1);
    load_ids(n2, row2, w2);
    constexpr int NPW = KT / kProducerWarps;  // neighbours per producer warp and stage
    while (cur.valid(a)) {
      flush_until(cur.j);  // everything before the current job is complete for this warp
      const int s = (int)(cur.it % STAGES);
      const uint32_t ph = (uint32_t)((cur.it / STAGES) & 1);
      const int m = min(KT, cur.je - cur.base);
      float4 v[NPW];
#pragma unroll
      for (int q = 0; q < NPW; q++) {
        const int t = q * kProducerWarps + pw;
        const int row = __shfl_sync(0xffffffffu, row0, t);
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < m && !(a.debug_flags & 16))
          v[q] = *reinterpret_cast<const float4 *>(a.Y + (size_t)row * a.ld + 4 * lane);
      }
      const StageCursor n3 = next_own(n2);
      int row3;
      float w3;
      load_ids(n3, row3, w3);

      mbar_wait(&empty[s], ph ^ 1);
      unsigned char *hi = tiles + s * kStageBytes + panel * kPanelBytes;
      unsigned char *lo = hi + kTileBytes;
      if (!(a.debug_flags & 4))  // (bring-up: 4 = skip the conversion / stores)
#pragma unroll
      for (int q = 0; q < NPW; q++) {
        const int t = q * kProducerWarps + pw;
        const float w = __shfl_sync(0xffffffffu, w0, t);
        const float sc = sqrtf(fmaxf(w, 0.f));
        float4 u = make_float4(sc * v[q].x, sc * v[q].y, sc * v[q].z, sc * v[q].w);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(u.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(u.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(u.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(u.w) & 0xffffe000u);
        l = make_float4(u.x - h.x, u.y - h.y, u.z - h.z, u.w - h.w);
        // Swizzle<2,5,2>: the 32-byte chunk index is XORed with the row index mod 4
        const int off = t * 128 + ((((chunk >> 1) ^ (t & 3)) << 5) | ((chunk & 1) << 4));
        *reinterpret_cast<float4 *>(hi + off) = h;
        *reinterpret_cast<float4 *>(lo + off) = l;
        const float cb = t < m ? a.bias + w : 0.f;
        bacc.x = fmaf(cb, v[q].x, bacc.x);
        bacc.y = fmaf(cb, v[q].y, bacc.y);
        bacc.z = fmaf(cb, v[q].z, bacc.z);
        bacc.w = fmaf(cb, v[q].w, bacc.w);
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      cur = n1; n1 = n2; n2 = n3;
      row0 = row1; row1 = row2; row2 = row3;
      w0 = w1; w1 = w2; w2 = w3;
    }
    flush_until((int)a.n_jobs);
  } else if (warp == kAllProducerWarps + kEpilogueWarps) {
    // ================================ MMA ISSUER ================================
    unsigned long long it = 0, jc = 0;
    for (long long j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
      const long long jb = a.job_begin[j], je = a.job_end[j];
      if (je <= jb) continue;  // nothing to accumulate: the epilogue writes zeros
      const int buf = (int)(jc & 1);
      mbar_wait(&accempty[buf], (uint32_t)(((jc >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_hh = tmem_base + (uint32_t)(buf * 256);
      uint32_t acc = 0;
      for (long long base = jb; base < je; base += KT, it++) {
        const int s = (int)(it % STAGES);
        mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
        tc_fence_after();
        if (lane == 0) {
          const uint32_t hi = smem_u32(tiles + s * kStageBytes);
          if (!(a.debug_flags & 8))  // (bring-up: 8 = issue no MMAs, only the commits)
#pragma unroll
          for (int k = 0; k < KT / 8; k++) {
            const uint64_t dh = make_desc(hi + k * 1024);
            mma_tf32(d_hh, dh, dh, acc);
            acc = 1;
          }
          tc_commit(&empty[s]);                       // the slot is free once these MMAs retire
          if (base + KT >= je) tc_commit(&accfull[buf]);  // ... and so is the job's accumulator
        }
        __syncwarp();
      }
      jc++;
    }
  } else {
    // ================================ EPILOGUE ================================
    const int ew = warp - kAllProducerWarps;  // == warp % 4: the TMEM lane quadrant of this warp
    const int row = ew * 32 + lane;        // accumulator row = feature index a
    unsigned long long jc = 0;
    for (long long j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
      float *out = a.W + (size_t)j * KP * KP + (size_t)row * KP;
      if (a.job_end[j] <= a.job_begin[j]) {
#pragma unroll 4
        for (int c = 0; c < KP; c += 4) *reinterpret_cast<float4 *>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const int buf = (int)(jc & 1);
      mbar_wait(&accfull[buf], (uint32_t)((jc >> 1) & 1));
      tc_fence_after();
      const uint32_t t_hh = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 256);
      if (a.debug_tmem != nullptr && blockIdx.x == 0 && jc == 0) {
        if (a.debug_flags & 2)  // probe: write a pattern into column 300 of every lane, read it back below
          tmem_st_probe(tmem_base + ((uint32_t)(ew * 32) << 16) + 300, 0x42280000u + (uint32_t)row);
        if (row == 0) a.debug_tmem[128 * 512] = __uint_as_float(tmem_base);
        for (int c = 0; c < kTmemCols; c += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + c, raw);
          tmem_ld_wait();
          for (int q = 0; q < 32; q++) a.debug_tmem[(size_t)row * kTmemCols + c + q] = __uint_as_float(raw[q]);
        }
      }
#pragma unroll 1
      for (int c = 0; c < KP; c += 16) {
        uint32_t hh[16], hl[16];
        tmem_ld16(t_hh + c, hh);
        tmem_ld16(t_hh + 128 + c, hl);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          float4 o;
          o.x = fmaf(0.5f, __uint_as_float(hh[q + 0]), __uint_as_float(hl[q + 0]));
          o.y = fmaf(0.5f, __uint_as_float(hh[q + 1]), __uint_as_float(hl[q + 1]));
          o.z = fmaf(0.5f, __uint_as_float(hh[q + 2]), __uint_as_float(hl[q + 2]));
          o.w = fmaf(0.5f, __uint_as_float(hh[q + 3]), __uint_as_float(hl[q + 3]));
          *reinterpret_cast<float4 *>(out + c + q) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[buf]);
      jc++;
    }
  }

  // teardown: everybody done with TMEM before the owner frees it
  tc_fence_before();
  __syncthreads();
  if (warp == kAllProducerWarps + kEpilogueWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// G = scale * sum_j (W_j + W_j^T) over a contiguous run of partials (K1 finalize).
__global__ void wgram_reduce_sym_kernel(const float *__restrict__ W, int n_parts, float scale,
                                        float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KP * KP) return;
  const int r = i / KP, c = i % KP;
  float acc = 0.f;
  for (int p = 0; p < n_parts; p++) {
    const float *w = W + (size_t)p * KP * KP;
    acc += w[r * KP + c] + w[c * KP + r];
  }
  out[i] = scale * acc;
}

// Contiguous row blocks (multiples of KT rows) for the plain Gram.
__global__ void block_jobs_kernel(int64_t begin, int64_t end, int n_jobs, int64_t *jb, int64_t *je) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const int64_t n = end - begin;
  int64_t per = (n + n_jobs - 1) / n_jobs;
  per = (per + KT - 1) / KT * KT;
  const int64_t b = min(begin + (int64_t)j * per, end);
  jb[j] = b;
  je[j] = min(b + per, end);
}

}  // namespace

void GramWorkspace::alloc(int jobs) {
  max_jobs = jobs;
  CUDA_CHECK(cudaMalloc(&job_begin, sizeof(int64_t) * jobs));
  CUDA_CHECK(cudaMalloc(&job_end, sizeof(int64_t) * jobs));
  CUDA_CHECK(cudaMalloc(&W, sizeof(float) * (size_t)jobs * KP * KP));
}
void GramWorkspace::free_all() {
  if (job_begin) cudaFree(job_begin);
  if (job_end) cudaFree(job_end);
  if (W) cudaFree(W);
  job_begin = job_end = nullptr;
  W = nullptr;
  max_jobs = 0;
}

void launch_gram_tc(const float *Y, int64_t row_begin, int64_t row_end, float alpha0,
                    const GramWorkspace &ws, float *P, cudaStream_t s) {
  const int64_t n = row_end - row_begin;
  if (n <= 0) {
    CUDA_CHECK(cudaMemsetAsync(P, 0, sizeof(float) * KP * KP, s));
    return;
  }
  // at least 8 stages of work per job, at most one job per workspace slot
  const int n_jobs = (int)std::max<int64_t>(1, std::min<int64_t>(ws.max_jobs, n / (8 * KT)));
  block_jobs_kernel<<<1, 256, 0, s>>>(row_begin, row_end, n_jobs, ws.job_begin, ws.job_end);
  count_launch();
  WGramArgs a{};
  a.Y = Y;
  a.ld = KP;
  a.job_begin = ws.job_begin;
  a.job_end = ws.job_end;
  a.n_jobs = n_jobs;
  a.W = ws.W;
  launch_wgram(a, s);
  launch_wgram_reduce_sym(ws.W, n_jobs, alpha0, P, s);
}

size_t wgram_smem_bytes() { return (size_t)STAGES * kStageBytes + 1024 + 12 * 8 + 16; }

// A/B variant with K-major operand tiles (wgram_k.cu, IALS_WGRAM=kmajor; not measured yet)
bool wgram_kmajor_enabled();
void launch_wgram_kmajor(const WGramArgs &a, cudaStream_t s);

void launch_wgram(const WGramArgs &a, cudaStream_t s) {
  if (a.n_jobs <= 0) return;
  if (a.ld != KP) throw NotImplemented("tensor-core Gram: n_components must pad to 128");
  if (wgram_kmajor_enabled() && a.debug_flags == 0) {  // (no bring-up switches, no TMEM dump there)
    launch_wgram_kmajor(a, s);
    return;
  }
  const size_t smem = wgram_smem_bytes();
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)std::min<int64_t>(a.n_jobs, sms);
  wgram_kernel<<<grid, kThreads, smem, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

void launch_wgram_reduce_sym(const float *W, int n_parts, float scale, float *out, cudaStream_t s) {
  wgram_reduce_sym_kernel<<<(KP * KP + 255) / 256, 256, 0, s>>>(W, n_parts, scale, out);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
