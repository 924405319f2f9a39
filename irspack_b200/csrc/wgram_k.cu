// A/B variant of the weighted gathered Gram (wgram.cu) with K-MAJOR operand tiles.
// Opt-in: IALS_WGRAM=kmajor.  NOT MEASURED YET (written after this round's GPU minutes were
// spent); same interface, same arithmetic, same pipeline as wgram_kernel -- see wgram.cu for the
// algorithm (error-compensated TF32: HH = sum hi hi^T, HL = sum hi lo^T, W = HH/2 + HL) and the
// reference functions it serves (/root/reference/cpp_source/als/IALSTrainer.hpp:78-115, 37-58).
//
// Why: wgram_kernel feeds the tensor core MN-major tiles (a gathered row IS feature-contiguous),
// and the MMA alone then runs at 59 % of the TF32 rate.  Here the producers
// transpose in registers and write the canonical K-major SWIZZLE_128B layout that score_tc.cu
// already uses (tc.cuh desc_kmajor_sw128): a tile row is one FEATURE, 128 bytes = the stage's 32
// neighbours; a stage is 128 rows (hi) + 128 rows (lo) = 32 KB, so the [hi | lo] B operand is
// still one N = 256 descriptor.
//   * a producer warp owns 8 CONSECUTIVE neighbours of the stage (k = 8 pw .. 8 pw + 7) and a
//     lane owns features {l, l+32, l+64, l+96}: four coalesced 128-byte LDG.32 per neighbour
//     (the same four L1 wavefronts as one LDG.128 per lane) -- then the 8 neighbours of one
//     feature sit in one lane's registers and leave as two 16-byte stores per tile;
//   * 16 STS.128 per lane and 8 neighbours, as in wgram_kernel; the lanes of a quarter warp
//     write rows with distinct (row mod 8), i.e. distinct swizzled 16-byte slots: conflict-free.
//
// FUSE instantiation (IALS_WGRAM=fused, also unmeasured): for a heavy row that is ONE job, the
// epilogue warps do not write W to global memory for dense_cg.cu to read back: they keep
// S = HH/2 + HL in shared memory (thread t owns row t, 129-float stride), release the TMEM
// accumulator, symmetrise S in place to A = S + S^T + P + reg_u I, and run the very CG
// recurrences of dense_cg.cu on it while the producers and the MMA warp are already on the next
// job.  The producers' partial sums of b still travel through global memory; a per-job arrival
// counter tells the epilogue when all 16 are there.  Rows cut into several jobs keep the
// W / dense_cg route.
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "tc.cuh"

namespace ials {
namespace {

using namespace tc;

constexpr int KP = 128;
constexpr int KT = 32;  // neighbours per stage = one 128-byte K-major row of tf32
constexpr int STAGES = 4;
constexpr int kProducerWarps = 4;
constexpr int kGroups = 4;
constexpr int kAllProducerWarps = kGroups * kProducerWarps;
static_assert(kAllProducerWarps == kWGramBParts, "bpart layout");
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kAllProducerWarps + kEpilogueWarps + 1) * kWarp;
constexpr int kTileBytes = KP * 128;          // 16 KB: hi or lo, 128 feature rows x 32 neighbours
constexpr int kStageBytes = 2 * kTileBytes;   // 32 KB
constexpr int kTmemCols = 512;
constexpr uint32_t kIdesc = idesc_tf32(KP, 2 * KP, false, false);  // both operands K-major

constexpr int LDS_ = KP + 1;  // row stride of S: row walks and column walks are conflict-free
constexpr size_t kFuseFloats = (size_t)KP * LDS_ + KP + 8;  // S, search direction, reduction scratch

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// barrier of the 4 epilogue warps only (named barrier 1; barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// sum over the 128 epilogue threads, result in every thread (as dense_cg.cu block_sum)
__device__ __forceinline__ float epi_sum(float v, float *scratch, int &phase, int et) {
  v = warp_sum(v);
  float *sc = scratch + 4 * (phase & 1);
  phase++;
  if ((et & 31) == 0) sc[et >> 5] = v;
  epi_bar();
  return (sc[0] + sc[1]) + (sc[2] + sc[3]);
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 16-byte store to a shared-window address (STS.128; a float4 store through the generic pointer
// derived from the aligned dynamic-shared base compiles to generic ST.E pieces)
__device__ __forceinline__ void sts4(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w)
               : "memory");
}

struct StageCursor {  // as in wgram.cu
  int j, base, je;
  unsigned it;
  __device__ __forceinline__ bool valid(const WGramArgs &a) const { return j < (int)a.n_jobs; }
  __device__ __forceinline__ void seek(const WGramArgs &a, int grid) {
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

template <bool FUSE>
__global__ void __launch_bounds__(kThreads, 1) wgram_kmajor_kernel(WGramArgs a, DenseSolveArgs d,
                                                                  unsigned *bcount) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + STAGES * kStageBytes);
  uint64_t *full = bars;
  uint64_t *empty = bars + STAGES;
  uint64_t *accfull = bars + 2 * STAGES;
  uint64_t *accempty = accfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accempty + 2);
  float *Sm = reinterpret_cast<float *>(tiles + STAGES * kStageBytes + 128);  // FUSE only

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
    mbar_init_fence();
  }
  if (warp == kAllProducerWarps + kEpilogueWarps) tmem_alloc(tmem_slot, kTmemCols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kAllProducerWarps) {
    // ================================ PRODUCERS ================================
    const int group = warp / kProducerWarps, pw = warp % kProducerWarps;
    const int grid = (int)gridDim.x;
    int flushed = (int)blockIdx.x - grid;
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};  // features lane + 32 j
    auto flush_until = [&](int j_stop) {
      for (int jj = flushed + grid; jj < j_stop && jj < (int)a.n_jobs; jj += grid) {
        if (a.bpart) {
          float *dst = a.bpart + ((size_t)jj * kAllProducerWarps + warp) * KP + lane;
#pragma unroll
          for (int j = 0; j < 4; j++) dst[32 * j] = bacc[j];
        }
        if (FUSE) {  // this warp's partial of job jj is in global memory: tell the epilogue
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(bcount + jj, 1u);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) bacc[j] = 0.f;
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
    constexpr int NPW = KT / kProducerWarps;  // 8 consecutive neighbours per warp and stage
    static_assert(NPW == 8, "two 16-byte chunks of four neighbours per feature row");
    while (cur.valid(a)) {
      flush_until(cur.j);
      const int s = (int)(cur.it % STAGES);
      const uint32_t ph = (uint32_t)((cur.it / STAGES) & 1);
      const int m = min(KT, cur.je - cur.base);
      float v[NPW][4];
#pragma unroll
      for (int q = 0; q < NPW; q++) {
        const int t = NPW * pw + q;
        const int row = __shfl_sync(0xffffffffu, row0, t);
        const float *src = a.Y + (size_t)row * a.ld + lane;
#pragma unroll
        for (int j = 0; j < 4; j++) v[q][j] = t < m ? __ldg(src + 32 * j) : 0.f;
      }
      const StageCursor n3 = next_own(n2);
      int row3;
      float w3;
      load_ids(n3, row3, w3);

      mbar_wait(&empty[s], ph ^ 1);
      const uint32_t hi = smem_u32(tiles + s * kStageBytes), lo = hi + kTileBytes;
      float sc[NPW];
#pragma unroll
      for (int q = 0; q < NPW; q++) {
        const int t = NPW * pw + q;
        const float w = __shfl_sync(0xffffffffu, w0, t);
        sc[q] = sqrtf(fmaxf(w, 0.f));
        const float cb = t < m ? a.bias + w : 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) bacc[j] = fmaf(cb, v[q][j], bacc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int f = lane + 32 * j;  // tile row = feature
#pragma unroll
        for (int half = 0; half < 2; half++) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int q = 4 * half + e;
            const float u = sc[q] * v[q][j];
            h[e] = __uint_as_float(__float_as_uint(u) & 0xffffe000u);
            l[e] = u - h[e];
          }
          const uint32_t off = sw128_offset(f, 2 * pw + half);  // 16-byte chunk = 4 neighbours
          sts4(hi + off, h[0], h[1], h[2], h[3]);
          sts4(lo + off, l[0], l[1], l[2], l[3]);
        }
      }
      fence_proxy_async_smem();
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
      if (je <= jb) continue;
      const int buf = (int)(jc & 1);
      mbar_wait(&accempty[buf], (uint32_t)(((jc >> 1) & 1) ^ 1));
      fence_after();
      const uint32_t d = tmem_base + (uint32_t)(buf * 256);
      uint32_t acc = 0;
      for (long long base = jb; base < je; base += KT, it++) {
        const int s = (int)(it % STAGES);
        mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
        fence_after();
        if (lane == 0) {
          const uint32_t hi = smem_u32(tiles + s * kStageBytes);
#pragma unroll
          for (int k = 0; k < KT / 8; k++) {
            // A = hi rows 0..127, B = rows 0..255 of the stage ([hi | lo]); 8 neighbours = 32 bytes
            const uint64_t dh = desc_kmajor_sw128(hi + k * 32);
            mma_tf32(d, dh, dh, acc, kIdesc);
            acc = 1;
          }
          commit(&empty[s]);
          if (base + KT >= je) commit(&accfull[buf]);
        }
        __syncwarp();
      }
      jc++;
    }
  } else {
    // ================================ EPILOGUE ================================ (as wgram.cu)
    const int ew = warp - kAllProducerWarps;
    const int row = ew * 32 + lane;
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
      fence_after();
      const uint32_t t_hh = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 256);
      if (FUSE) {
        // heavy row of this job: heavy_first_job[h] <= j < heavy_first_job[h + 1]
        int lo_h = 0, hi_h = (int)d.n_heavy;
        while (hi_h - lo_h > 1) {
          const int mid = (lo_h + hi_h) >> 1;
          if (d.heavy_first_job[mid] <= (int)j) lo_h = mid; else hi_h = mid;
        }
        const int h = lo_h;
        if (d.heavy_first_job[h + 1] - d.heavy_first_job[h] == 1) {
          const SolveArgs &sa = d.base;
          float *pv = Sm + KP * LDS_, *red = pv + KP;
          epi_bar();  // every reader of the previous row's S / pv is done
#pragma unroll 1
          for (int c = 0; c < KP; c += 16) {
            uint32_t hh[16], hl[16];
            tmem_ld16(t_hh + c, hh);
            tmem_ld16(t_hh + 128 + c, hl);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; q++)
              Sm[row * LDS_ + c + q] = fmaf(0.5f, __uint_as_float(hh[q]), __uint_as_float(hl[q]));
          }
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&accempty[buf]);  // the MMA warp may reuse the accumulator
          jc++;
          // b = sum of the 16 producer warps' partials (global memory, arrival counter)
          if (lane == 0) {
            unsigned spins = 0;
            while (ld_acquire_u32(bcount + j) < (unsigned)kAllProducerWarps) {
              if (++spins > (1u << 24)) __trap();
            }
          }
          __syncwarp();
          float b = 0.f;
#pragma unroll
          for (int q = 0; q < kAllProducerWarps; q++)
            b += __ldcg(a.bpart + ((size_t)j * kAllProducerWarps + q) * KP + row);
          const int64_t u = sa.order[h];
          const int64_t gu = sa.row_base + u;
          const int64_t nnz = sa.indptr[u + 1] - sa.indptr[u];
          const float reg_u = sa.reg * powf(sa.alpha0 * (float)sa.n_other + (float)nnz, sa.nu);  // :117-120
          float x = sa.target[gu * KP + row];
          pv[row] = x;
          epi_bar();  // S is complete
          // A = S + S^T + P + reg_u I in place.  The unordered pair {i, j} belongs to thread i
          // with (j - i) mod 128 in [1, 63], or 64 for i < 64: 63.5 pairs per thread, no races.
          for (int k = 1; k <= KP / 2; k++) {
            if (k == KP / 2 && row >= KP / 2) break;
            const int c = (row + k) & (KP - 1);
            const float v = (Sm[row * LDS_ + c] + Sm[c * LDS_ + row]) + __ldg(sa.P + (size_t)row * KP + c);
            Sm[row * LDS_ + c] = v;
            Sm[c * LDS_ + row] = v;
          }
          Sm[row * LDS_ + row] = fmaf(2.f, Sm[row * LDS_ + row], __ldg(sa.P + (size_t)row * KP + row)) + reg_u;
          epi_bar();  // A and pv are complete
          int phase = 0;
          auto matvec = [&]() {  // (A pv)[row]
            const float *arow = Sm + row * LDS_;
            float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
            for (int c = 0; c < KP; c += 4) {
              const float4 p4 = *reinterpret_cast<const float4 *>(pv + c);
              acc0 = fmaf(arow[c + 0], p4.x, acc0);
              acc1 = fmaf(arow[c + 1], p4.y, acc1);
              acc2 = fmaf(arow[c + 2], p4.z, acc2);
              acc3 = fmaf(arow[c + 3], p4.w, acc3);
            }
            return (acc0 + acc1) + (acc2 + acc3);
          };
          float r = b - matvec();  // r = b - A x   (IALSTrainer.hpp:216-228)
          float p = r;
          bool failed = false;
          for (int it = 0; it < sa.max_cg_steps; it++) {
            const float r2 = epi_sum(r * r, red, phase, row);
            if (r2 <= 1e-20f) break;  // :237-240
            epi_bar();                // everyone has consumed the previous pv
            pv[row] = p;
            epi_bar();
            const float Ap = matvec();
            const float den = epi_sum(p * Ap, red, phase, row);
            if (!(den > 0.f) || !isfinite(den)) { failed = true; break; }  // :249-254
            const float alpha = r2 / den;
            x = fmaf(alpha, p, x);
            r = fmaf(-alpha, Ap, r);
            const float r2n = epi_sum(r * r, red, phase, row);
            if (r2n <= 1e-20f) break;  // :258-260
            p = fmaf(r2n / r2, p, r);
          }
          if (failed) {
            if (row == 0) atomicExch(&sa.err_flags[kErrCgSingular], 1);
          } else {  // the reference throws before writing the row back
            sa.target[gu * KP + row] = x;
            for (int pi = 0; pi < sa.n_peers; pi++) sa.peers[pi][gu * KP + row] = x;
          }
          continue;
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
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[buf]);
      jc++;
    }
  }

  fence_before();
  __syncthreads();
  if (warp == kAllProducerWarps + kEpilogueWarps) tmem_dealloc(tmem_base, kTmemCols);
}


// ---------------------------------------------------------------------------------------------
// Cross block of a 256-column Gram (K = 256 Cholesky, BASELINE configs[2]; IALS_CHOL=tc, also
// unmeasured).  A 256-float factor row is two halves y0 | y1; the diagonal blocks
// G00 = sum c y0 y0^T and G11 are wgram_kmajor_kernel<false> run on Y and on Y + 128 with
// ld = 256.  This kernel forms  G01 = sum c y0 y1^T  (not symmetric):
//     u = sqrt(c) y,  u0 u1^T = hi0 hi1^T + hi0 lo1^T + lo0 hi1^T + O(2^-22)
// with four K-major tiles per stage [hi0 | lo0 | hi1 | lo1] (64 KB, 3 stages) and two
// instructions per 8 neighbours:  D0 (256 TMEM columns) += hi0 x [hi1 | lo1],
// D1 (128 columns) += lo0 x hi1.  One accumulator set (384 of 512 columns): the epilogue of a
// job and the MMAs of the next one do not overlap here.
// ---------------------------------------------------------------------------------------------
constexpr int XSTAGES = 3;
constexpr int kXStageBytes = 4 * kTileBytes;  // 64 KB
constexpr uint32_t kIdescN128 = idesc_tf32(KP, KP, false, false);

__global__ void __launch_bounds__(kThreads, 1) wgram_cross_kernel(WGramArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + XSTAGES * kXStageBytes);
  uint64_t *full = bars;               // [XSTAGES]
  uint64_t *empty = bars + XSTAGES;    // [XSTAGES]
  uint64_t *accfull = bars + 2 * XSTAGES;
  uint64_t *accempty = accfull + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accempty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < XSTAGES; s++) {
      mbar_init(&full[s], kProducerWarps);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accfull, 1);
    mbar_init(accempty, kEpilogueWarps);
    mbar_init_fence();
  }
  if (warp == kAllProducerWarps + kEpilogueWarps) tmem_alloc(tmem_slot, kTmemCols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kAllProducerWarps) {
    // producers: as wgram_kmajor_kernel, both halves of every neighbour row, no b
    const int group = warp / kProducerWarps, pw = warp % kProducerWarps;
    const int grid = (int)gridDim.x;
    auto next_own = [&](StageCursor c) {
      for (int g = 0; g < kGroups && c.valid(a); g++) c.advance(a, grid);
      return c;
    };
    auto load_ids = [&](const StageCursor &c, int &row, float &w) {
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
    StageCursor n1 = next_own(cur);
    int row0, row1;
    float w0, w1;
    load_ids(cur, row0, w0);
    load_ids(n1, row1, w1);
    constexpr int NPW = KT / kProducerWarps;
    while (cur.valid(a)) {
      const int s = (int)(cur.it % XSTAGES);
      const uint32_t ph = (uint32_t)((cur.it / XSTAGES) & 1);
      const int m = min(KT, cur.je - cur.base);
      const StageCursor n2 = next_own(n1);
      int row2;
      float w2;
      load_ids(n2, row2, w2);
      float sc[NPW];
#pragma unroll
      for (int q = 0; q < NPW; q++) sc[q] = sqrtf(fmaxf(__shfl_sync(0xffffffffu, w0, NPW * pw + q), 0.f));
      bool waited = false;
#pragma unroll 1
      for (int hf = 0; hf < 2; hf++) {  // feature half: columns [128 hf, 128 hf + 128) of the row
        float v[NPW][4];
#pragma unroll
        for (int q = 0; q < NPW; q++) {
          const int t = NPW * pw + q;
          const int row = __shfl_sync(0xffffffffu, row0, t);
          const float *src = a.Y + (size_t)row * a.ld + KP * hf + lane;
#pragma unroll
          for (int j = 0; j < 4; j++) v[q][j] = t < m ? __ldg(src + 32 * j) : 0.f;
        }
        if (!waited) {
          mbar_wait(&empty[s], ph ^ 1);
          waited = true;
        }
        const uint32_t hi = smem_u32(tiles + s * kXStageBytes) + (uint32_t)(2 * hf) * kTileBytes;
        const uint32_t lo = hi + kTileBytes;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int f = lane + 32 * j;
#pragma unroll
          for (int half = 0; half < 2; half++) {
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int q = 4 * half + e;
              const float u = sc[q] * v[q][j];
              h[e] = __uint_as_float(__float_as_uint(u) & 0xffffe000u);
              l[e] = u - h[e];
            }
            const uint32_t off = sw128_offset(f, 2 * pw + half);
            sts4(hi + off, h[0], h[1], h[2], h[3]);
            sts4(lo + off, l[0], l[1], l[2], l[3]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      cur = n1; n1 = n2;
      row0 = row1; row1 = row2;
      w0 = w1; w1 = w2;
    }
  } else if (warp == kAllProducerWarps + kEpilogueWarps) {
    // MMA issuer
    unsigned long long it = 0, jc = 0;
    for (long long j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
      const long long jb = a.job_begin[j], je = a.job_end[j];
      if (je <= jb) continue;
      mbar_wait(accempty, (uint32_t)((jc & 1) ^ 1));
      fence_after();
      uint32_t acc = 0;
      for (long long base = jb; base < je; base += KT, it++) {
        const int s = (int)(it % XSTAGES);
        mbar_wait(&full[s], (uint32_t)((it / XSTAGES) & 1));
        fence_after();
        if (lane == 0) {
          const uint32_t hi0 = smem_u32(tiles + s * kXStageBytes);
          const uint32_t lo0 = hi0 + kTileBytes, hi1 = hi0 + 2 * kTileBytes;
#pragma unroll
          for (int k = 0; k < KT / 8; k++) {
            const uint64_t db = desc_kmajor_sw128(hi1 + k * 32);  // rows 0..255: hi1 then lo1
            mma_tf32(tmem_base, desc_kmajor_sw128(hi0 + k * 32), db, acc, kIdesc);
            mma_tf32(tmem_base + 256, desc_kmajor_sw128(lo0 + k * 32), db, acc, kIdescN128);
            acc = 1;
          }
          commit(&empty[s]);
          if (base + KT >= je) commit(accfull);
        }
        __syncwarp();
      }
      jc++;
    }
  } else {
    // epilogue: G01 row `row` = D0[0:128] + D0[128:256] + D1[0:128]
    const int ew = warp - kAllProducerWarps;
    const int row = ew * 32 + lane;
    unsigned long long jc = 0;
    for (long long j = blockIdx.x; j < a.n_jobs; j += gridDim.x) {
      float *out = a.W + (size_t)j * KP * KP + (size_t)row * KP;
      if (a.job_end[j] <= a.job_begin[j]) {
#pragma unroll 4
        for (int c = 0; c < KP; c += 4) *reinterpret_cast<float4 *>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      mbar_wait(accfull, (uint32_t)(jc & 1));
      fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < KP; c += 16) {
        uint32_t x0[16], x1[16], x2[16];
        tmem_ld16(t0 + c, x0);
        tmem_ld16(t0 + 128 + c, x1);
        tmem_ld16(t0 + 256 + c, x2);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          float4 o;
          o.x = (__uint_as_float(x0[q + 0]) + __uint_as_float(x1[q + 0])) + __uint_as_float(x2[q + 0]);
          o.y = (__uint_as_float(x0[q + 1]) + __uint_as_float(x1[q + 1])) + __uint_as_float(x2[q + 1]);
          o.z = (__uint_as_float(x0[q + 2]) + __uint_as_float(x1[q + 2])) + __uint_as_float(x2[q + 2]);
          o.w = (__uint_as_float(x0[q + 3]) + __uint_as_float(x1[q + 3])) + __uint_as_float(x2[q + 3]);
          *reinterpret_cast<float4 *>(out + c + q) = o;
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);
      jc++;
    }
  }
  fence_before();
  __syncthreads();
  if (warp == kAllProducerWarps + kEpilogueWarps) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// IALS_WGRAM=kmajor | fused (read once).  launch_wgram (wgram.cu) asks this before its own launch.
static int wgram_variant() {
  static const int v = [] {
    const char *e = std::getenv("IALS_WGRAM");
    const std::string m = e ? e : "";
    return m == "kmajor" ? 1 : (m == "fused" ? 2 : 0);
  }();
  return v;
}
bool wgram_kmajor_enabled() { return wgram_variant() != 0; }
bool wgram_fused_enabled() { return wgram_variant() == 2; }

namespace {
constexpr size_t kSmemPlain = (size_t)STAGES * kStageBytes + 1024 + 12 * 8 + 16;
constexpr size_t kSmemFused = (size_t)STAGES * kStageBytes + 1024 + 128 + sizeof(float) * kFuseFloats;
static_assert(kSmemFused <= 232448, "fused epilogue does not fit shared memory");

unsigned grid_for(int64_t n_jobs) {
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return (unsigned)std::min<int64_t>(n_jobs, sms);
}

// rows are sorted by descending degree, so the rows cut into several jobs come first
__global__ void count_multi_job_rows_kernel(const int32_t *first, int n_heavy, int *out) {
  int lo = 0, hi = n_heavy;  // first h in [0, n_heavy] whose row is a single job
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (first[mid + 1] - first[mid] > 1) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}
}  // namespace

void launch_wgram_kmajor(const WGramArgs &a, cudaStream_t s) {
  if (a.n_jobs <= 0) return;
  if (a.ld != KP) throw NotImplemented("tensor-core Gram: n_components must pad to 128");
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_kmajor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmemPlain));
  });
  wgram_kmajor_kernel<false><<<grid_for(a.n_jobs), kThreads, kSmemPlain, s>>>(a, DenseSolveArgs{}, nullptr);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

// Symmetric block of a wider factor matrix: a.Y may point into the row (Y + 128), a.ld is the
// true row stride (256).  W / bpart as in launch_wgram.
void launch_wgram_kmajor_strided(const WGramArgs &a, cudaStream_t s) {
  if (a.n_jobs <= 0) return;
  if (a.ld < KP || a.ld % 4 != 0) throw InvalidArgument("strided Gram: row stride must be >= 128");
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_kmajor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmemPlain));
  });
  wgram_kmajor_kernel<false><<<grid_for(a.n_jobs), kThreads, kSmemPlain, s>>>(a, DenseSolveArgs{}, nullptr);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

// Cross block G01 = sum c y[0:128] y[128:256]^T of rows with stride a.ld >= 256; a.W receives
// the full 128 x 128 block per job (no symmetrisation), a.bpart is not written.
void launch_wgram_cross(const WGramArgs &a, cudaStream_t s) {
  if (a.n_jobs <= 0) return;
  if (a.ld < 2 * KP || a.ld % 4 != 0) throw InvalidArgument("cross Gram: row stride must be >= 256");
  constexpr size_t smem = (size_t)XSTAGES * kXStageBytes + 1024 + 128;
  static_assert(smem <= 232448, "cross Gram stages do not fit shared memory");
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  wgram_cross_kernel<<<grid_for(a.n_jobs), kThreads, smem, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

// Heavy rows of one half-epoch: Gram of every job, CG of the single-job rows in the epilogue.
// Returns the number of leading heavy rows (those cut into several jobs) whose W / bpart were
// written for dense_cg.cu.  d.W / d.bpart must be the buffers a.W / a.bpart point to.
int64_t launch_wgram_fused(const WGramArgs &a, const DenseSolveArgs &d, cudaStream_t s) {
  if (a.n_jobs <= 0 || d.n_heavy <= 0) return 0;
  if (a.ld != KP) throw NotImplemented("tensor-core Gram: n_components must pad to 128");
  if (a.bpart == nullptr || a.W == nullptr) throw InvalidArgument("fused heavy path needs W and bpart");
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_kmajor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmemFused));
  });
  unsigned *bcount = nullptr;
  int *d_multi = nullptr;
  CUDA_CHECK(cudaMallocAsync(&bcount, sizeof(unsigned) * a.n_jobs + sizeof(int), s));
  d_multi = reinterpret_cast<int *>(bcount + a.n_jobs);
  int multi = 0;
  try {
    CUDA_CHECK(cudaMemsetAsync(bcount, 0, sizeof(unsigned) * a.n_jobs + sizeof(int), s));
    count_multi_job_rows_kernel<<<1, 1, 0, s>>>(d.heavy_first_job, (int)d.n_heavy, d_multi);
    count_launch();
    wgram_kmajor_kernel<true><<<grid_for(a.n_jobs), kThreads, kSmemFused, s>>>(a, d, bcount);
    count_launch();
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(&multi, d_multi, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));  // (A/B variant: one host round trip per half-epoch)
  } catch (...) {
    cudaFreeAsync(bcount, s);
    throw;
  }
  CUDA_CHECK(cudaFreeAsync(bcount, s));
  return multi;
}

}  // namespace ials
