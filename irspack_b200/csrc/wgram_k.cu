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

__global__ void __launch_bounds__(kThreads, 1) wgram_kmajor_kernel(WGramArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>(
      ((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(tiles + STAGES * kStageBytes);
  uint64_t *full = bars;
  uint64_t *empty = bars + STAGES;
  uint64_t *accfull = bars + 2 * STAGES;
  uint64_t *accempty = accfull + 2;
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

}  // namespace

// IALS_WGRAM=kmajor (read once).  launch_wgram (wgram.cu) asks this before its own launch.
bool wgram_kmajor_enabled() {
  static const bool on = [] {
    const char *e = std::getenv("IALS_WGRAM");
    return e != nullptr && std::string(e) == "kmajor";
  }();
  return on;
}

void launch_wgram_kmajor(const WGramArgs &a, cudaStream_t s) {
  if (a.n_jobs <= 0) return;
  if (a.ld != KP) throw NotImplemented("tensor-core Gram: n_components must pad to 128");
  const size_t smem = (size_t)STAGES * kStageBytes + 1024 + 12 * 8 + 16;
  static PerDeviceOnce configured;
  configured.run([&] {
    CUDA_CHECK(cudaFuncSetAttribute(wgram_kmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  int dev = 0, sms = kNumSMsB200;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)std::min<int64_t>(a.n_jobs, sms);
  wgram_kmajor_kernel<<<grid, kThreads, smem, s>>>(a);
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
