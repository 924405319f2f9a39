// K3 Cholesky row solve for 256-column factors, after the tensor-core Gram (replaces
// Solver::step_cholesky, /root/reference/cpp_source/als/IALSTrainer.hpp:273-331, for the rows whose
// rank updates -- BatchedRankUpdater, :37-58 -- were done by wgram.cu; api.cu solve_cholesky_tensor).
//
// Per row:  A = P + G + reg_u I  (G = sum c y y^T from the Gram workspace),  A = U^T U,
// U^T z = b,  U x = z.
//
// profiles/r02h: the register-tiled kernel (cholesky_tile.cu, one CTA of 544 threads = one row per
// SM, the matrix in registers, the factor in 131 KB of shared memory) spent 410 k cycles per row
// where the FMA work needs 11 k: one row per SM means every serial chain -- the 8-pivot diagonal
// eliminations, the 64 block barriers, the one-warp backward substitution -- is exposed.  Here the
// factor lives in a per-CTA scratch in GLOBAL memory (L2-resident: 152 KB per CTA), a CTA is 160
// threads and holds only ONE BLOCK ROW (32 x 256) of the matrix in registers, so three to four
// rows are in flight per SM and the chains of one row hide behind the FMAs of the others:
//   * blocked LEFT-looking factorisation, 8 block rows of 32: block row p starts from zero,
//     accumulates  S = sum_{k < 32 p} U(k, rows)^T U(k, cols)  from the finished block rows, which
//     are streamed from the scratch through shared memory (cp.async, 16 pivots per stage, double
//     buffered: 4 LDS.128 per 32 packed FFMA2), then takes  A - S  (A read once from the Gram
//     workspace + P, L2-prefetched while the stages run);
//   * thread (tr, tc) = (t & 3, t >> 2) owns the 8 x 8 tile (tr, tc) of the block row: 4 tile rows
//     x 32 tile columns = 128 threads, threads 128..131 own the right-hand side as a 33rd tile
//     column (column 0 = b, which makes the forward substitution part of the factorisation);
//     the tile columns left of the diagonal block do not exist, so the live threads of the later
//     block rows fill whole warps and the idle warps cost nothing;
//   * inside the block row, 4 steps of 8 pivots: the owner of the diagonal tile factors it in
//     registers (pivot row scaled by 1 / sqrt(a_ii): the published rows ARE the rows of U), the
//     tile row applies the same row operations with the diagonal tile's entries and publishes its
//     8 final rows to shared memory and to the scratch, the tile rows below take the rank-8
//     update;
//   * backward substitution block by block from the bottom: a 32 x rem matrix-vector product by
//     128 threads, then the 32 x 32 triangle by one warp with its rows in registers (one shuffle
//     and one FMA per unknown).
// Factor columns >= K (K < 256) are zero in P and G: their diagonal is set to 1, their solution
// is 0.  Failure rules of the reference: pivot not > 0 -> "Cholesky decomposition failed.",
// non-finite solution -> "Cholesky solve failed." (:316-323).
#include "common.cuh"

namespace ials {
namespace {

constexpr int kN = 256;            // order of the (padded) system = row stride of the factors
constexpr int kNB = 32;            // block row height
constexpr int kBlocks = kN / kNB;  // 8
constexpr int kLLThreads = 160;
constexpr int kStagePivots = 16;
constexpr int kStripLd = 232;  // staged rows: block rows p >= 1 are at most 224 + 8 floats wide
constexpr int kProwLd = 264;

// scratch of one CTA: block row q = 32 rows of row_len(q) floats  [U(k, 32 q .. 255) | z_k | 1 / U_kk | 0 x 6]
__host__ __device__ constexpr int row_len(int q) { return kN - kNB * q + 8; }
__host__ __device__ constexpr int block_base(int q) { return 32 * (264 * q - 16 * q * (q - 1)); }
constexpr int kScratchFloats = block_base(kBlocks);  // 38 912

#define EL(i, j) (((j) & 1) ? acc[i][(j) >> 1].y : acc[i][(j) >> 1].x)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long *>(&d))
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
        "l"(*reinterpret_cast<unsigned long long *>(&c)));
  return d;
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct LLSmem {
  float strip[2][kStagePivots][kStripLd];  // staged rows of a finished block row (also: U_pp of the back substitution)
  float prow[8][kProwLd];                  // the 8 rows published by the current step
  float dtile[64];                         // the factored diagonal tile (row li, column r)
  float dinv[8];                           // 1 / U_ii of its pivots
  float x[kN];                             // solution
  float rs[kNB];                           // right-hand side of the current triangle
  float sinv[kNB];
};

// Workspace of one chunk of Gram jobs (api.cu solve_cholesky_tensor):
//   W00 [JC][128][128] | W11 [JC][128][128] | G01 [JC][128][128] | b0 [JC][16][128] | b1 likewise
// diagonal blocks: G = W + W^T; G01: rows in the first half of the factor, columns in the second.
struct LLArgs {
  const float *ws;
  int job0;        // first job of the chunk (row_jobs holds absolute job ids)
  int job_cap;     // JC
  float *scratch;  // [gridDim.x][kScratchFloats]
};

__global__ void __launch_bounds__(kLLThreads, 3) cholesky_ll_kernel(SolveArgs a, LLArgs g) {
  __shared__ __align__(16) LLSmem sm;
  __shared__ long long s_slot;
  __shared__ int s_fail;
  const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
  const int tr = tid & 3, tcr = tid >> 2;  // tile row / tile column (relative to the block row); 32 = rhs
  const bool rhs = tcr == 32;
  const int K = a.K;
  float *const scratch = g.scratch + (size_t)blockIdx.x * kScratchFloats;
  const size_t blk = (size_t)g.job_cap * 128 * 128;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_slot = (long long)atomicAdd(a.work_counter, 1ull);
      s_fail = 0;
    }
    __syncthreads();
    const int64_t slot = s_slot;
    if (slot >= a.n_sched) break;
    const int64_t u = a.order ? (int64_t)a.order[slot] : slot;
    const int64_t gu = a.row_base + u;
    const int64_t nnz = a.indptr[u + 1] - a.indptr[u];
    const float reg_u = a.reg * powf(a.alpha0 * (float)a.n_other + (float)nnz, a.nu);  // :309-310
    const int gj0 = a.row_jobs[slot] - g.job0, gj1 = a.row_jobs[slot + 1] - g.job0;
    bool failed = false;

    for (int p = 0; p < kBlocks && !failed; p++) {
      const int Wp = kN - kNB * p, Lp = Wp + 8, ntc = Wp / 8;
      const bool active = rhs || (tcr < ntc && (tcr >= 4 || tcr >= tr));
      const int colofs = rhs ? Wp : 8 * tcr;
      const int gi0 = kNB * p + 8 * tr, gjc = kNB * p + 8 * tcr;  // global row / column of the tile
      const int bi = gi0 >> 7, bj = gjc >> 7, li0 = gi0 & 127, lj0 = gjc & 127;

      // warp 4, lane = row of the block row: b = the Gram producers' partial sums (:301-307)
      float bl = 0.f;
      if (warp == 4) {
        const int row = kNB * p + lane;
        const float *bp = g.ws + 3 * blk + (size_t)(row >> 7) * g.job_cap * kWGramBParts * 128 + (row & 127);
        for (int jb = gj0; jb < gj1; jb++)
#pragma unroll
          for (int q = 0; q < kWGramBParts; q++) bl += bp[((size_t)jb * kWGramBParts + q) * 128];
      }
      // the tile's part of A: into L2 while the finished block rows stream through
      if (active && !rhs && p > 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          prefetch_l2(a.P + (size_t)(gi0 + i) * kN + gjc);
          if (bi == bj) {
            prefetch_l2(g.ws + (size_t)bi * blk + (size_t)gj0 * 16384 + (li0 + i) * 128 + lj0);
            prefetch_l2(g.ws + (size_t)bi * blk + (size_t)gj0 * 16384 + (lj0 + i) * 128 + li0);
          } else {
            prefetch_l2(g.ws + 2 * blk + (size_t)gj0 * 16384 + (li0 + i) * 128 + lj0);
          }
        }
      }

      float2 acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

      // ---- S = sum over the finished block rows, 16 pivots per stage ----
      const int n_stages = 2 * p, chunks_per_row = Lp / 4;
      auto issue_stage = [&](int st) {
        const int q = st >> 1, k0 = kStagePivots * (st & 1);
        const float *src = scratch + block_base(q) + (size_t)k0 * row_len(q) + kNB * (p - q);
        float *dst = &sm.strip[st & 1][0][0];
        for (int c = tid; c < kStagePivots * chunks_per_row; c += kLLThreads) {
          const int k = c / chunks_per_row, cc = c - k * chunks_per_row;
          cp_async16(dst + k * kStripLd + 4 * cc, src + (size_t)k * row_len(q) + 4 * cc);
        }
      };
      if (n_stages > 0) issue_stage(0);
      for (int st = 0; st < n_stages; st++) {
        cp_async_commit_wait_all();
        __syncthreads();  // stage st has landed; everybody is done with the other buffer
        if (st + 1 < n_stages) issue_stage(st + 1);
        if (active) {
          const float *sb = &sm.strip[st & 1][0][0];
#pragma unroll 4
          for (int k = 0; k < kStagePivots; k++) {
            const float4 r0 = *reinterpret_cast<const float4 *>(sb + k * kStripLd + 8 * tr);
            const float4 r1 = *reinterpret_cast<const float4 *>(sb + k * kStripLd + 8 * tr + 4);
            const float4 c0 = *reinterpret_cast<const float4 *>(sb + k * kStripLd + colofs);
            const float4 c1 = *reinterpret_cast<const float4 *>(sb + k * kStripLd + colofs + 4);
            const float ra[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            const float2 cb[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y),
                                  make_float2(c1.z, c1.w)};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
              for (int j = 0; j < 4; j++) acc[i][j] = fma2(make_float2(ra[i], ra[i]), cb[j], acc[i][j]);
          }
        }
      }

      // ---- acc = A - S ----
      if (warp == 4) {  // the rhs threads are lanes 0..3 of warp 4
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float bv = __shfl_sync(0xffffffffu, bl, 8 * (lane & 3) + i);
          if (rhs) EL(i, 0) = bv - EL(i, 0);
        }
      }
      if (active && !rhs) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float4 p0 = *reinterpret_cast<const float4 *>(a.P + (size_t)(gi0 + i) * kN + gjc);
          const float4 p1 = *reinterpret_cast<const float4 *>(a.P + (size_t)(gi0 + i) * kN + gjc + 4);
          acc[i][0] = make_float2(p0.x - acc[i][0].x, p0.y - acc[i][0].y);
          acc[i][1] = make_float2(p0.z - acc[i][1].x, p0.w - acc[i][1].y);
          acc[i][2] = make_float2(p1.x - acc[i][2].x, p1.y - acc[i][2].y);
          acc[i][3] = make_float2(p1.z - acc[i][3].x, p1.w - acc[i][3].y);
        }
        for (int jb = gj0; jb < gj1; jb++) {
          const float *W = g.ws + (bi == bj ? (size_t)bi * blk : 2 * blk) + (size_t)jb * 16384;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 w0 = __ldcs(reinterpret_cast<const float4 *>(W + (li0 + i) * 128 + lj0));
            const float4 w1 = __ldcs(reinterpret_cast<const float4 *>(W + (li0 + i) * 128 + lj0 + 4));
            acc[i][0].x += w0.x; acc[i][0].y += w0.y; acc[i][1].x += w0.z; acc[i][1].y += w0.w;
            acc[i][2].x += w1.x; acc[i][2].y += w1.y; acc[i][3].x += w1.z; acc[i][3].y += w1.w;
          }
          if (bi == bj) {  // + W^T: row lj0 + j of W holds column j of the tile
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 w0 = __ldcs(reinterpret_cast<const float4 *>(W + (lj0 + j) * 128 + li0));
              const float4 w1 = __ldcs(reinterpret_cast<const float4 *>(W + (lj0 + j) * 128 + li0 + 4));
              EL(0, j) += w0.x; EL(1, j) += w0.y; EL(2, j) += w0.z; EL(3, j) += w0.w;
              EL(4, j) += w1.x; EL(5, j) += w1.y; EL(6, j) += w1.z; EL(7, j) += w1.w;
            }
          }
        }
        if (gi0 == gjc) {  // :312-314; the padding columns get a unit diagonal
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (gi0 + i < K) EL(i, i) += reg_u;
            else EL(i, i) = 1.f;
          }
        }
      }

      // ---- the block row itself: 4 steps of 8 pivots ----
      for (int s = 0; s < 4; s++) {
        if (tr == s && tcr == s) {  // factor the diagonal tile in registers
#pragma unroll
          for (int li = 0; li < 8; li++) {
            const float d = EL(li, li);
            if (!(d > 0.f)) s_fail = 1;
            const float inv = __frsqrt_rn(d);
            sm.dinv[li] = inv;
#pragma unroll
            for (int j = 0; j < 4; j++) acc[li][j] = make_float2(acc[li][j].x * inv, acc[li][j].y * inv);
#pragma unroll
            for (int r = 0; r < 8; r++) {
              if (r > li) {
                const float m = EL(li, r);
                sm.dtile[li * 8 + r] = m;
#pragma unroll
                for (int j = 0; j < 4; j++) acc[r][j] = fma2(make_float2(-m, -m), acc[li][j], acc[r][j]);
              }
            }
          }
        }
        __syncthreads();
        if (s_fail) {  // uniform: written before the barrier every thread just passed
          failed = true;
          break;
        }
        if (active && tr == s) {
          if (tcr != s) {  // the same row operations with the diagonal tile's entries
#pragma unroll
            for (int li = 0; li < 8; li++) {
              const float inv = sm.dinv[li];
#pragma unroll
              for (int j = 0; j < 4; j++) acc[li][j] = make_float2(acc[li][j].x * inv, acc[li][j].y * inv);
#pragma unroll
              for (int r = 0; r < 8; r++) {
                if (r > li) {
                  const float m = sm.dtile[li * 8 + r];
#pragma unroll
                  for (int j = 0; j < 4; j++) acc[r][j] = fma2(make_float2(-m, -m), acc[li][j], acc[r][j]);
                }
              }
            }
          }
          float *gdst = scratch + block_base(p) + (size_t)(8 * s) * Lp + colofs;
#pragma unroll
          for (int li = 0; li < 8; li++) {
            float4 v0 = make_float4(acc[li][0].x, acc[li][0].y, acc[li][1].x, acc[li][1].y);
            float4 v1 = make_float4(acc[li][2].x, acc[li][2].y, acc[li][3].x, acc[li][3].y);
            if (rhs) {  // [z | 1 / U_ii | 0 ...]
              v0 = make_float4(v0.x, sm.dinv[li], 0.f, 0.f);
              v1 = make_float4(0.f, 0.f, 0.f, 0.f);
            } else if (tcr == s) {  // below the diagonal: zero
              if (li > 0) v0.x = 0.f;
              if (li > 1) v0.y = 0.f;
              if (li > 2) v0.z = 0.f;
              if (li > 3) v0.w = 0.f;
              if (li > 4) v1.x = 0.f;
              if (li > 5) v1.y = 0.f;
              if (li > 6) v1.z = 0.f;
            }
            *reinterpret_cast<float4 *>(&sm.prow[li][colofs]) = v0;
            *reinterpret_cast<float4 *>(&sm.prow[li][colofs + 4]) = v1;
            *reinterpret_cast<float4 *>(gdst + (size_t)li * Lp) = v0;
            *reinterpret_cast<float4 *>(gdst + (size_t)li * Lp + 4) = v1;
          }
        }
        __syncthreads();
        if (active && tr > s) {  // rank-8 update of the tile rows below
#pragma unroll
          for (int li = 0; li < 8; li++) {
            const float4 r0 = *reinterpret_cast<const float4 *>(&sm.prow[li][8 * tr]);
            const float4 r1 = *reinterpret_cast<const float4 *>(&sm.prow[li][8 * tr + 4]);
            const float4 c0 = *reinterpret_cast<const float4 *>(&sm.prow[li][colofs]);
            const float4 c1 = *reinterpret_cast<const float4 *>(&sm.prow[li][colofs + 4]);
            const float ra[8] = {-r0.x, -r0.y, -r0.z, -r0.w, -r1.x, -r1.y, -r1.z, -r1.w};
            const float2 cb[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y),
                                  make_float2(c1.z, c1.w)};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
              for (int j = 0; j < 4; j++) acc[i][j] = fma2(make_float2(ra[i], ra[i]), cb[j], acc[i][j]);
          }
        }
      }
    }
    __syncthreads();  // the last block row is in the scratch
    if (failed) {
      if (tid == 0) atomicExch(&a.err_flags[kErrCholDecomp], 1);
      continue;
    }

    // ---- backward substitution  U x = z, block by block from the bottom ----
    for (int p = kBlocks - 1; p >= 0; p--) {
      const int Wp = kN - kNB * p, Lp = Wp + 8;
      const float *Ub = scratch + block_base(p);
      if (tid < 128) {  // 4 threads per row: z - U(row, columns right of the diagonal block) . x
        const int row = tid >> 2, part = tid & 3;
        const float *Ur = Ub + (size_t)row * Lp;
        float dot = 0.f;
        for (int c4 = kNB / 4 + part; c4 < Wp / 4; c4 += 4) {
          const float4 uv = __ldcg(reinterpret_cast<const float4 *>(Ur + 4 * c4));
          const float4 xv = *reinterpret_cast<const float4 *>(&sm.x[kNB * p + 4 * c4]);
          dot = fmaf(uv.x, xv.x, dot);
          dot = fmaf(uv.y, xv.y, dot);
          dot = fmaf(uv.z, xv.z, dot);
          dot = fmaf(uv.w, xv.w, dot);
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        if (part == 0) {
          sm.rs[row] = __ldcg(Ur + Wp) - dot;
          sm.sinv[row] = __ldcg(Ur + Wp + 1);
        }
        // the triangle, padded to 33 floats per row (the strip buffers are free now)
        float *tri = &sm.strip[0][0][0];
        const float4 t0 = __ldcg(reinterpret_cast<const float4 *>(Ur + 8 * part));
        const float4 t1 = __ldcg(reinterpret_cast<const float4 *>(Ur + 8 * part + 4));
        float *td = tri + row * 33 + 8 * part;
        td[0] = t0.x; td[1] = t0.y; td[2] = t0.z; td[3] = t0.w;
        td[4] = t1.x; td[5] = t1.y; td[6] = t1.z; td[7] = t1.w;
      }
      __syncthreads();
      if (warp == 0) {  // lane = row j of the triangle, its 32 entries in registers
        const float *tri = &sm.strip[0][0][0] + lane * 33;
        float ur[kNB];
#pragma unroll
        for (int i = 0; i < kNB; i++) ur[i] = tri[i];
        float r = sm.rs[lane];
        const float inv = sm.sinv[lane];
        float xmine = 0.f;
#pragma unroll
        for (int i = kNB - 1; i >= 0; i--) {
          const float xi = __shfl_sync(0xffffffffu, r * inv, i);
          if (lane == i) xmine = xi;
          if (lane < i) r = fmaf(-ur[i], xi, r);
        }
        sm.x[kNB * p + lane] = xmine;
      }
      __syncthreads();
    }
    bool finite = true;
    for (int k = tid; k < K; k += kLLThreads) finite = finite && isfinite(sm.x[k]);
    if (!finite) s_fail = 1;
    __syncthreads();
    if (s_fail) {  // :320-323
      if (tid == 0) atomicExch(&a.err_flags[kErrCholSolve], 1);
      continue;
    }
    for (int k = tid; k < kN; k += kLLThreads) {
      const float v = k < K ? sm.x[k] : 0.f;
      a.target[gu * kN + k] = v;
      for (int pi = 0; pi < a.n_peers; pi++) a.peers[pi][gu * kN + k] = v;
    }
  }
}

}  // namespace

// resident CTAs of cholesky_ll_kernel on the current device and its scratch size
static int ll_grid() {
  int dev = 0, sms = kNumSMsB200, per_sm = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cholesky_ll_kernel, kLLThreads, 0));
  return sms * std::max(per_sm, 1);
}
size_t cholesky_ll_scratch_bytes() { return (size_t)ll_grid() * kScratchFloats * sizeof(float); }

// Cholesky rows whose Gram blocks are in `workspace`: a.order / a.n_sched = the chunk's rows,
// first_job[0 .. n_sched] their absolute job ranges, scratch >= cholesky_ll_scratch_bytes().
void launch_solve_cholesky_from_gram(const SolveArgs &a, const int32_t *first_job, int job0, int job_cap,
                                     const float *workspace, float *scratch, cudaStream_t s) {
  if (a.n_sched <= 0) return;
  if (a.ld != kN || a.K > kN) throw NotImplemented("Cholesky from Gram blocks: the row stride must be 256");
  SolveArgs args = a;
  args.row_jobs = first_job;
  CUDA_CHECK(cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), s));
  const unsigned grid = (unsigned)std::min<int64_t>(a.n_sched, ll_grid());
  cholesky_ll_kernel<<<grid, kLLThreads, 0, s>>>(args, LLArgs{workspace, job0, job_cap, scratch});
  count_launch();
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace ials
