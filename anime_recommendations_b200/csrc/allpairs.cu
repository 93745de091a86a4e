// Half B, tensor-core part: many-query cosine candidates on the 5th-gen tensor cores.
//
// S = Qn * Cn^T for row-normalised bf16 tables (dim 128), fp32 accumulation in TMEM, fused per-row
// running top-k' selection in the epilogue -- the score matrix is never written to memory.
// Callers: all-pairs anime x anime / users x users (BASELINE cfg3; the reference's per-query
// np.dot + np.argsort of similar_anime.py:404-409 / similar_users.py:293-296 looped over all rows)
// and model_recs scoring (model_recs.py:394-396) with a watched mask.  The k' bf16-ranked candidates
// per row are re-ranked exactly in fp32 by ar_cosine_rerank (topk.cu).
//
// Kernel anatomy (one persistent CTA per SM, 320 threads, ~224 KB smem, all 512 TMEM columns):
//   warp 0      TMA producer: A = 2 x [128 query rows x 128 K] stays resident per work item;
//               B = [128 candidate rows x 128 K] tiles stream through a 3/4-stage mbarrier ring
//               (cp.async.bulk.tensor, SWIZZLE_128B, two 64-element K blocks per tile)
//   warp 1      MMA issuer (one elected lane): per B tile 2 x 8 tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=128, K=16) into a double-buffered TMEM accumulator, tcgen05.commit to the
//               smem-empty and tmem-full barriers
//   warps 2..9  epilogue: thread = one query row (TMEM lane); tcgen05.ld 32 columns at a time, a max tree
//               against the row's current k'-th best rejects almost every chunk; hits are inserted
//               into the row's private k'-entry list in shared memory (replace-min)
// A work item is (256-query tile, candidate chunk); items are striped over the CTAs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace ar {

constexpr int BM = 128;        // query rows per MMA tile (= TMEM lanes)
constexpr int BN = 128;        // candidates per B tile (= accumulator columns)
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int KBLK = 2;        // dim 128 = 2 swizzle blocks
constexpr int MT = 2;          // m-tiles per CTA: 256 query rows share every B tile
constexpr int QT = BM * MT;
constexpr int UMMA_K = 16;
constexpr int kApThreads = 320;
constexpr int kEpiWarps = 8;
constexpr uint32_t kSubTileBytes = BM * BK * 2;            // 16 KB: 128 rows x 128 B
constexpr uint32_t kABytes = MT * KBLK * kSubTileBytes;    // 64 KB
constexpr uint32_t kBStageBytes = KBLK * kSubTileBytes;    // 32 KB
constexpr uint32_t kTmemCols = 512;                        // 2 stages x 2 m-tiles x 128 columns

// bf16 x bf16 -> f32, M=128, N=128, both operands K-major (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct ApParams {
  int64_t q0, n_q;          // query rows [q0, q0+n_q) of the Q table
  int64_t c0, n_c;          // candidate rows [c0, c0+n_c) of the C table
  int n_qtiles, n_chunks, tiles_per_chunk, n_tiles;
  int exclude_self;
  const int32_t* self_ids;  // optional [n_q]: candidate id to drop per query row (overrides q0 + row)
  const uint32_t* watched;  // optional [n_q][watched_stride] bit rows over candidate ids; set bit = drop
  int64_t watched_stride;
  const float* thr_init;    // optional [n_q]: only candidates with score > thr_init[row] are listed
  int32_t* out_idx;         // [n_chunks][n_q][CAP]
  float* out_score;         // [n_chunks][n_q][CAP]
  int32_t* out_cnt;         // [n_chunks][n_q] entries kept
  float* out_thr;           // [n_chunks][n_q] every unlisted candidate of the chunk has score <= this
  float* dump;              // optional raw scores [n_q][n_c] (tests)
  int dbg;                  // perf attribution (AR_AP_DEBUG): 0 full, 1 no slow path, 2 TMEM loads only, 3 no TMEM loads
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 3-input maximum: one FMNMX3 (sm_100a) absorbs two new values per ALU slot
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Shared-memory plan: KP = entries a compaction keeps (16 or 24), list capacity 32 entries per query row.
constexpr int CAP = 32;
struct ApSmem {
  static constexpr int kStages = 3;
  static constexpr uint32_t kListBytes = CAP * QT * 8;
  static constexpr uint32_t kOffB = kABytes;
  static constexpr uint32_t kOffList = kOffB + kStages * kBStageBytes;
  static constexpr uint32_t kOffBar = kOffList + kListBytes;
  static constexpr uint32_t kNumBars = 2 * kStages + 2 + 8;
  static constexpr uint32_t kBytes = kOffBar + kNumBars * 8 + 16;
};

__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Row-private candidate list in shared memory (entry j of a row at byte offset j*QT*4 from the row's base:
// bank = thread, conflict-free).  Invariant: every admissible candidate seen so far with score > thr is in
// the list.  list_compact() pulls the held scores into registers, bisects (10 rolled steps of 32 compares)
// for the lowest cut with at most KP scores above it (within (hi-lo)/1024), raises thr to the cut and
// squeezes the list.  Scores tied at the cut are dropped together (list under-full: the re-rank cannot
// certify such a row and the caller's exact path takes it).
// Code size matters as much as instruction count here: an earlier version that inlined an unrolled scan at
// every use was 120 KB of code and stalled on instruction fetch (ncu: icc hit rate 60 %, `no_instruction`
// the top stall); out-of-line functions cost stack spills that go to L2 (this kernel leaves ~2 KB of L1).
template <int KP>
__device__ __forceinline__ void list_compact(uint32_t sa, uint32_t ia, int& cnt, float& thr) {
  float s[CAP];
  float hi = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    s[j] = (j < cnt) ? lds_f32(sa + j * QT * 4) : -CUDART_INF_F;
    hi = fmaxf(hi, s[j]);
  }
  // bisect for the LOWEST cut that leaves at most KP scores above it (a lower thr certifies more rows):
  // invariant count(s > lo) > KP (true for lo = thr: cnt > KP entries, all above thr) and count(s > hi) <= KP
  float lo = fmaxf(thr, -1.01f);
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    const float mid = 0.5f * (lo + hi);
    int c = 0;
#pragma unroll
    for (int j = 0; j < CAP; ++j) c += (s[j] > mid) ? 1 : 0;
    if (c > KP) lo = mid; else hi = mid;
  }
  const float nt = hi;  // <= KP entries survive, whatever the ties
  int w = 0;
#pragma unroll
  for (int j = 0; j < CAP; ++j) {
    if (s[j] > nt) {  // padding entries are -inf: never kept
      const int id = lds_s32(ia + j * QT * 4);
      sts_f32(sa + w * QT * 4, s[j]);
      sts_s32(ia + w * QT * 4, id);
      ++w;
    }
  }
  cnt = w;
  thr = nt;
}

template <int KP, bool DUMP>
__global__ void __launch_bounds__(kApThreads, 1)
allpairs_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, ApParams p) {
  using L = ApSmem;
  constexpr int S = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment; the launch reserves 1 KB of slack for this
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* list_s = reinterpret_cast<float*>(smem + L::kOffList);            // [CAP][QT]
  int* list_i = reinterpret_cast<int*>(smem + L::kOffList + CAP * QT * 4);  // [CAP][QT]
  const uint32_t bar0 = sbase + L::kOffBar;
  auto full_b = [&](int s) { return bar0 + 8u * s; };
  auto empty_b = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t a_full = bar0 + 8u * (2 * S), a_empty = bar0 + 8u * (2 * S + 1);
  // four accumulator buffers of 128 TMEM columns, buffer = 2*(tile parity) + m-tile, each with its own
  // full/empty barrier: a slow epilogue warp only holds back the MMAs of its own m-tile
  auto tmem_full = [&](int b) { return bar0 + 8u * (2 * S + 2 + b); };
  auto tmem_empty = [&](int b) { return bar0 + 8u * (2 * S + 6 + b); };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + L::kOffBar + L::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int b = 0; b < 2 * MT; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), kEpiWarps / MT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int n_items = p.n_qtiles * p.n_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
      uint32_t bj = 0, n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int qt = item / p.n_chunks, ch = item - qt * p.n_chunks;
        mbar_wait(a_empty, (n & 1) ^ 1);  // previous item's MMAs are done with A
        mbar_expect_tx(a_full, kABytes);
        const int qrow = (int)(p.q0 + (int64_t)qt * QT);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int kb = 0; kb < KBLK; ++kb)
            tma_load_2d(sbase + (mt * KBLK + kb) * kSubTileBytes, &tmQ, kb * BK, qrow + mt * BM, a_full);
        const int t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++bj) {
          const int s = bj % S;
          mbar_wait(empty_b(s), ((bj / S) & 1) ^ 1);
          mbar_expect_tx(full_b(s), kBStageBytes);
          const int crow = (int)(p.c0 + (int64_t)t * BN);
#pragma unroll
          for (int kb = 0; kb < KBLK; ++kb)
            tma_load_2d(sbase + L::kOffB + s * kBStageBytes + kb * kSubTileBytes, &tmC, kb * BK, crow, full_b(s));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t bj = 0, it = 0, n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int qt = item / p.n_chunks, ch = item - qt * p.n_chunks;
        (void)qt;
        mbar_wait(a_full, n & 1);
        const int t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
        for (int t = t0; t < t1; ++t, ++bj, ++it) {
          const int s = bj % S, as = it & 1;
          mbar_wait(full_b(s), (bj / S) & 1);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mbar_wait(tmem_empty(as * MT + mt), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)((as * MT + mt) * BN);
#pragma unroll
            for (int k = 0; k < KBLK * (BK / UMMA_K); ++k) {
              const int kb = k / (BK / UMMA_K), kk = k % (BK / UMMA_K);
              const uint64_t ad = smem_desc(sbase + (mt * KBLK + kb) * kSubTileBytes + kk * UMMA_K * 2);
              const uint64_t bd = smem_desc(sbase + L::kOffB + s * kBStageBytes + kb * kSubTileBytes + kk * UMMA_K * 2);
              tc_mma(d, ad, bd, kInstrDesc, k > 0 ? 1u : 0u);
            }
            tc_commit(tmem_full(as * MT + mt));  // this m-tile's accumulators are ready for its epilogue warps
          }
          tc_commit(empty_b(s));                 // B stage reusable once these MMAs retire
        }
        tc_commit(a_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = query row
    const int quarter = warp & 3;               // TMEM lane quarter this warp may touch
    const int mt = (warp - 2) >> 2;
    const int row_in_cta = mt * BM + quarter * 32 + lane;
    const uint32_t my_s = smem_u32(list_s + row_in_cta);  // stride QT between entries: bank = thread
    const uint32_t my_i = smem_u32(list_i + row_in_cta);
    const int dbg = p.dbg;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qt = item / p.n_chunks, ch = item - qt * p.n_chunks;
      const int64_t qlocal = (int64_t)qt * QT + row_in_cta;  // row within [0, n_q)
      const bool row_ok = qlocal < p.n_q;
      int self_id = -1;
      if (p.exclude_self && row_ok) self_id = p.self_ids ? p.self_ids[qlocal] : (int)(p.q0 + qlocal);
      const uint32_t* wrow = (p.watched && row_ok) ? p.watched + qlocal * p.watched_stride : nullptr;
      float thr = CUDART_INF_F;  // rows past n_q list nothing
      if (row_ok) thr = p.thr_init ? p.thr_init[qlocal] : -CUDART_INF_F;
      if (dbg == 1) thr = CUDART_INF_F;
      int cnt = 0;
      const int t0 = ch * p.tiles_per_chunk, t1 = min(p.n_tiles, t0 + p.tiles_per_chunk);
      const int c_end = (int)(p.c0 + p.n_c);

      // One 32-column chunk held in registers.  Fast path: FMNMX3 tree (18 ALU ops) against the row's
      // threshold and one vote.  Slow path (some row of the warp beat its threshold; warp-uniform entry):
      // vote per group of eight, then per hot group select its eight scores (no dynamic register indexing),
      // make room if a row needs it -- the WHOLE warp compacts, same instruction stream, so the other rows'
      // squeezes are free and their thresholds tighten early -- and append the admissible hits.
      auto consume = [&](const uint32_t (&v)[32], int cb) {
        if (DUMP && row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cid = cb + i;
            if (cid < c_end) p.dump[qlocal * p.n_c + (cid - p.c0)] = __uint_as_float(v[i]);
          }
        }
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float r1 = fmax3(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]), __uint_as_float(v[8 * j + 2]));
          const float r2 = fmax3(__uint_as_float(v[8 * j + 3]), __uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
          g[j] = fmax3(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]), fmaxf(r1, r2));
        }
        const float m = fmaxf(fmax3(g[0], g[1], g[2]), g[3]);
        if (!__any_sync(0xffffffffu, m > thr)) return;
        // (measured alternatives, 350k x 350k: four static-index copies of the scan instead of the select chain
        // -- 70 KB of code, 91 ms; static copies with compaction hoisted to one site per chunk pair and a
        // drop-and-remember overflow rule -- 56 KB, 46.6 ms; fully predicated appends -- 44 ms; this version
        // -- 35.8 ms.  Every variant that grew the code lost more to instruction fetch than it saved.  Also tried:
        // moving ALL of this to dedicated scanner warps fed through shared-memory descriptor rings by
        // never-diverging filter warps (scanner re-reads the chunk from TMEM): 62 ms with 4 scanners, 40 ms
        // with 8, 49 ms with descriptors pushed per chunk -- the rescans hold the accumulator buffer longer
        // than the two-tile slack allows.)
        uint32_t gmask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) gmask |= __any_sync(0xffffffffu, g[j] > thr) ? (1u << j) : 0u;
#pragma unroll 1
        while (gmask) {
          const int gq = __ffs(gmask) - 1;
          gmask &= gmask - 1;
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t lo = (gq & 1) ? v[8 + e] : v[e], hi = (gq & 1) ? v[24 + e] : v[16 + e];
            x[e] = __uint_as_float((gq & 2) ? hi : lo);
          }
          const float glo = (gq & 1) ? g[1] : g[0], ghi = (gq & 1) ? g[3] : g[2];
          const bool hot = ((gq & 2) ? ghi : glo) > thr;
          if (__any_sync(0xffffffffu, hot && cnt > CAP - 8)) {
            if (cnt > KP) list_compact<KP>(my_s, my_i, cnt, thr);
          }
          // (a straight-line, fully predicated version of this append loop was measured 23 % SLOWER -- 44.1 vs
          // 35.9 ms on 350k x 350k -- the skipped instructions matter more than the reconvergence stalls)
          if (hot) {
            const int cid0 = cb + 8 * gq;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (x[e] > thr) {
                const int cid = cid0 + e;
                bool ok = (cid < c_end) && (cid != self_id);
                if (ok && wrow) ok = !((wrow[cid >> 5] >> (cid & 31)) & 1u);
                if (ok) {
                  sts_f32(my_s + cnt * QT * 4, x[e]);
                  sts_s32(my_i + cnt * QT * 4, cid);
                  ++cnt;
                }
              }
            }
          }
        }
      };

      for (int t = t0; t < t1; ++t, ++it) {
        const int as = it & 1;
        mbar_wait(tmem_full(as * MT + mt), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((as * MT + mt) * BN);
        const int cbase = (int)(p.c0 + (int64_t)t * BN);
        if (dbg >= 3) {  // attribution: MMA + TMA only
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(as * MT + mt));
          continue;
        }
        // software pipeline over the four 32-column chunks (two per trip of a rolled loop, so that the slow
        // path exists twice, not four times): the next tcgen05.ld is in flight while the current chunk is scanned
        uint32_t va[32], vb[32];
        __syncwarp();
        tmem_ld32(tbase, va);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          tmem_ld_wait();
          __syncwarp();
          tmem_ld32(tbase + 64 * h + 32, vb);
          if (dbg < 2) consume(va, cbase + 64 * h);
          tmem_ld_wait();
          if (h == 0) {
            __syncwarp();
            tmem_ld32(tbase + 64, va);
          } else {
            // every column of this accumulator buffer is now in registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(as * MT + mt));
          }
          if (dbg < 2) consume(vb, cbase + 64 * h + 32);
        }
      }
      if (row_ok) {
        const int64_t r = (int64_t)ch * p.n_q + qlocal;
        const int64_t o = r * CAP;
#pragma unroll 4
        for (int j = 0; j < CAP; ++j) {
          const bool in = j < cnt;
          p.out_idx[o + j] = in ? lds_s32(my_i + j * QT * 4) : -1;
          p.out_score[o + j] = in ? lds_f32(my_s + j * QT * 4) : -CUDART_INF_F;
        }
        p.out_cnt[r] = cnt;
        p.out_thr[r] = thr;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static int make_map(CUtensorMap* tm, const void* base, int64_t rows, int dim) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return AR_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for %lld x %d bf16", (int)r, (long long)rows, dim);
    return AR_ERR_CUDA;
  }
  return AR_OK;
}

static int ap_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int KP, bool DUMP>
static int launch_allpairs(const CUtensorMap& tq, const CUtensorMap& tc, const ApParams& p, cudaStream_t st) {
  using L = ApSmem;
  static bool attr = false;
  if (!attr) {
    AR_CUDA(cudaFuncSetAttribute(allpairs_topk_kernel<KP, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes + 1024));
    attr = true;
  }
  const int items = p.n_qtiles * p.n_chunks;
  const int grid = std::max(1, std::min(items, ap_sm_count()));
  allpairs_topk_kernel<KP, DUMP><<<grid, kApThreads, L::kBytes + 1024, st>>>(tq, tc, p);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

}  // namespace ar

using namespace ar;

// Candidate chunks that keep every SM busy: one chunk when there are enough 256-query tiles, otherwise split
// the candidate range so that items ~ 2 x SMs (each chunk keeps its own list; the re-rank sees them all).
extern "C" int32_t ar_allpairs_chunks(int64_t n_q, int64_t n_c) {
  if (n_q <= 0 || n_c <= 0) return 1;
  const int64_t qtiles = (n_q + QT - 1) / QT, tiles = (n_c + BN - 1) / BN;
  const int sms = 148;
  if (qtiles >= 2 * sms) return 1;
  int64_t want = (2 * sms + qtiles - 1) / qtiles;
  want = std::min<int64_t>(want, std::max<int64_t>(1, tiles / 8));  // at least 8 B tiles per chunk
  want = std::max<int64_t>(1, std::min<int64_t>(want, 64));
  const int64_t per = (tiles + want - 1) / want;
  return (int32_t)((tiles + per - 1) / per);                        // the count the kernel will actually use
}

extern "C" int32_t ar_allpairs_list_cap(int32_t kprime) {
  return (kprime == 16 || kprime == 24) ? CAP : 0;
}

extern "C" int ar_cosine_topk_allpairs(const void* Qn_bf16, int64_t q_rows_total, int64_t q0, int64_t n_q,
                                       const void* Cn_bf16, int64_t c_rows_total, int64_t c0, int64_t n_c,
                                       int32_t dim, int32_t kprime, int32_t exclude_self, const int32_t* self_ids,
                                       const uint32_t* watched, int64_t watched_stride, const float* thr_init,
                                       int32_t n_chunks, int32_t* out_idx, float* out_score, int32_t* out_cnt,
                                       float* out_thr, float* dump_scores, void* stream) {
  const char* who = "ar_cosine_topk_allpairs";
  AR_REQUIRE(out_idx && out_score && out_cnt && out_thr, "%s: null output", who);
  AR_REQUIRE(kprime == 16 || kprime == 24, "%s: kprime must be 16 or 24 (got %d)", who, kprime);
  AR_REQUIRE(Qn_bf16 && Cn_bf16, "%s: null table", who);
  if (dim != KBLK * BK) {
    set_error("%s: the tcgen05 path is built for dim %d (got %d)", who, KBLK * BK, dim);
    return AR_ERR_UNSUPPORTED;
  }
  AR_REQUIRE(q0 >= 0 && n_q >= 0 && q0 + n_q <= q_rows_total, "%s: query range outside the table", who);
  AR_REQUIRE(c0 >= 0 && n_c >= 0 && c0 + n_c <= c_rows_total, "%s: candidate range outside the table", who);
  AR_REQUIRE(q_rows_total < (1ll << 31) && c_rows_total < (1ll << 31), "%s: table too large", who);
  AR_REQUIRE(n_chunks >= 1, "%s: bad n_chunks", who);
  AR_REQUIRE(((uintptr_t)Qn_bf16 & 15) == 0 && ((uintptr_t)Cn_bf16 & 15) == 0, "%s: tables must be 16-byte aligned", who);
  if (n_q == 0 || n_c == 0) return AR_OK;
  ApParams p{};
  p.q0 = q0; p.n_q = n_q; p.c0 = c0; p.n_c = n_c;
  p.n_qtiles = (int)((n_q + QT - 1) / QT);
  p.n_tiles = (int)((n_c + BN - 1) / BN);
  const int want = std::min<int>(n_chunks, p.n_tiles);
  p.tiles_per_chunk = (p.n_tiles + want - 1) / want;
  p.n_chunks = (p.n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  AR_REQUIRE(p.n_chunks == n_chunks, "%s: n_chunks %d does not tile %d candidate tiles; use %d", who, n_chunks,
             p.n_tiles, p.n_chunks);
  p.exclude_self = exclude_self;
  p.self_ids = self_ids;
  p.watched = watched;
  p.watched_stride = watched_stride;
  p.thr_init = thr_init;
  p.out_idx = out_idx;
  p.out_score = out_score;
  p.out_cnt = out_cnt;
  p.out_thr = out_thr;
  p.dump = dump_scores;
  static const int dbg = [] { const char* e = getenv("AR_AP_DEBUG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  CUtensorMap tq, tc;
  int rc;
  if ((rc = make_map(&tq, Qn_bf16, q_rows_total, dim))) return rc;
  if ((rc = make_map(&tc, Cn_bf16, c_rows_total, dim))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dump_scores) return kprime == 16 ? launch_allpairs<16, true>(tq, tc, p, st) : launch_allpairs<24, true>(tq, tc, p, st);
  return kprime == 16 ? launch_allpairs<16, false>(tq, tc, p, st) : launch_allpairs<24, false>(tq, tc, p, st);
}
