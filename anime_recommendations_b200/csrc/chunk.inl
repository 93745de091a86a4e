// Single-GPU training: ONE persistent kernel per chunk of steps (included by train.cu).
//
// One CTA per SM.  Warps 0..AR_STEP_WARPS-1 of every CTA are STEP warps: they walk the chunk step by step,
//     F        warp per sample (a contiguous share per warp, two samples in flight): wait until both rows of the
//              sample are at optimizer step t-1 (AR_ADAM_REPLAY: last_step[row] is the readiness flag -- the row
//              was either updated by the previous step or brought there by a replay item), gather them with
//              128-bit loads, l2-normalise, dot; normalised rows, c, 1/||u||, 1/||a|| to the step scratch
//              (double-buffered by step parity); per-CTA (sum c, sum c^2) in a fixed order
//     -- grid barrier --   the only one per step: BatchNorm needs the statistics of the whole batch
//     head     every CTA, redundantly and bit-identically: batch statistics from the per-CTA partials, one pass
//              over the step's (c, label) for the five backward sums, Adam on the four head scalars (kept in
//              shared memory for the whole chunk), the step scalars; each CTA adds the reported BCE / MSE of ITS
//              share of the samples
//     U        warp per distinct row (contiguous share): atomic-free segment reduction over the plan's sorted
//              samples + L2 term + Adam, one RMW of (W, m, v), then (behind a fence) last_step[row] = t; rows hit by
//              more than AR_HEAVY_LEN samples are cut into pieces, the last piece to arrive (ticket) adds them
//              in piece order
// (AR_ADAM_TOUCHED and AR_ADAM_DENSE have no per-row flags: a second grid barrier follows U; DENSE adds a pass over
// every other row and a third barrier.)  The other warps are REPLAY warps (AR_ADAM_REPLAY): they pull items of the
// plan-time schedule (ar_plan_sched: "replay the missed pure-L2 Adam steps of row r up to step t-1"),
// earliest deadline first, up to `depth` steps ahead of the step warps, and publish last_step[row] behind a fence;
// the replay is bound by the special-function pipe (sqrt + reciprocal per element-step) and fills it while the
// step warps are bound by memory latency.  Every wait has a time-out that raises ctl->abort instead of hanging the
// GPU.  Registers move from the replay warps to the step warps at kernel entry (setmaxnreg).
//
// Arithmetic follows oracle/train.py; all sums have a fixed order => bit-reproducible.

namespace ar {

// Build-time knobs (A/B builds: -DAR_STEP_WARPS=... etc.; the defaults are the measured best)
#ifndef AR_STEP_WARPS
#define AR_STEP_WARPS 16          // step warps per CTA (a multiple of 4: setmaxnreg works on warpgroups)
#endif
#ifndef AR_CHUNK_THREADS
#define AR_CHUNK_THREADS 768      // NV == 1: step warps + replay warps
#endif
#ifndef AR_REGS_LAUNCH
#define AR_REGS_LAUNCH 64         // NV == 1: registers per thread at launch
#endif
#ifndef AR_REGS_REPLAY
#define AR_REGS_REPLAY 40         // ... the replay warps keep
#endif
#ifndef AR_REGS_STEP
#define AR_REGS_STEP 72           // ... the step warps grow to
#endif
constexpr int kStepWarps = AR_STEP_WARPS;
constexpr int kStepThreads = kStepWarps * 32;
static_assert(kStepWarps % 4 == 0 && kStepWarps <= 32, "step warps come in warpgroups");
constexpr int kMaxCtas = 160;      // B200: 148 SMs
constexpr int kStamps = 8;
constexpr unsigned kCodeRowMask = (1u << 26) - 1u;
constexpr unsigned kCodeSplit = 1u << 30;
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;

struct ChunkCtl {            // first 256 bytes of the workspace; zeroed before every launch
  unsigned int arrive;       // grid-barrier arrivals since launch (monotone)
  unsigned int gen;          // completed grid barriers
  int abort;                 // a wait timed out: every loop exits
  int pad0;
  unsigned long long stats[8];
  unsigned long long pad1[22];
};
static_assert(sizeof(ChunkCtl) == 256, "ChunkCtl is 256 bytes");

struct ChunkLayout {
  size_t ctl, htick, fpart, hsum, mpart, stamps, hpart, stash1, total, zero_bytes;
  int n_pieces;              // capacity of hpart / htick per table
};
static ChunkLayout chunk_layout(int n_slots, int batch_cap, int dim) {
  ChunkLayout l;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  l.n_pieces = 2 * (batch_cap / AR_HEAVY_LEN) + 2;
  l.ctl = 0;
  l.htick = 256;
  l.zero_bytes = up(l.htick + (size_t)2 * l.n_pieces * sizeof(int32_t));
  l.fpart = l.zero_bytes;
  l.hsum = up(l.fpart + (size_t)kMaxCtas * 2 * sizeof(double));
  l.mpart = up(l.hsum + (size_t)kMaxCtas * 8 * sizeof(double));
  l.stamps = up(l.mpart + (size_t)n_slots * kMaxCtas * 2 * sizeof(double));
  l.hpart = up(l.stamps + (size_t)n_slots * kStamps * sizeof(long long));
  l.stash1 = up(l.hpart + (size_t)2 * l.n_pieces * (dim + 4) * sizeof(float));
  l.total = up(l.stash1 + (size_t)batch_cap * (2 * dim + 4) * sizeof(float));   // odd steps' uh, ah, c, ru, ra
  return l;
}

struct ChunkArgs {
  ar_table tab[2];
  ar_plan plan[2];
  ar_sched sched;
  const int32_t* iu;         // first sample of the chunk
  const int32_t* ia;
  const float* label;
  int64_t t0;
  int n_steps, batch, mode, depth;
  float l2x2;
  const float* alpha;
  float* head;
  float* head_m;
  float* head_v;
  float* bn_moving;
  float* uh[2];              // step scratch, by step parity
  float* ah[2];
  float* c[2];
  float* ru[2];
  float* ra[2];
  double* fwd_part;
  double* hsum;              // [cta][8] per-CTA partial sums of the head's second pass
  float* metrics;
  RegAcc reg;
  int32_t* health;
  ChunkCtl* ctl;
  int32_t* htick;
  double* mpart;
  long long* stamps;
  float* hpart;
  int n_pieces;
};

// Multi-GPU (peer-memory) extension of the same kernel, csrc/peer.inl: the tables are row-sharded, this rank works on
// the samples of the GLOBAL batch that touch its rows (two selection lists per step) and reads the other table's row
// -- and its readiness flag -- straight from the owner's HBM over NVLink.
constexpr int kPeerRanks = AR_PEER_MAX_RANKS;
constexpr int kPeerErrWordC = 32;         // flags[32]: sticky "a cross-rank wait timed out" (shared with the staged path)
constexpr int kPeerLenWord = 56;          // flags[56 + parity]: length of the list this rank sent two steps ago (local)
constexpr unsigned long long kPeerWaitTimeoutNs = 20000000000ull;
// Tagged words: one 64-bit store carries the payload and the step it belongs to, so the reader needs neither a flag
// nor a fence -- it polls the word itself.  Pair word: low half = position in the global batch (18 bits) | tag << 18,
// high half = the cosine's bits; the tag tells step t from step t-2, the only other step whose word can still sit in
// the slot (a slot that drops out of the list is overwritten with kPairInvalid).  Header word: low half = 32 bits of
// payload, high half = the optimizer step.
constexpr int kPairPosBits = 18;
constexpr unsigned kPairPosMask = (1u << kPairPosBits) - 1u;
constexpr unsigned long long kPairInvalid = ~0ull;
static_assert(AR_PEER_MAX_RANKS * AR_MAX_BATCH < (int)kPairPosMask, "positions of the global batch must fit the pair word");
constexpr int kHdrWords = 8;              // 5 used: list length, sum c (lo, hi), sum c^2 (lo, hi)
__device__ __forceinline__ unsigned pair_tag(int64_t t) { return (unsigned)((t >> 1) & ((1u << (32 - kPairPosBits)) - 1u)); }
struct PeerExt {
  const float* W_peer[2][kPeerRanks];          // [table][rank] shard bases (own entry = local pointer)
  int32_t* rowflag_peer[2][kPeerRanks];        // [table][rank] per-row "at step" words the owners publish for the peers
  unsigned long long* pairs_peer[kPeerRanks];  // [rank] its inbox of pair words: [2 parities][G senders][cap]
  unsigned long long* hdrin_peer[kPeerRanks];  // [rank] its inbox of header words: [2 parities][G senders][kHdrWords]
  int32_t* flags;                              // my flag words (error word, kPeerLenWord)
  const int32_t* key[2];                       // selection lists of the chunk, [slot][cap]: local row of my table
  const int32_t* oth[2];                       // ... global row of the other table
  const int32_t* samp0;                        // ... position in the global batch (user side)
  const int32_t* cnt[2];                       // [slot]
  const float* lab[2];                         // [slot][cap] label of the listed sample
  const float* label_step;                     // [slot][G*B] labels in global-batch order
  float* cl1[2];                               // [parity][cap] cosines of the anime-side list (the user side's, the
                                               // normalised other rows and 1/||my row|| live in ChunkArgs' c, ah/uh, ru/ra)
  int G, me, cap, gb;                          // gb = G * per-rank batch
  int strict;                                  // 1: system-scope fences in front of the row words (see below)
};
__device__ __forceinline__ int ld_sys_s32(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_sys_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys_s32(int32_t* p, int v) { asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// Row flags in peer mode: last_step[row] is the LOCAL flag; the peers poll rowflag[row] (own array, so that the two
// can be published differently) and then read the row from the owner's HBM with loads that bypass their own caches.
//   default   both words are written behind the same DEVICE-scope fence.  The row lives in the owner's memory: once
//             the fence has completed, the row's stores have been performed at the owner's L2, which is where every
//             reader of that memory -- local SM or NVLink peer -- is served from, so a peer that sees the word reads
//             the new row.  This is an argument about the hardware; the PTX model asks for a system-scope fence here.
//   strict    (AR_PEER_STRICT=1) the PTX-conformant version: rowflag[row] is written behind a SYSTEM-scope fence -- one
//             per step warp and step (after its last row), one per batch of replay items.  MEMBAR.SYS stalls far more
//             than the warp that issues it: measured 93.7 us/step against 57.8 (1 rank, cfg2 shapes).
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void step_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kStepThreads) : "memory"); }

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
template <int NV>
__device__ __forceinline__ void tile_load_cg(RowTile<NV>& t, const float* row, int d4, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int j = lane + 32 * k;
    t.x[k] = (j < d4) ? ldcg4(row + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NV>
__device__ __forceinline__ float tile_partial_dot(const RowTile<NV>& a, const RowTile<NV>& b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += dot4(a.x[k], b.x[k]);
  return s;
}

// Thread 0 of the step warps spins until `*p >= want`; false = timed out or aborted elsewhere.
__device__ __forceinline__ bool spin_until_ge(const int* p, int want, ChunkCtl* ctl) {
  unsigned spins = 0;
  unsigned long long t_begin = 0;
  while (ld_relaxed_s32(p) < want) {
    __nanosleep(32);
    if ((++spins & 255u) == 0u) {
      if (ld_relaxed_s32(&ctl->abort)) return false;
      const unsigned long long now = globaltimer_ns();
      if (!t_begin) t_begin = now;
      else if (now - t_begin > kWaitTimeoutNs) return false;
    }
  }
  return ld_acquire_s32(p) >= want;
}

// Grid barrier over the step warps of all CTAs.  Arrivals are counted monotonically (no reset race): barrier b is
// complete when arrive == n_ctas * b.  One release-reduction per CTA, everyone polls the counter itself (no
// "last arriver publishes a generation" hop).  Returns false once the chunk is aborted.
__device__ __forceinline__ bool grid_bar(ChunkCtl* ctl, unsigned& bar, int n_ctas, volatile int* abort_s) {
  step_bar();
  ++bar;
  if (threadIdx.x == 0) {
    // release: this CTA's writes (ordered before by the CTA barrier above) are visible before the arrival
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ctl->arrive) : "memory");
    if (!spin_until_ge(reinterpret_cast<const int*>(&ctl->arrive), (int)((unsigned)n_ctas * bar), ctl)) {
      *abort_s = 1;
      atomicExch(&ctl->abort, 1);
    }
  }
  step_bar();
  return *abort_s == 0;
}

// sum over the step warps of NVAL doubles per thread; every thread gets the totals (fixed order)
template <int NVAL>
__device__ __forceinline__ void step_block_sum(double (&v)[NVAL], double* smem /* [NVAL][kStepWarps] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NVAL; ++i) v[i] = warp_sum(v[i]);
  step_bar();  // protect smem reuse
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NVAL; ++i) smem[i * kStepWarps + wid] = v[i];
  }
  step_bar();
#pragma unroll
  for (int i = 0; i < NVAL; ++i) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kStepWarps; ++w) t += smem[i * kStepWarps + w];
    v[i] = t;
  }
}

// Shared state of a CTA.  The step warps pass every grid barrier, so they know how far the chunk is; the other
// warps read it here instead of polling global memory (2368 warps polling one L2 line starve everything else).
constexpr int kWin = 1024;   // alpha / step-weight window kept in shared memory: the chunk's steps and the ~750 before
struct StepSmem {
  int steps_done;            // steps whose row update is complete grid-wide (written by thread 0 after the barrier)
  int dense_go;              // AR_ADAM_DENSE: dense pass of step dense_go-1 may start
  int dense_arrive;          // AR_ADAM_DENSE: helper warps that finished their share (monotone over the chunk)
  int quit;                  // the step warps have left the loop (abort)
  int abort;
  long long win_base;        // alpha_w[i] = alpha[win_base + i]
  float head[4], hm[4], hv[4], bn[2];
  float stepc[K_STEPC + 2];
  double red[8 * kStepWarps];
  float alpha_w[kWin];
  float stepw_w[kWin];
  // peer mode only (kept behind everything the single-GPU kernel uses)
  unsigned long long pmbar[kStepWarps][4];   // one mbarrier per step warp and pair buffer (bulk copies)
  double ps0[kPeerRanks], ps1[kPeerRanks];   // every rank's batch sums / published list length of the current step
  int pcnt[kPeerRanks];
};
__device__ __forceinline__ void dense_arrive(StepSmem& sm) { atomicAdd(&sm.dense_arrive, 1); }
struct StepTabs {            // alpha[t], stepw[t] through the shared-memory window
  const float* alpha;
  const float* stepw;        // may be null (weight 1)
  const StepSmem* sm;
  __device__ __forceinline__ float a_at(int64_t t) const {
    const long long i = t - sm->win_base;
    return (unsigned long long)i < (unsigned long long)kWin ? sm->alpha_w[i] : __ldg(alpha + t);
  }
  __device__ __forceinline__ float w_at(int64_t t) const {
    if (!stepw) return 1.f;
    const long long i = t - sm->win_base;
    return (unsigned long long)i < (unsigned long long)kWin ? sm->stepw_w[i] : __ldg(stepw + t);
  }
};

// Replay pure-L2 Adam steps (from, to] of a full row held as float4 tiles; regd += the replayed
// steps' share of the regulariser term.  The common all-weights-one blocks take a shorter path.
template <int NV>
__device__ __forceinline__ void replay_tile(RowTile<NV>& w, RowTile<NV>& m, RowTile<NV>& v, const StepTabs& tb,
                                            int64_t from, int64_t to, float l2x2, int lane, double& regd) {
#pragma unroll 1
  for (int64_t t0 = from + 1; t0 <= to; t0 += 32) {
    const int64_t tl = t0 + lane;
    const float a_l = (tl <= to) ? tb.a_at(tl) : 0.f;
    const float w_l = (tl <= to) ? tb.w_at(tl) : 1.f;
    const int cnt = (int)min((int64_t)32, to - t0 + 1);
    float accf = 0.f;
    if (__all_sync(0xffffffffu, w_l == 1.f)) {
#pragma unroll 2
      for (int s = 0; s < cnt; ++s) {
        const float a = __shfl_sync(0xffffffffu, a_l, s);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          accf = fmaf(w.x[k].x, w.x[k].x, accf);
          accf = fmaf(w.x[k].y, w.x[k].y, accf);
          accf = fmaf(w.x[k].z, w.x[k].z, accf);
          accf = fmaf(w.x[k].w, w.x[k].w, accf);
          const float4 g = make_float4(__fmul_rn(l2x2, w.x[k].x), __fmul_rn(l2x2, w.x[k].y),
                                       __fmul_rn(l2x2, w.x[k].z), __fmul_rn(l2x2, w.x[k].w));
          adam4(w.x[k], m.x[k], v.x[k], g, a);
        }
      }
    } else {
#pragma unroll 1
      for (int s = 0; s < cnt; ++s) {
        const float a = __shfl_sync(0xffffffffu, a_l, s);
        const float sw = __shfl_sync(0xffffffffu, w_l, s);
        float p = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          p = fmaf(w.x[k].x, w.x[k].x, p);
          p = fmaf(w.x[k].y, w.x[k].y, p);
          p = fmaf(w.x[k].z, w.x[k].z, p);
          p = fmaf(w.x[k].w, w.x[k].w, p);
          const float4 g = make_float4(__fmul_rn(l2x2, w.x[k].x), __fmul_rn(l2x2, w.x[k].y),
                                       __fmul_rn(l2x2, w.x[k].z), __fmul_rn(l2x2, w.x[k].w));
          adam4(w.x[k], m.x[k], v.x[k], g, a);
        }
        accf = fmaf(sw, p, accf);
      }
    }
    regd += (double)accf;
  }
}
// ... of one element per lane (the 32-element parts of a split row)
__device__ __forceinline__ void replay_lane(float& w, float& m, float& v, const StepTabs& tb, int64_t from, int64_t to,
                                            float l2x2, int lane, double& regd) {
#pragma unroll 1
  for (int64_t t0 = from + 1; t0 <= to; t0 += 32) {
    const int64_t tl = t0 + lane;
    const float a_l = (tl <= to) ? tb.a_at(tl) : 0.f;
    const float w_l = (tl <= to) ? tb.w_at(tl) : 1.f;
    const int cnt = (int)min((int64_t)32, to - t0 + 1);
    float accf = 0.f;
    if (__all_sync(0xffffffffu, w_l == 1.f)) {
#pragma unroll 4
      for (int s = 0; s < cnt; ++s) {
        const float a = __shfl_sync(0xffffffffu, a_l, s);
        accf = fmaf(w, w, accf);
        adam1(w, m, v, __fmul_rn(l2x2, w), a);
      }
    } else {
#pragma unroll 1
      for (int s = 0; s < cnt; ++s) {
        const float a = __shfl_sync(0xffffffffu, a_l, s);
        const float sw = __shfl_sync(0xffffffffu, w_l, s);
        accf = fmaf(sw * w, w, accf);
        adam1(w, m, v, __fmul_rn(l2x2, w), a);
      }
    }
    regd += (double)accf;
  }
}

__device__ __forceinline__ void reg_fix_add(double regd, const RegAcc& reg, unsigned long long& regfix) {
  // one conversion per row visit; integer adds are associative, so the total is independent of which warp does what
  regfix += (unsigned long long)__double2ll_rn(regd * (double)reg.scale);
}

// One schedule item: bring the row (or 32 of its elements) from step t_to-(g-1) to t_to.  The item's (W, m, v) are
// staged in shared memory with cp.async (no registers while in flight, L2-coherent .cg path): the NEXT item of the
// batch is on its way while this one is replayed.  The caller publishes last_step[row] behind a fence.
constexpr int kPartShift = 27;   // split rows: the parts count their arrivals in the top bits of last_step[row]
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Rows travel global -> shared with cp.async (no registers while in flight) and are read back by the lane that
// copied them, so a wait_group is all the synchronisation a TILE needs.  Two buffers per warp while shared memory
// allows it (the next row / sample pair is in flight during the current one's arithmetic).
template <int NV> struct StageCfg { static constexpr int kBufs = NV <= 2 ? 2 : 1; };
template <int NV>
__device__ __forceinline__ void stage_tile(float4* dst, const float* row, int d4, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int j = lane + 32 * k;
    if (j < d4) cp_async16(dst + j, row + 4 * j);
  }
}
template <int NV>
__device__ __forceinline__ void tile_from_stage(RowTile<NV>& t, const float4* src, int d4, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int j = lane + 32 * k;
    t.x[k] = (j < d4) ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// stage: [3][32*NV] float4 (W, m, v of one row)
template <int NV>
__device__ __forceinline__ void item_prefetch(const ChunkArgs& a, int code, float4* stage, int lane) {
  const bool second = code < 0;
  const int row = (int)((unsigned)code & kCodeRowMask);
  const int dim = a.tab[0].dim;
  const float* W = second ? a.tab[1].W : a.tab[0].W;
  const float* M = second ? a.tab[1].m : a.tab[0].m;
  const float* V = second ? a.tab[1].v : a.tab[0].v;
  if ((unsigned)code & kCodeSplit) {
    const int e0 = (int)(((unsigned)code >> 26) & 15u) * 32;
    const size_t o = (size_t)row * dim + e0 + 4 * lane;
    if (lane < 8 && e0 + 4 * lane < dim) {
      cp_async16(stage + lane, W + o);
      cp_async16(stage + 32 * NV + lane, M + o);
      cp_async16(stage + 64 * NV + lane, V + o);
    }
  } else {
    const int d4 = dim >> 2;
    const size_t o = (size_t)row * dim;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      if (j < d4) {
        cp_async16(stage + j, W + o + 4 * j);
        cp_async16(stage + 32 * NV + j, M + o + 4 * j);
        cp_async16(stage + 64 * NV + j, V + o + 4 * j);
      }
    }
  }
  cp_async_commit();
}

template <int NV>
__device__ __forceinline__ void replay_item(const ChunkArgs& a, const StepTabs& tabs, int code, int g, int64_t t_to,
                                            const float4* stage, int lane, unsigned long long& regfix) {
  const bool second = code < 0;
  const int row = (int)((unsigned)code & kCodeRowMask);
  const int dim = a.tab[0].dim;
  float* __restrict__ W = second ? a.tab[1].W : a.tab[0].W;
  float* __restrict__ M = second ? a.tab[1].m : a.tab[0].m;
  float* __restrict__ V = second ? a.tab[1].v : a.tab[0].v;
  const int64_t from = t_to - (int64_t)(g - 1);
  double regd = 0.0;
  if ((unsigned)code & kCodeSplit) {
    const int e = (int)(((unsigned)code >> 26) & 15u) * 32 + lane;
    const bool live = e < dim;
    const size_t o = (size_t)row * dim + e;
    const float* sf = reinterpret_cast<const float*>(stage);
    float w = 0.f, m = 0.f, v = 0.f;
    if (live) {
      w = sf[lane];
      m = sf[128 * NV + lane];
      v = sf[256 * NV + lane];
    }
    replay_lane(w, m, v, tabs, from, t_to, a.l2x2, lane, regd);
    if (live) {
      W[o] = w;
      M[o] = m;
      V[o] = v;
    }
  } else {
    const int d4 = dim >> 2;
    const size_t o = (size_t)row * dim;
    RowTile<NV> w, m, v;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      const bool in = j < d4;
      w.x[k] = in ? stage[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      m.x[k] = in ? stage[32 * NV + j] : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x[k] = in ? stage[64 * NV + j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    replay_tile<NV>(w, m, v, tabs, from, t_to, a.l2x2, lane, regd);
    w.store(W + o, d4, lane);
    m.store(M + o, d4, lane);
    v.store(V + o, d4, lane);
  }
  if (a.reg.acc) {
    regd = warp_sum(regd);
    reg_fix_add(regd, a.reg, regfix);
  }
}
// After a fence: the row is at t_to.  A split row becomes current when its last part arrives.
template <bool PEER>
__device__ __forceinline__ void publish_item(const ChunkArgs& a, const PeerExt* px, int code, int64_t t_to, int n_parts) {
  const int row = (int)((unsigned)code & kCodeRowMask);
  int32_t* ls = (code < 0 ? a.tab[1].last_step : a.tab[0].last_step) + row;
  if ((unsigned)code & kCodeSplit) {
    const int old = atomicAdd(ls, 1 << kPartShift);
    if ((old >> kPartShift) == n_parts - 1) {
      // (the other parts' stores: ordered before their arrivals by their fences, cumulative through this one)
      if (PEER && px->strict) __threadfence_system(); else __threadfence();
      *(volatile int32_t*)ls = (int32_t)t_to;
      if (PEER) st_sys_s32(px->rowflag_peer[code < 0 ? 1 : 0][px->me] + row, (int32_t)t_to);
    }
  } else {
    *(volatile int32_t*)ls = (int32_t)t_to;
    if (PEER) st_sys_s32(px->rowflag_peer[code < 0 ? 1 : 0][px->me] + row, (int32_t)t_to);
  }
}

// ---------------------------------------------------------------------------------------------
// replay warps
constexpr int kReplayBatch = 4;  // items per grab
template <int NV, bool PEER>
__device__ void replay_role(const ChunkArgs& a, const PeerExt* px, StepSmem& sm, float4* stage_all) {
  const int lane = threadIdx.x & 31;
  constexpr int kBufs = StageCfg<NV>::kBufs;
  float4* stage = stage_all + (size_t)((threadIdx.x >> 5) - kStepWarps) * (kBufs * 96 * NV);   // row buffers of this warp
  ChunkCtl* ctl = a.ctl;
  const int D = a.depth, Wd = D + 1;
  const int aidx = lane / Wd, k = lane - aidx * Wd;
  const int n_parts = (a.tab[0].dim + 31) / 32;
  const StepTabs tabs{a.alpha, a.reg.stepw, &sm};
  unsigned long long regfix = 0, busy = 0, items = 0, elsteps = 0;
  int sd_idle = -1;            // nothing was available at this many completed steps: wait for the next one
  int sd_seen = -1;
  for (;;) {
    if (*(volatile int*)&sm.quit) break;
    const int sd = *(volatile int*)&sm.steps_done;
    if (sd == sd_idle) {
      __nanosleep(500);
      continue;
    }
    if (sd != sd_seen) {       // order the row loads below after the step warps' barrier acquire
      __threadfence();
      sd_seen = sd;
    }
    // lane (aidx, k): sublist k of step sd + aidx; released iff aidx <= k.  Lowest lane = earliest deadline.
    const int sp = sd + aidx;
    bool ok = lane < Wd * Wd && aidx <= k && sp < a.n_steps;
    int beg = 0, cnt = 0;
    if (ok) {
      const int32_t* sb = a.sched.sub + (size_t)sp * AR_SCHED_SUB;
      beg = sb[k];
      cnt = sb[k + 1] - beg;
      ok = ld_relaxed_s32(a.sched.cursor + (size_t)sp * AR_SCHED_SUB + k) < cnt;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (!mask) {
      if (sd >= a.n_steps - 1) break;      // every item of the chunk has been released and handed out
      sd_idle = sd;                        // cursors only grow: nothing new before the next step completes
      continue;
    }
    const int src = __ffs(mask) - 1;
    const int sp_s = __shfl_sync(0xffffffffu, sp, src), k_s = __shfl_sync(0xffffffffu, k, src);
    const int beg_s = __shfl_sync(0xffffffffu, beg, src), cnt_s = __shfl_sync(0xffffffffu, cnt, src);
    const int R = kReplayBatch;
    int i0 = 0;
    if (lane == 0) i0 = atomicAdd(a.sched.cursor + (size_t)sp_s * AR_SCHED_SUB + k_s, R);
    i0 = __shfl_sync(0xffffffffu, i0, 0);
    if (i0 >= cnt_s) continue;
    const int nit = min(R, cnt_s - i0);
    int code = 0, g = 0;
    if (lane < nit) {
      const size_t at = (size_t)sp_s * a.sched.cap + beg_s + i0 + lane;
      code = a.sched.codes[at];
      g = a.sched.glen[at];
    }
    const long long c0 = clock64();
    const int64_t t_to = a.t0 + sp_s;
    __syncwarp();                // the previous batch's shared-memory reads are done
    item_prefetch<NV>(a, __shfl_sync(0xffffffffu, code, 0), stage, lane);
    for (int j = 0; j < nit; ++j) {
      const int cj = __shfl_sync(0xffffffffu, code, j), gj = __shfl_sync(0xffffffffu, g, j);
      const int cn = __shfl_sync(0xffffffffu, code, (j + 1) & 31);
      if (kBufs == 2 && j + 1 < nit) {   // next item on its way, then wait for this one only
        item_prefetch<NV>(a, cn, stage + (size_t)((j + 1) & 1) * (96 * NV), lane);
        cp_async_wait<1>();
      } else {
        if (kBufs == 1 && j > 0) item_prefetch<NV>(a, cj, stage, lane);
        cp_async_wait<0>();
      }
      __syncwarp();
      replay_item<NV>(a, tabs, cj, gj, t_to, stage + (size_t)(j & (kBufs - 1)) * (96 * NV), lane, regfix);
      elsteps += (unsigned long long)(gj - 1) * (((unsigned)cj & kCodeSplit) ? 32u : (unsigned)a.tab[0].dim);
      __syncwarp();              // buffer j&1 is free for item j+2
    }
    __syncwarp();
    if (PEER && px->strict) __threadfence_system(); else __threadfence();   // every lane's row stores before the flags
    if (lane < nit) publish_item<PEER>(a, px, code, t_to, n_parts);
    busy += (unsigned long long)(clock64() - c0);
    items += nit;
  }
  if (lane == 0) {
    if (a.reg.acc && regfix) atomicAdd(a.reg.acc, regfix);
    atomicAdd(&ctl->stats[0], busy);
    atomicAdd(&ctl->stats[1], items);
    atomicAdd(&ctl->stats[2], elsteps);
    atomicAdd(&ctl->stats[4], 1ull);
  }
}

// AR_ADAM_DENSE: every row the step did not touch takes the same Adam step with the pure L2 gradient.  All warps
// of the grid share the rows (32 consecutive rows per warp visit, one coalesced last_step load).
template <int NV>
__device__ __forceinline__ void dense_pass(const ChunkArgs& a, const StepTabs& tabs, int64_t t, int gwarp,
                                           int n_gwarps, int lane, unsigned long long& regfix, float4* stage) {
  // stage: this warp's staging buffers, kBufs x [W | m | v]; the next row is on its way (cp.async) while the current
  // one takes its step -- the pass is a pure HBM stream (1.13 GB per step at cfg2) and wants bytes in flight
  constexpr int kBufs = StageCfg<NV>::kBufs, kRow4 = 96 * NV;
  const int dim = a.tab[0].dim, d4 = dim >> 2;
#pragma unroll 1
  for (int w = 0; w < 2; ++w) {
    const ar_table tb = a.tab[w];
    auto prefetch = [&](int64_t row, float4* buf) {
      const size_t o = (size_t)row * dim;
      stage_tile<NV>(buf, tb.W + o, d4, lane);
      stage_tile<NV>(buf + 32 * NV, tb.m + o, d4, lane);
      stage_tile<NV>(buf + 64 * NV, tb.v + o, d4, lane);
      cp_async_commit();
    };
    for (int64_t r0 = (int64_t)gwarp * 32; r0 < tb.n_rows; r0 += (int64_t)n_gwarps * 32) {
      const int64_t mine = r0 + lane;
      const int last_l = mine < tb.n_rows ? __ldcg(tb.last_step + mine) : 0x7fffffff;
      unsigned todo = __ballot_sync(0xffffffffu, (int64_t)last_l < t);
      int k = 0;
      if (todo) prefetch(r0 + __ffs(todo) - 1, stage);
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t last = __shfl_sync(0xffffffffu, last_l, j);
        const size_t o = (size_t)(r0 + j) * dim;
        if (kBufs == 2 && todo) {
          prefetch(r0 + __ffs(todo) - 1, stage + (size_t)((k + 1) & 1) * kRow4);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        const float4* b4 = stage + (size_t)(k & (kBufs - 1)) * kRow4;
        RowTile<NV> x, m, v;
        tile_from_stage<NV>(x, b4, d4, lane);
        tile_from_stage<NV>(m, b4 + 32 * NV, d4, lane);
        tile_from_stage<NV>(v, b4 + 64 * NV, d4, lane);
        if (kBufs == 1 && todo) prefetch(r0 + __ffs(todo) - 1, stage);
        double regd = 0.0;
        replay_tile<NV>(x, m, v, tabs, last, t, a.l2x2, lane, regd);
        x.store(tb.W + o, d4, lane);
        m.store(tb.m + o, d4, lane);
        v.store(tb.v + o, d4, lane);
        if (a.reg.acc) {
          regd = warp_sum(regd);
          reg_fix_add(regd, a.reg, regfix);
        }
        ++k;
      }
      if (mine < tb.n_rows && (int64_t)last_l < t) tb.last_step[mine] = (int32_t)t;
    }
  }
}

// The non-step warps in AR_ADAM_DENSE: join the dense pass of every step.
template <int NV>
__device__ void dense_helper_role(const ChunkArgs& a, StepSmem& sm, int n_threads, float4* stage_all) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int wpc = n_threads >> 5;
  const StepTabs tabs{a.alpha, a.reg.stepw, &sm};
  unsigned long long regfix = 0;
  for (int s = 0; s < a.n_steps; ++s) {
    // the dense pass of step s starts when grid barrier 3s+2 has completed
    bool quit = false;
    while (*(volatile int*)&sm.dense_go < s + 1) {
      if (*(volatile int*)&sm.quit) { quit = true; break; }
      __nanosleep(200);
    }
    if (quit) break;
    __threadfence();
    dense_pass<NV>(a, tabs, a.t0 + s + 1, blockIdx.x * wpc + wid, gridDim.x * wpc, lane, regfix,
                   stage_all + (size_t)(wid - kStepWarps) * (StageCfg<NV>::kBufs * 96 * NV));
    __threadfence();
    if (lane == 0) dense_arrive(sm);
  }
  if (lane == 0 && a.reg.acc && regfix) atomicAdd(a.reg.acc, regfix);
}

// ---------------------------------------------------------------------------------------------
// step warps
// finish one row: gradient from the reduced samples + L2 term, Adam step t, store.  w/m/v are already loaded.
// `pend` (AR_ADAM_REPLAY): last_step entry of the row this warp stored BEFORE this one; its flag is published here,
// behind a fence that finds those stores long complete, and this row's flag becomes the pending one.
template <int NV, bool PEER>
__device__ __forceinline__ void finish_loaded(const ChunkArgs& a, const StepTabs& tabs, const ar_table& tb, int row,
                                              const RowTile<NV>& acc, float q, float rinv, RowTile<NV>& w,
                                              RowTile<NV>& m, RowTile<NV>& v, int last, int64_t t, int lane,
                                              double& reg_lane, int32_t*& pend, int32_t*& pend_peer,
                                              int32_t* rowflag = nullptr) {
  const int d4 = tb.dim >> 2;
  const size_t o = (size_t)row * tb.dim;
  const bool flags = a.mode == AR_ADAM_REPLAY;
  if (flags && (int64_t)last != t - 1) {
    // the forward has waited for this row to be at t-1: cannot happen unless the schedule and the plans disagree
    if (lane == 0) atomicAdd(a.health, 1);
  }
  // regulariser term of the step: this lane's share; the warp adds its lanes once per step (its rows are a fixed
  // share of the plan, so the order of the sum is fixed)
  if (a.reg.acc) reg_lane += (double)(tabs.w_at(t) * tile_partial_dot<NV>(w, w));
  const float al = tabs.a_at(t);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 wk = w.x[k], ak = acc.x[k];
    float4 g;
    g.x = __fadd_rn(rinv * (ak.x - q * (wk.x * rinv)), __fmul_rn(a.l2x2, wk.x));
    g.y = __fadd_rn(rinv * (ak.y - q * (wk.y * rinv)), __fmul_rn(a.l2x2, wk.y));
    g.z = __fadd_rn(rinv * (ak.z - q * (wk.z * rinv)), __fmul_rn(a.l2x2, wk.z));
    g.w = __fadd_rn(rinv * (ak.w - q * (wk.w * rinv)), __fmul_rn(a.l2x2, wk.w));
    adam4(w.x[k], m.x[k], v.x[k], g, al);
  }
  if (flags && pend) {
    __threadfence();
    if (lane == 0) {
      *(volatile int32_t*)pend = (int32_t)t;
      if (PEER && pend_peer) st_sys_s32(pend_peer, (int32_t)t);
    }
  }
  w.store(tb.W + o, d4, lane);
  m.store(tb.m + o, d4, lane);
  v.store(tb.v + o, d4, lane);
  if (flags) pend = tb.last_step + row;                       // published later, behind a fence
  else if (lane == 0) tb.last_step[row] = (int32_t)t;         // no per-row flags: a grid barrier follows the update
  if (PEER) pend_peer = rowflag ? rowflag + row : nullptr;
}

// Bulk asynchronous copies (TMA engine, no tensor map): a row travels global -> shared as ONE request that completes on
// an mbarrier.  The per-lane 16-byte cp.async used elsewhere in this file keeps too few bytes in flight per SM once the
// source is microseconds away over NVLink (measured: the forward took 27 / 38 / 43 us at 2 / 4 / 8 GPUs with it).
// For LOCAL rows it is the other way round: the same change in the single-GPU forward and row update (one lane issuing
// four 512-byte bulk copies per pair / row) measured 58.6 us/step against 46.0 -- those stay on cp.async.
#ifndef AR_PEER_TMA
#define AR_PEER_TMA 1
#endif
__device__ __forceinline__ unsigned smem_addr32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_row(unsigned dst, const float* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// pair buffers of the peer forward per step warp (each 4 row tiles); dim <= 128: 3 x 2 KB x 16 warps = 96 KB (measured:
// five buffers request a whole share at once but shrink the L1 and slow the head and the row update more than they gain)
#ifndef AR_PEER_PAIR_BUFS
#define AR_PEER_PAIR_BUFS 3
#endif
template <int NV> struct PeerCfg { static constexpr int kPairBufs = NV == 1 ? AR_PEER_PAIR_BUFS : StageCfg<NV>::kBufs; };
static_assert(PeerCfg<1>::kPairBufs <= 4, "four mbarriers per step warp");
template <int NV, bool PEER> struct StepStage {   // float4 per step warp
  static constexpr int k4 = (PEER ? PeerCfg<NV>::kPairBufs : StageCfg<NV>::kBufs) * 128 * NV;
};

// Peer mode, forward of step s: this warp's share of the rank's two selection lists (entries of the user-side list
// first).  Entry = (my row, local) + (the sample's row of the OTHER table, in its owner's HBM): wait until both are at
// step t-1 (the owner's per-row flags are polled over NVLink), gather, normalise, dot.  Kept for the row update: the
// normalised other row, 1/||my row|| and the cosine, per list entry; the user side also appends (position, cosine) to
// the list the other ranks read, and only its cosines enter the batch sums (every sample once).
template <int NV>
__device__ __forceinline__ void peer_forward(const ChunkArgs& a, const PeerExt& px, StepSmem& sm, int s, int64_t t, int par,
                                             int gw, int ngw, int lane, int wid, float4* sbuf, long long* stamps,
                                             unsigned& ring_phase) {
  constexpr int NB = PeerCfg<NV>::kPairBufs, kBuf4 = 128 * NV;
  const unsigned mb0 = smem_addr32(&sm.pmbar[wid][0]);
  const unsigned sb0 = smem_addr32(sbuf);
  const unsigned row_bytes = (unsigned)a.tab[0].dim * 4u;
  const int dim = a.tab[0].dim, d4 = dim >> 2;
  const int G = px.G;
  const int cu = min(px.cnt[0][s], px.cap), ca = min(px.cnt[1][s], px.cap);
  // this warp: a contiguous share of the user-side list, then one of the anime-side list.  The user-side entries
  // come first because they send words to the other ranks: those stores are long acknowledged when the grid
  // barrier's release fence has to wait for them
  const int RFu = (cu + ngw - 1) / ngw, RFa = (ca + ngw - 1) / ngw;
  const int u0 = min(cu, gw * RFu), u1 = min(cu, u0 + RFu);
  const int a0 = min(ca, gw * RFa), a1 = min(ca, a0 + RFa);
  const int nuw = u1 - u0;
  const int f0 = 0, f1 = nuw + (a1 - a0);
  const size_t my_box = ((size_t)par * G + px.me) * px.cap;   // my sender slot in every rank's inbox of this parity
  const unsigned tag = pair_tag(t) << kPairPosBits;
  {
    // slots that held a word two steps ago and are past the end of this step's list: invalid from now on
    const int prev = __ldcg(px.flags + kPeerLenWord + par);
    for (int k = cu + gw * 32 + lane; k < min(prev, px.cap); k += ngw * 32)
      for (int d = 0; d < G; ++d) st_sys_u64(px.pairs_peer[d] + my_box + k, kPairInvalid);
  }
  float* __restrict__ cl0 = a.c[par];
  float* __restrict__ cl1 = px.cl1[par];
  double sc = 0.0, sc2 = 0.0;
  for (int base = f0; base < f1; base += 32) {
    const int cnt = min(32, f1 - base);
    int T_l = 0, k_l = 0, key_l = 0, own_l = px.me, ol_l = 0, samp_l = -1;
    if (lane < cnt) {
      const int e = base + lane;
      T_l = e >= nuw ? 1 : 0;
      k_l = T_l ? a0 + (e - nuw) : u0 + e;
      const size_t at = (size_t)s * px.cap + k_l;
      key_l = px.key[T_l][at];
      const int og = px.oth[T_l][at];
      own_l = og % G;
      ol_l = og / G;
      if (!T_l) samp_l = px.samp0[at];
    }
    {
      const int want = (int)(t - 1);
      const int32_t* lmine = a.tab[T_l].last_step + key_l;
      const int32_t* lpeer = px.rowflag_peer[1 - T_l][own_l] + ol_l;
      // (polled with relaxed loads -- an acquire load per poll also invalidates the SM's L1 and measured 3 us slower --
      // then read once more with acquire semantics: the row reads below come after it, at system scope)
      bool ready = lane >= cnt || (ld_sys_s32(lpeer) == want && ld_relaxed_s32(lmine) == want);
      unsigned spins = 0;
      unsigned long long t_begin = 0;
      while (!__all_sync(0xffffffffu, ready)) {
        __nanosleep(64);
        if (!ready) ready = ld_sys_s32(lpeer) == want && ld_relaxed_s32(lmine) == want;
        if ((++spins & 255u) == 0u) {
          const unsigned long long now = globaltimer_ns();
          if (!t_begin) t_begin = now;
          else if (now - t_begin > kPeerWaitTimeoutNs || ld_relaxed_s32(&a.ctl->abort)) {
            sm.abort = 1;            // keep going (the barriers must stay matched); the loop ends at the barrier
            atomicExch(&a.ctl->abort, 1);
            st_sys_s32(px.flags + kPeerErrWordC, (int)t);
            break;
          }
        }
      }
      if (lane < cnt) (void)ld_acquire_sys_s32(lpeer);
      __threadfence();
    }
    if (stamps && base == f0) stamps[1] = (long long)globaltimer_ns();
    // entry pairs through the staging buffers: [mine0 | other0 | mine1 | other1]
    const int np = (cnt + 1) >> 1;
    auto prefetch_pair = [&](int pi, float4* buf) {
      const int j = 2 * pi, j1 = min(j + 1, cnt - 1);
      const int b = pi % NB;
#if AR_PEER_TMA
      if (lane == 0) mbar_expect(mb0 + 8u * b, (AR_PEER_TMA == 2 ? 2u : 4u) * row_bytes);
#endif
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jj = h ? j1 : j;
        const int T = __shfl_sync(0xffffffffu, T_l, jj), key = __shfl_sync(0xffffffffu, key_l, jj);
        const int q = __shfl_sync(0xffffffffu, own_l, jj), ol = __shfl_sync(0xffffffffu, ol_l, jj);
#if AR_PEER_TMA
        if (lane == 0) {
          const unsigned dst = sb0 + (unsigned)((b * kBuf4 + (2 * h) * 32 * NV) * sizeof(float4));
          if (AR_PEER_TMA != 2) bulk_row(dst, a.tab[T].W + (size_t)key * dim, row_bytes, mb0 + 8u * b);
          bulk_row(dst + 32u * NV * (unsigned)sizeof(float4), px.W_peer[1 - T][q] + (size_t)ol * dim, row_bytes, mb0 + 8u * b);
        }
        if (AR_PEER_TMA == 2) stage_tile<NV>(buf + (2 * h) * 32 * NV, a.tab[T].W + (size_t)key * dim, d4, lane);
#else
        stage_tile<NV>(buf + (2 * h) * 32 * NV, a.tab[T].W + (size_t)key * dim, d4, lane);
        stage_tile<NV>(buf + (2 * h + 1) * 32 * NV, px.W_peer[1 - T][q] + (size_t)ol * dim, d4, lane);
#endif
      }
#if AR_PEER_TMA != 1
      cp_async_commit();
#endif
    };
#if AR_PEER_TMA
    fence_proxy_async();   // the bulk copies read what the row words announced (generic-proxy acquire above)
#endif
    // ring of NB pair buffers, one mbarrier each: the rows come over NVLink (microseconds away), so as many pairs as
    // shared memory allows are requested before the first one is waited for
    for (int pi = 0; pi < min(NB, np); ++pi) prefetch_pair(pi, sbuf + (size_t)pi * kBuf4);
    for (int pi = 0; pi < np; ++pi) {
      const int j = 2 * pi;
      const bool two = j + 1 < cnt;
#if AR_PEER_TMA
      {
        const int b = pi % NB;
        unsigned spins = 0;
        while (!mbar_try(mb0 + 8u * b, (ring_phase >> b) & 1u)) {
          if ((++spins & 0xfffffu) == 0u && ld_relaxed_s32(&a.ctl->abort)) break;
        }
        ring_phase ^= 1u << b;
      }
#endif
#if AR_PEER_TMA != 1
      const int pending = min(np, pi + NB) - (pi + 1);   // younger groups that may stay in flight
      if (pending <= 0) cp_async_wait<0>();
      else if (pending == 1) cp_async_wait<1>();
      else if (pending == 2) cp_async_wait<2>();
      else cp_async_wait<3>();
#endif
      const float4* b4 = sbuf + (size_t)(pi % NB) * kBuf4;
      RowTile<NV> x0, y0, x1, y1;   // x = my row, y = the other table's row
      tile_from_stage<NV>(x0, b4, d4, lane);
      tile_from_stage<NV>(y0, b4 + 32 * NV, d4, lane);
      tile_from_stage<NV>(x1, b4 + 64 * NV, d4, lane);
      tile_from_stage<NV>(y1, b4 + 96 * NV, d4, lane);
      if (pi + NB < np) {   // the buffer is in registers now
#if AR_PEER_TMA
        __syncwarp();
        fence_proxy_async();   // every lane's reads of the buffer before the engine writes it again
#endif
        prefetch_pair(pi + NB, sbuf + (size_t)(pi % NB) * kBuf4);
      }
      float sx0 = tile_partial_dot<NV>(x0, x0), sy0 = tile_partial_dot<NV>(y0, y0);
      float sx1 = tile_partial_dot<NV>(x1, x1), sy1 = tile_partial_dot<NV>(y1, y1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sx0 += __shfl_xor_sync(0xffffffffu, sx0, o);
        sy0 += __shfl_xor_sync(0xffffffffu, sy0, o);
        sx1 += __shfl_xor_sync(0xffffffffu, sx1, o);
        sy1 += __shfl_xor_sync(0xffffffffu, sy1, o);
      }
      const float rx0 = 1.0f / sqrtf(fmaxf(sx0, kL2NormEps)), ry0 = 1.0f / sqrtf(fmaxf(sy0, kL2NormEps));
      const float rx1 = 1.0f / sqrtf(fmaxf(sx1, kL2NormEps)), ry1 = 1.0f / sqrtf(fmaxf(sy1, kL2NormEps));
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        x0.x[k] = scale4(x0.x[k], rx0);
        y0.x[k] = scale4(y0.x[k], ry0);
        x1.x[k] = scale4(x1.x[k], rx1);
        y1.x[k] = scale4(y1.x[k], ry1);
      }
      // (the products commute, so the anime side gets the very bits the user side publishes)
      float cs0 = tile_partial_dot<NV>(x0, y0), cs1 = tile_partial_dot<NV>(x1, y1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
        cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h && !two) break;
        const int jj = j + h;
        const int T = __shfl_sync(0xffffffffu, T_l, jj), k = __shfl_sync(0xffffffffu, k_l, jj);
        const int sp = __shfl_sync(0xffffffffu, samp_l, jj);
        const float cs = h ? cs1 : cs0, rx = h ? rx1 : rx0;
        float* stash = (T ? a.uh[par] : a.ah[par]) + (size_t)k * dim;
        if (h) y1.store(stash, d4, lane); else y0.store(stash, d4, lane);
        if (lane == 0) {
          (T ? a.ra[par] : a.ru[par])[k] = rx;
          (T ? cl1 : cl0)[k] = cs;
        }
        if (!T && lane < G) {   // to every rank's inbox: position, tag and cosine in ONE 64-bit store
          const unsigned long long word = (unsigned long long)(((unsigned)sp & kPairPosMask) | tag) |
                                          ((unsigned long long)__float_as_uint(cs) << 32);
          st_sys_u64(px.pairs_peer[lane] + my_box + k, word);
        }
        if (!T) {
          sc += (double)cs;
          sc2 += (double)cs * (double)cs;
        }
      }
    }
  }
  if (lane == 0) {
    sm.red[2 * wid] = sc;
    sm.red[2 * wid + 1] = sc2;
  }
}

template <int NV, bool PEER>
__device__ void step_role(const ChunkArgs& a, const PeerExt* px, StepSmem& sm, int n_threads, float4* step_stage) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float4* sbuf = step_stage + (size_t)wid * StepStage<NV, PEER>::k4;
  const int n_ctas = gridDim.x;
  const int gw = blockIdx.x * kStepWarps + wid, ngw = n_ctas * kStepWarps;
  const int dim = a.tab[0].dim, d4 = dim >> 2;
  const bool flags = a.mode == AR_ADAM_REPLAY, dense = a.mode == AR_ADAM_DENSE;
  ChunkCtl* ctl = a.ctl;
  const StepTabs tabs{a.alpha, a.reg.stepw, &sm};
  const bool stamp = blockIdx.x == 0 && tid == 0;
  constexpr int kBufs = StageCfg<NV>::kBufs, kBuf4 = 128 * NV;   // staging buffers of 4 tiles (float4 units)
  unsigned bar = 0;
  unsigned long long regfix = 0;
  const unsigned long long wall0 = stamp ? globaltimer_ns() : 0ull;
  const long long clk0 = stamp ? clock64() : 0ll;
  if (tid < 4) {
    sm.head[tid] = a.head[tid];
    sm.hm[tid] = a.head_m[tid];
    sm.hv[tid] = a.head_v[tid];
  }
  if (tid < 2) sm.bn[tid] = a.bn_moving[tid];
  unsigned ring_phase = 0;      // peer mode: phase parity of this warp's pair-buffer mbarriers
  if (PEER && lane == 0) {
    for (int b = 0; b < 4; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr32(&sm.pmbar[wid][b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  step_bar();

  for (int s = 0; s < a.n_steps; ++s) {
    const int n = min(a.batch, a.plan[0].meta[(size_t)s * 4 + 2]);
    const int64_t t = a.t0 + s + 1;
    // scratch parity: by chunk step on one GPU; by optimizer step in peer mode (the published lists outlive a launch)
    const int par = PEER ? (int)(t & 1) : (s & 1);
    float* __restrict__ uh = a.uh[par];
    float* __restrict__ ah = a.ah[par];
    float* __restrict__ cc = a.c[par];
    float* __restrict__ ruv = a.ru[par];
    float* __restrict__ rav = a.ra[par];
    long long* stamps = a.stamps + (size_t)s * kStamps;
    if (stamp) stamps[0] = (long long)globaltimer_ns();

    // ---- F: forward of this warp's samples
    if constexpr (PEER) {
      peer_forward<NV>(a, *px, sm, s, t, par, gw, ngw, lane, wid, sbuf, stamp ? stamps : nullptr, ring_phase);
    } else {
      const int32_t* __restrict__ iu = a.iu + (size_t)s * a.batch;
      const int32_t* __restrict__ ia = a.ia + (size_t)s * a.batch;
      const float* __restrict__ U = a.tab[0].W;
      const float* __restrict__ A = a.tab[1].W;
      const int RF = (n + ngw - 1) / ngw;
      const int f0 = min(n, gw * RF), f1 = min(n, f0 + RF);
      double sc = 0.0, sc2 = 0.0;
      for (int base = f0; base < f1; base += 32) {
        const int cnt = min(32, f1 - base);
        int my_u = 0, my_a = 0;
        if (lane < cnt) {
          my_u = iu[base + lane];
          my_a = ia[base + lane];
        }
        if (flags) {
          // both rows of every sample must be at step t-1: updated by the previous step, or replayed to there
          const int want = (int)(t - 1);
          const int32_t* lu = a.tab[0].last_step + my_u;
          const int32_t* la = a.tab[1].last_step + my_a;
          bool ready = lane >= cnt || (ld_relaxed_s32(lu) == want && ld_relaxed_s32(la) == want);
          unsigned spins = 0;
          unsigned long long t_begin = 0;
          while (!__all_sync(0xffffffffu, ready)) {
            __nanosleep(64);
            if (!ready) ready = ld_relaxed_s32(lu) == want && ld_relaxed_s32(la) == want;
            if ((++spins & 255u) == 0u) {
              const unsigned long long now = globaltimer_ns();
              if (!t_begin) t_begin = now;
              else if (now - t_begin > kWaitTimeoutNs || ld_relaxed_s32(&ctl->abort)) {
                sm.abort = 1;            // keep going (the barriers below must stay matched); the loop ends at the barrier
                atomicExch(&ctl->abort, 1);
                break;
              }
            }
          }
          __threadfence();
        }
        if (stamp && base == f0) stamps[1] = (long long)globaltimer_ns();
        // sample pairs through the staging buffers: [u0 | a0 | u1 | a1]
        const int np = (cnt + 1) >> 1;
        auto prefetch_pair = [&](int pi, float4* buf) {
          const int j = 2 * pi, j1 = min(j + 1, cnt - 1);
          const int u0 = __shfl_sync(0xffffffffu, my_u, j), a0 = __shfl_sync(0xffffffffu, my_a, j);
          const int u1 = __shfl_sync(0xffffffffu, my_u, j1), a1 = __shfl_sync(0xffffffffu, my_a, j1);
          stage_tile<NV>(buf, U + (size_t)u0 * dim, d4, lane);
          stage_tile<NV>(buf + 32 * NV, A + (size_t)a0 * dim, d4, lane);
          stage_tile<NV>(buf + 64 * NV, U + (size_t)u1 * dim, d4, lane);
          stage_tile<NV>(buf + 96 * NV, A + (size_t)a1 * dim, d4, lane);
          cp_async_commit();
        };
        prefetch_pair(0, sbuf);
        for (int pi = 0; pi < np; ++pi) {
          const int j = 2 * pi;
          const bool two = j + 1 < cnt;
          if (kBufs == 2 && pi + 1 < np) {
            prefetch_pair(pi + 1, sbuf + (size_t)((pi + 1) & 1) * kBuf4);
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          const float4* b4 = sbuf + (size_t)(pi & (kBufs - 1)) * kBuf4;
          RowTile<NV> x0, y0, x1, y1;
          tile_from_stage<NV>(x0, b4, d4, lane);
          tile_from_stage<NV>(y0, b4 + 32 * NV, d4, lane);
          tile_from_stage<NV>(x1, b4 + 64 * NV, d4, lane);
          tile_from_stage<NV>(y1, b4 + 96 * NV, d4, lane);
          if (kBufs == 1 && pi + 1 < np) prefetch_pair(pi + 1, sbuf);   // the buffer is in registers now
          float su0 = tile_partial_dot<NV>(x0, x0), sa0 = tile_partial_dot<NV>(y0, y0);
          float su1 = tile_partial_dot<NV>(x1, x1), sa1 = tile_partial_dot<NV>(y1, y1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            su0 += __shfl_xor_sync(0xffffffffu, su0, o);
            sa0 += __shfl_xor_sync(0xffffffffu, sa0, o);
            su1 += __shfl_xor_sync(0xffffffffu, su1, o);
            sa1 += __shfl_xor_sync(0xffffffffu, sa1, o);
          }
          const float ru0 = 1.0f / sqrtf(fmaxf(su0, kL2NormEps)), ra0 = 1.0f / sqrtf(fmaxf(sa0, kL2NormEps));
          const float ru1 = 1.0f / sqrtf(fmaxf(su1, kL2NormEps)), ra1 = 1.0f / sqrtf(fmaxf(sa1, kL2NormEps));
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            x0.x[k] = scale4(x0.x[k], ru0);
            y0.x[k] = scale4(y0.x[k], ra0);
            x1.x[k] = scale4(x1.x[k], ru1);
            y1.x[k] = scale4(y1.x[k], ra1);
          }
          float cs0 = tile_partial_dot<NV>(x0, y0), cs1 = tile_partial_dot<NV>(x1, y1);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
            cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
          }
          const int s0 = base + j;
          x0.store(uh + (size_t)s0 * dim, d4, lane);
          y0.store(ah + (size_t)s0 * dim, d4, lane);
          if (lane == 0) {
            cc[s0] = cs0;
            ruv[s0] = ru0;
            rav[s0] = ra0;
          }
          sc += (double)cs0;
          sc2 += (double)cs0 * (double)cs0;
          if (two) {
            x1.store(uh + (size_t)(s0 + 1) * dim, d4, lane);
            y1.store(ah + (size_t)(s0 + 1) * dim, d4, lane);
            if (lane == 0) {
              cc[s0 + 1] = cs1;
              ruv[s0 + 1] = ru1;
              rav[s0 + 1] = ra1;
            }
            sc += (double)cs1;
            sc2 += (double)cs1 * (double)cs1;
          }
        }
      }
      if (lane == 0) {
        sm.red[2 * wid] = sc;
        sm.red[2 * wid + 1] = sc2;
      }
    }
    if (PEER && stamp) stamps[7] = (long long)globaltimer_ns();   // this warp's own forward is done; the rest is waiting
    step_bar();
    if (tid == 0) {
      double b0 = 0.0, b1 = 0.0;
#pragma unroll
      for (int w = 0; w < kStepWarps; ++w) {
        b0 += sm.red[2 * w];
        b1 += sm.red[2 * w + 1];
      }
      a.fwd_part[2 * blockIdx.x] = b0;
      a.fwd_part[2 * blockIdx.x + 1] = b1;
    }
    if (!grid_bar(ctl, bar, n_ctas, &sm.abort)) break;
    // every CTA has finished the row update of step s-1 (it came before this step's forward): release the items
    // that were waiting for it
    if (flags && tid == 0) *(volatile int*)&sm.steps_done = s;
    if (stamp) stamps[2] = (long long)globaltimer_ns();

    // ---- U, part 0: what this warp's rows need that does not depend on the head -- issued now, so that the loads
    // (and the first row's gathers) are in flight while the head is computed
    const float* __restrict__ label = PEER ? nullptr : a.label + (size_t)s * a.batch;
    // per listed sample: cosine and label.  One GPU: both tables index the batch.  Peer mode: each table has its own
    // selection list (the anime side computed the same cosine bits itself)
    auto cc_of = [&](int w) -> const float* { return PEER ? (w ? px->cl1[par] : cc) : cc; };
    auto lab_of = [&](int w) -> const float* { return PEER ? px->lab[w] + (size_t)s * px->cap : label; };
    const int nu = a.plan[0].meta[(size_t)s * 4], na = a.plan[1].meta[(size_t)s * 4];
    const int tot = nu + na;
    // item k of this warp is segment gw + k*ngw (users first, then anime): popular rows have neighbouring ids, a
    // contiguous share would hand all of them to one warp
    const int RU = tot > gw ? (tot - gw + ngw - 1) / ngw : 0;
    const int r0 = 0, r1 = RU;
    // lane j: everything row `base + j` needs before its gathers
    int row_l = 0, beg_l = 0, len_l = 0, s0_l = 0, last_l = 0, which_l = 0;
    float lab_l = 0.f, c0_l = 0.f, rinv_l = 0.f;
    auto load_meta = [&](int base, int cnt) {
      const int idx = gw + (base + lane) * ngw;
      which_l = idx >= nu ? 1 : 0;
      if (lane < cnt) {
        const ar_plan& pl = a.plan[which_l];
        const int seg = which_l ? idx - nu : idx;
        row_l = pl.uniq[(size_t)s * pl.batch_cap + seg];
        const int32_t* off = pl.off + (size_t)s * (pl.batch_cap + 1);
        beg_l = off[seg];
        len_l = off[seg + 1] - beg_l;
        s0_l = pl.order[(size_t)s * pl.batch_cap + beg_l];
        c0_l = __ldcg(cc_of(which_l) + s0_l);
        lab_l = __ldg(lab_of(which_l) + s0_l);
        rinv_l = __ldcg((which_l ? rav : ruv) + s0_l);
        last_l = __ldcg(a.tab[which_l].last_step + row_l);
      }
    };
    // row j of the block -> staging buffer [other row of its first sample | W | m | v]
    auto prefetch_row = [&](int j, float4* buf) {
      const int w = __shfl_sync(0xffffffffu, which_l, j), row = __shfl_sync(0xffffffffu, row_l, j);
      const int len = __shfl_sync(0xffffffffu, len_l, j), sx = __shfl_sync(0xffffffffu, s0_l, j);
      if (len <= AR_HEAVY_LEN) {
        const ar_table& tj = a.tab[w];
        stage_tile<NV>(buf, (w ? uh : ah) + (size_t)sx * dim, d4, lane);
        stage_tile<NV>(buf + 32 * NV, tj.W + (size_t)row * dim, d4, lane);
        stage_tile<NV>(buf + 64 * NV, tj.m + (size_t)row * dim, d4, lane);
        stage_tile<NV>(buf + 96 * NV, tj.v + (size_t)row * dim, d4, lane);
      }
      cp_async_commit();
    };
    if (r0 < r1) {
      load_meta(r0, min(32, r1 - r0));
      prefetch_row(0, sbuf);
    }

    // ---- head: identical arithmetic, identical order in every CTA
    {
      double s0 = 0.0, s1 = 0.0;
      int nh = n;               // samples the batch statistics run over (peer mode: the global batch)
      if (!PEER || (blockIdx.x == 0 && wid == 0)) {
        double2 pv[kMaxCtas / 32];
#pragma unroll
        for (int r = 0; r < kMaxCtas / 32; ++r) {     // all loads in flight together
          const int i = lane + 32 * r;
          pv[r] = i < n_ctas ? __ldcg(reinterpret_cast<const double2*>(a.fwd_part) + i) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int r = 0; r < kMaxCtas / 32; ++r) {
          s0 += pv[r].x;
          s1 += pv[r].y;
        }
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if constexpr (PEER) {
        // Cross-rank exchange, no flags and no fences: CTA 0 sends this rank's list length and batch sums to every
        // rank's header inbox as step-tagged words; everyone polls its own inbox until all ranks' words of step t
        // are there.  (The pair words were sent during the forward and are validated one by one below.)
        const int G = px->G;
        if (blockIdx.x == 0 && tid == 0) {
          const unsigned long long b0 = (unsigned long long)__double_as_longlong(s0);
          const unsigned long long b1 = (unsigned long long)__double_as_longlong(s1);
          const int cu = min(px->cnt[0][s], px->cap);
          const unsigned pay[5] = {(unsigned)cu, (unsigned)b0, (unsigned)(b0 >> 32), (unsigned)b1, (unsigned)(b1 >> 32)};
          const unsigned long long tg = (unsigned long long)(unsigned)t << 32;
          for (int d = 0; d < G; ++d) {
            unsigned long long* box = px->hdrin_peer[d] + ((size_t)par * G + px->me) * kHdrWords;
#pragma unroll
            for (int i = 0; i < 5; ++i) st_sys_u64(box + i, (unsigned long long)pay[i] | tg);
          }
          px->flags[kPeerLenWord + par] = cu;   // what the forward of step t+2 has to invalidate behind
        }
        if (tid < G) {
          const unsigned long long* box = px->hdrin_peer[px->me] + ((size_t)par * G + tid) * kHdrWords;
          unsigned long long w[5];
          unsigned spins = 0;
          unsigned long long t_begin = 0;
          for (;;) {
            bool all = true;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              w[i] = ld_sys_u64(box + i);
              all = all && (unsigned)(w[i] >> 32) == (unsigned)t;
            }
            if (all) break;
            if ((++spins & 63u) == 0u) {
              const unsigned long long now = globaltimer_ns();
              if (!t_begin) t_begin = now;
              else if (now - t_begin > kPeerWaitTimeoutNs || ld_relaxed_s32(&ctl->abort)) {
                st_sys_s32(px->flags + kPeerErrWordC, (int)t);
                sm.abort = 1;          // the barriers below stay matched; the loop ends at the next one
                atomicExch(&ctl->abort, 1);
                break;
              }
            }
          }
          sm.pcnt[tid] = (int)(unsigned)w[0];
          sm.ps0[tid] = __longlong_as_double((long long)((w[1] & 0xffffffffull) | (w[2] << 32)));
          sm.ps1[tid] = __longlong_as_double((long long)((w[3] & 0xffffffffull) | (w[4] << 32)));
        }
        step_bar();
        s0 = 0.0;
        s1 = 0.0;
        nh = 0;
        for (int r = 0; r < G; ++r) {   // rank order: identical sums on every rank
          s0 += sm.ps0[r];
          s1 += sm.ps1[r];
          nh += min(max(sm.pcnt[r], 0), px->cap);
        }
        nh = max(nh, 1);
      }
      const int n = nh;   // (shadows the per-rank count from here to the end of the head)
      const HeadScalars h = head_scalars(sm.head, s0, s1, n);
      // second pass, distributed: this CTA's share of the samples (one per thread) -> 5 backward sums + the reported
      // BCE / squared error; per-CTA partials, a second grid barrier, then every CTA adds them in CTA order
      const int mshare = (n + n_ctas - 1) / n_ctas;
      double acc[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) acc[i] = 0.0;
      // peer mode: the samples are every rank's published (position, cosine) pairs, read over NVLink; CTA b takes
      // pairs [b*share, (b+1)*share) of every list, flattened over (rank, pair) so that the loads are in flight together
      int shmax = 0;
      if constexpr (PEER) {
        for (int r = 0; r < px->G; ++r) shmax = max(shmax, (min(max(sm.pcnt[r], 0), px->cap) + n_ctas - 1) / n_ctas);
      }
      const unsigned i_hi = PEER ? (unsigned)(px->G * shmax) : min(n, (blockIdx.x + 1) * mshare);
      for (int i = PEER ? tid : blockIdx.x * mshare + tid; i < i_hi; i += kStepThreads) {
        float ci, ti;
        if constexpr (PEER) {
          const int r = i / shmax, k = blockIdx.x * shmax + (i - r * shmax);
          if (k >= min(sm.pcnt[r], px->cap)) continue;
          const unsigned long long* pp = px->pairs_peer[px->me] + ((size_t)par * px->G + r) * px->cap + k;
          const unsigned want = pair_tag(t);
          unsigned long long pw = ld_sys_u64(pp);
          unsigned spins = 0;
          unsigned long long t_begin = 0;
          // valid = this step's tag and a position inside the global batch (an invalidated slot has neither)
          while ((unsigned)pw >> kPairPosBits != want || ((unsigned)pw & kPairPosMask) >= (unsigned)px->gb) {
            if ((++spins & 63u) == 0u) {
              const unsigned long long now = globaltimer_ns();
              if (!t_begin) t_begin = now;
              else if (now - t_begin > kPeerWaitTimeoutNs || ld_relaxed_s32(&ctl->abort)) {
                st_sys_s32(px->flags + kPeerErrWordC, (int)t);
                sm.abort = 1;
                atomicExch(&ctl->abort, 1);
                break;
              }
            }
            pw = ld_sys_u64(pp);
          }
          const int j = (int)min((unsigned)pw & kPairPosMask, (unsigned)px->gb - 1u);
          ci = __uint_as_float((unsigned)(pw >> 32));
          ti = __ldg(px->label_step + (size_t)s * px->gb + j);
        } else {
          ci = __ldcg(cc + i);
          ti = __ldg(label + i);
        }
        const float zh = ((h.w * ci + h.b) - h.mu) * h.inv;
        const float y = h.gamma * zh + h.beta;
        const float p = sigmoidf_(y);
        const float dy = (p - ti) * h.rn;
        acc[0] += (double)dy;
        acc[1] += (double)(dy * zh);
        acc[2] += (double)zh;
        acc[3] += (double)(dy * ci);
        acc[4] += (double)(zh * ci);
        acc[5] += (double)bce_logits(y, ti);
        acc[6] += (double)((ti - p) * (ti - p));
      }
      step_block_sum<7>(acc, sm.red);
      if (tid < 7) a.hsum[(size_t)blockIdx.x * 8 + tid] = acc[tid];
      if (tid == 64) {
        double* mp = a.mpart + ((size_t)s * kMaxCtas + blockIdx.x) * 2;
        mp[0] = acc[5];
        mp[1] = acc[6];
      }
      if (!grid_bar(ctl, bar, n_ctas, &sm.abort)) break;
      if (wid < 5) {            // warp i adds sum i over the CTAs, in CTA order
        double pv[kMaxCtas / 32];
#pragma unroll
        for (int r = 0; r < kMaxCtas / 32; ++r) {     // all loads in flight together
          const int c = lane + 32 * r;
          pv[r] = c < n_ctas ? __ldcg(a.hsum + (size_t)c * 8 + wid) : 0.0;
        }
        double t5 = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxCtas / 32; ++r) t5 += pv[r];
        t5 = warp_sum(t5);
        if (lane == 0) sm.red[wid] = t5;
      }
      step_bar();
#pragma unroll
      for (int i = 0; i < 5; ++i) acc[i] = sm.red[i];
      const double S1 = acc[0], S2 = acc[1], Szh = acc[2], Sdyc = acc[3], Szhc = acc[4], Sc = s0;
      const double ig = (double)h.inv * (double)h.gamma;
      if (tid < 4) {
        // oracle head_backward(): dgamma = sum dy*zh, dbeta = sum dy, dz_i = inv*gamma*(dy_i - S1/n - zh_i*S2/n)
        const float g = tid == 0 ? (float)(ig * (Sdyc - S1 / n * Sc - S2 / n * Szhc))   // dw = sum dz*c
                      : tid == 1 ? (float)(-ig * (S2 / n) * Szh)                        // db = sum dz
                      : tid == 2 ? (float)S2 : (float)S1;
        float th = sm.head[tid], hm = sm.hm[tid], hv = sm.hv[tid];
        adam1(th, hm, hv, g, tabs.a_at(t));
        sm.head[tid] = th;
        sm.hm[tid] = hm;
        sm.hv[tid] = hv;
      } else if (tid == 32) {
        const float mm = sm.bn[0], mv = sm.bn[1];
        sm.bn[0] = mm - (mm - h.mu) * kBnOneMinusMomentum;
        sm.bn[1] = mv - (mv - h.var) * kBnOneMinusMomentum;
      } else if (tid == 64) {
        sm.stepc[K_COEF] = (float)((double)h.w * ig);
        sm.stepc[K_S1N] = (float)(S1 / n);
        sm.stepc[K_S2N] = (float)(S2 / n);
        sm.stepc[K_MU] = h.mu;
        sm.stepc[K_INV] = h.inv;
        sm.stepc[K_W] = h.w;
        sm.stepc[K_B] = h.b;
        sm.stepc[K_GAMMA] = h.gamma;
        sm.stepc[K_BETA] = h.beta;
        sm.stepc[K_FN] = h.fn;
        sm.stepc[K_RN] = h.rn;
        if (blockIdx.x == 0) {
          float* row = a.metrics + t * 4;
          row[2] = h.fn;
          row[3] = h.mu;
        }
      }
      step_bar();
    }
    if (stamp) stamps[3] = (long long)globaltimer_ns();

    // ---- U: this warp's share of the step's distinct rows (users first, then anime), through the staging buffers
    {
      const float* kk = sm.stepc;   // read at the use sites: shared-memory loads are cheaper than 11 live registers
      int32_t* pend = nullptr;      // flag of the row stored last, published behind the next row's arithmetic
      double reg_lane = 0.0;
      int hv_n = 0, hv_row = 0;     // strict peer mode: heavy rows this warp finished (lane i: the i-th; bit 31 = anime)
      int32_t* pend_peer = nullptr; // default peer mode: the pending row's word for the peers
      const bool strict = PEER && px->strict;
      auto rowflag_of = [&](int w) -> int32_t* { return (PEER && !strict) ? px->rowflag_peer[w][px->me] : nullptr; };
      // heavy rows FIRST (their flags are what the next forward waits for longest): pieces of AR_HEAVY_LEN samples
      // spread over all step warps of the grid, the last piece to arrive finishes the row
      for (int w = 0; w < 2; ++w) {
        const ar_plan& pl = a.plan[w];
        const int nh = pl.meta[(size_t)s * 4 + 1];
        if (nh == 0) continue;
        const ar_table& tb = a.tab[w];
        const float* __restrict__ other = w ? uh : ah;
        const float* __restrict__ ccw = cc_of(w);
        const float* __restrict__ labw = lab_of(w);
        const int32_t* off = pl.off + (size_t)s * (pl.batch_cap + 1);
        const int32_t* order = pl.order + (size_t)s * pl.batch_cap;
        int piece0 = 0;   // pieces of the heavy rows before this round
        for (int hb = 0; hb < nh; hb += 32) {
          int seg_l = 0, beg_h = 0, len_h = 0, np_l = 0;
          if (hb + lane < nh) {
            seg_l = pl.heavy[(size_t)s * pl.heavy_cap + hb + lane];
            beg_h = off[seg_l];
            len_h = off[seg_l + 1] - beg_h;
            np_l = (len_h + AR_HEAVY_LEN - 1) / AR_HEAVY_LEN;
          }
          int incl = np_l;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
          }
          const int excl = incl - np_l + piece0;
          const int round_total = __shfl_sync(0xffffffffu, incl, 31);
          // pieces [piece0, piece0 + round_total): mine are those == gw (mod ngw)
          int p = piece0 + ((gw - piece0) % ngw + ngw) % ngw;
          for (; p < piece0 + round_total; p += ngw) {
            const unsigned le = __ballot_sync(0xffffffffu, np_l > 0 && excl <= p);
            const int owner = 31 - __clz(le);
            const int seg = __shfl_sync(0xffffffffu, seg_l, owner), beg = __shfl_sync(0xffffffffu, beg_h, owner);
            const int len = __shfl_sync(0xffffffffu, len_h, owner), np = __shfl_sync(0xffffffffu, np_l, owner);
            const int first = __shfl_sync(0xffffffffu, excl, owner);
            const int piece = p - first;
            const int lo = beg + piece * AR_HEAVY_LEN, hi = min(beg + len, lo + AR_HEAVY_LEN);
            RowTile<NV> acc;
            acc.zero();
            float q = 0.f;
            for (int e = lo; e < hi; ++e) {
              const int sx = order[e];
              RowTile<NV> o;
              tile_load_cg<NV>(o, other + (size_t)sx * dim, d4, lane);
              const float cx = __ldcg(ccw + sx);
              const float dx = dc_of_label(cx, __ldg(labw + sx), kk);
              q = fmaf(dx, cx, q);
#pragma unroll
              for (int k = 0; k < NV; ++k) acc.x[k] = fma4(dx, o.x[k], acc.x[k]);
            }
            const int pslot = min(p, a.n_pieces - 1);
            float* hp = a.hpart + ((size_t)w * a.n_pieces + pslot) * (dim + 4);
            acc.store(hp, d4, lane);
            if (lane == 0) hp[dim] = q;
            __threadfence();
            int old = 0;
            int32_t* tick = a.htick + (size_t)w * a.n_pieces + min(first, a.n_pieces - 1);
            if (lane == 0) old = atomicAdd(tick, 1);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == np - 1) {   // last piece of the row: add the pieces in order, finish the row
              __threadfence();
              if (lane == 0) *tick = 0;
              acc.zero();
              q = 0.f;
              for (int e = 0; e < np; ++e) {
                const float* src = a.hpart + ((size_t)w * a.n_pieces + first + e) * (dim + 4);
                RowTile<NV> pp;
                tile_load_cg<NV>(pp, src, d4, lane);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                  acc.x[k].x += pp.x[k].x;
                  acc.x[k].y += pp.x[k].y;
                  acc.x[k].z += pp.x[k].z;
                  acc.x[k].w += pp.x[k].w;
                }
                q += __ldcg(src + dim);
              }
              const int row = pl.uniq[(size_t)s * pl.batch_cap + seg];
              RowTile<NV> x, m, v;
              tile_load_cg<NV>(x, tb.W + (size_t)row * dim, d4, lane);
              tile_load_cg<NV>(m, tb.m + (size_t)row * dim, d4, lane);
              tile_load_cg<NV>(v, tb.v + (size_t)row * dim, d4, lane);
              const float rinv = __ldcg((w ? rav : ruv) + order[beg]);
              finish_loaded<NV, PEER>(a, tabs, tb, row, acc, q, rinv, x, m, v, __ldcg(tb.last_step + row), t, lane, reg_lane, pend,
                                      pend_peer, rowflag_of(w));
              if (strict) {
                if (hv_n == 32) {   // (never in practice: a warp finishing more than 32 heavy rows of one step)
                  __threadfence_system();
                  st_sys_s32(px->rowflag_peer[hv_row < 0 ? 1 : 0][px->me] + (hv_row & 0x7fffffff), (int32_t)t);
                  hv_n = 0;
                }
                if (lane == hv_n) hv_row = row | (w ? (int)0x80000000 : 0);
                ++hv_n;
              }
            }
          }
          piece0 += round_total;
        }
      }
      for (int base = r0; base < r1; base += 32) {
        const int cnt = min(32, r1 - base);
        if (base != r0) {
          load_meta(base, cnt);
          prefetch_row(0, sbuf);
        }
        const float d0_l = dc_of_label(c0_l, lab_l, kk);
        for (int j = 0; j < cnt; ++j) {
          if (kBufs == 2 && j + 1 < cnt) {
            prefetch_row(j + 1, sbuf + (size_t)((j + 1) & 1) * kBuf4);
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          const float4* b4 = sbuf + (size_t)(j & (kBufs - 1)) * kBuf4;
          const int wA = __shfl_sync(0xffffffffu, which_l, j), rowA = __shfl_sync(0xffffffffu, row_l, j);
          const int lenA = __shfl_sync(0xffffffffu, len_l, j);
          if (lenA <= AR_HEAVY_LEN) {
            RowTile<NV> accA, wa, ma, va;
            tile_from_stage<NV>(accA, b4, d4, lane);
            tile_from_stage<NV>(wa, b4 + 32 * NV, d4, lane);
            tile_from_stage<NV>(ma, b4 + 64 * NV, d4, lane);
            tile_from_stage<NV>(va, b4 + 96 * NV, d4, lane);
            if (kBufs == 1 && j + 1 < cnt) prefetch_row(j + 1, sbuf);   // the buffer is in registers now
            const ar_table& tA = a.tab[wA];
            const float* __restrict__ otherA = wA ? uh : ah;
            const float* __restrict__ ccA = cc_of(wA);
            const float* __restrict__ labA = lab_of(wA);
            const float d0 = __shfl_sync(0xffffffffu, d0_l, j), c0 = __shfl_sync(0xffffffffu, c0_l, j);
            float q = fmaf(d0, c0, 0.f);
#pragma unroll
            for (int k = 0; k < NV; ++k) accA.x[k] = fma4(d0, accA.x[k], make_float4(0.f, 0.f, 0.f, 0.f));
            if (lenA > 1) {
              const ar_plan& pl = a.plan[wA];
              const int32_t* order = pl.order + (size_t)s * pl.batch_cap + __shfl_sync(0xffffffffu, beg_l, j);
              for (int e = 1; e < lenA; ++e) {
                const int sx = order[e];
                RowTile<NV> o;
                tile_load_cg<NV>(o, otherA + (size_t)sx * dim, d4, lane);
                const float cx = __ldcg(ccA + sx);
                const float dx = dc_of_label(cx, __ldg(labA + sx), kk);
                q = fmaf(dx, cx, q);
#pragma unroll
                for (int k = 0; k < NV; ++k) accA.x[k] = fma4(dx, o.x[k], accA.x[k]);
              }
            }
            finish_loaded<NV, PEER>(a, tabs, tA, rowA, accA, q, __shfl_sync(0xffffffffu, rinv_l, j), wa, ma, va,
                              __shfl_sync(0xffffffffu, last_l, j), t, lane, reg_lane, pend, pend_peer, rowflag_of(wA));
          } else if (kBufs == 1 && j + 1 < cnt) {
            prefetch_row(j + 1, sbuf);
          }
        }
      }
      if (flags && pend) {          // the last row this warp stored
        __threadfence();
        if (lane == 0) {
          *(volatile int32_t*)pend = (int32_t)t;
          if (PEER && pend_peer) st_sys_s32(pend_peer, (int32_t)t);
        }
      }
      if (strict) {
        // for the peers: ONE system-scope fence behind all of this warp's rows of the step, then their row words
        __threadfence_system();
        if (lane < hv_n) st_sys_s32(px->rowflag_peer[hv_row < 0 ? 1 : 0][px->me] + (hv_row & 0x7fffffff), (int32_t)t);
        for (int base = r0; base < r1; base += 32) {
          const int k = base + lane;
          if (k < r1) {
            const int idx = gw + k * ngw;
            const int w = idx >= nu ? 1 : 0;
            const ar_plan& pl = a.plan[w];
            const int seg = w ? idx - nu : idx;
            const int32_t* off = pl.off + (size_t)s * (pl.batch_cap + 1);
            if (off[seg + 1] - off[seg] <= AR_HEAVY_LEN)
              st_sys_s32(px->rowflag_peer[w][px->me] + pl.uniq[(size_t)s * pl.batch_cap + seg], (int32_t)t);
          }
        }
      }
      if (a.reg.acc) reg_fix_add(warp_sum(reg_lane), a.reg, regfix);
    }
    if (stamp) stamps[4] = (long long)globaltimer_ns();
    if (!flags) {                   // no per-row flags: everyone's rows before anyone's next forward
      if (!grid_bar(ctl, bar, n_ctas, &sm.abort)) break;
      if (tid == 0 && dense) *(volatile int*)&sm.dense_go = s + 1;
    }
    if (stamp) stamps[5] = (long long)globaltimer_ns();
    if (dense) {
      const int wpc = n_threads >> 5;
      dense_pass<NV>(a, tabs, t, blockIdx.x * wpc + wid, n_ctas * wpc, lane, regfix, sbuf);
      __threadfence();
      if (tid == 0) {   // the helper warps of this CTA have finished their share of the pass too
        const int want = (s + 1) * (wpc - kStepWarps);
        unsigned spins = 0;
        unsigned long long t_begin = 0;
        while (*(volatile int*)&sm.dense_arrive < want) {
          if ((++spins & 1023u) == 0u) {
            const unsigned long long now = globaltimer_ns();
            if (!t_begin) t_begin = now;
            else if (now - t_begin > kWaitTimeoutNs || ld_relaxed_s32(&ctl->abort)) {
              sm.abort = 1;
              atomicExch(&ctl->abort, 1);
              break;
            }
          }
        }
      }
      if (!grid_bar(ctl, bar, n_ctas, &sm.abort)) break;
      if (stamp) stamps[6] = (long long)globaltimer_ns();
    }
  }
  // the last step's row updates, grid-wide, before the kernel's results are read (and the metrics reduced below)
  if (flags && !sm.abort) {
    grid_bar(ctl, bar, n_ctas, &sm.abort);
    if (tid == 0) *(volatile int*)&sm.steps_done = a.n_steps;
  }

  // ---- epilogue
  if (tid == 0 && sm.abort) *(volatile int*)&sm.quit = 1;
  if (sm.abort) {
    if (tid == 0) atomicAdd(a.health + 1, 1);
    if (lane == 0 && a.reg.acc && regfix) atomicAdd(a.reg.acc, regfix);
    return;
  }
  if (blockIdx.x == 0) {
    if (tid < 4) {
      a.head[tid] = sm.head[tid];
      a.head_m[tid] = sm.hm[tid];
      a.head_v[tid] = sm.hv[tid];
    }
    if (tid < 2) a.bn_moving[tid] = sm.bn[tid];
  }
  // reported metrics: per-step sums of the per-CTA partials, fixed order
  for (int s = gw; s < a.n_steps; s += ngw) {
    // (peer mode: the global batch's size, as CTA 0 recorded it with the step's statistics)
    const int n = PEER ? max(1, (int)__ldcg(a.metrics + (a.t0 + s + 1) * 4 + 2))
                       : min(a.batch, a.plan[0].meta[(size_t)s * 4 + 2]);
    double b0 = 0.0, b1 = 0.0;
    for (int i = lane; i < n_ctas; i += 32) {
      b0 += __ldcg(a.mpart + ((size_t)s * kMaxCtas + i) * 2);
      b1 += __ldcg(a.mpart + ((size_t)s * kMaxCtas + i) * 2 + 1);
    }
    b0 = warp_sum(b0);
    b1 = warp_sum(b1);
    if (lane == 0) {
      float* row = a.metrics + (a.t0 + s + 1) * 4;
      row[0] = (float)(b0 / n);
      row[1] = (float)(b1 / n);
    }
  }
  if (lane == 0 && a.reg.acc && regfix) atomicAdd(a.reg.acc, regfix);
  if (stamp) {
    ctl->stats[3] = globaltimer_ns() - wall0;
    ctl->stats[5] = (unsigned long long)(clock64() - clk0);
  }
}

// Register budget.  NV == 1 (dim <= 128, the headline shape): AR_CHUNK_THREADS launched at AR_REGS_LAUNCH; the
// replay warps shrink to AR_REGS_REPLAY, the step warps grow to AR_REGS_STEP; what the CTA leaves of the SM's 64 K
// registers (and 1280 of 2048 threads) lets one 512-thread plan CTA of the NEXT chunk run beside it.
// Larger rows: 768 threads at 80 registers, 56 / 88 after the hand-over (the whole register file).
template <int NV> struct ChunkRegs {
  static constexpr int kReplay = NV == 1 ? AR_REGS_REPLAY : 56;
  static constexpr int kStep = NV == 1 ? AR_REGS_STEP : 88;
};
static_assert((AR_CHUNK_THREADS - kStepThreads) * (AR_REGS_LAUNCH - AR_REGS_REPLAY) >= kStepThreads * (AR_REGS_STEP - AR_REGS_LAUNCH),
              "the replay warps do not free enough registers for the step warps");
static_assert((768 - kStepThreads) * (80 - 56) >= kStepThreads * (88 - 80), "register hand-over (NV > 1)");

template <int NV, int THREADS, bool PEER>
__device__ __forceinline__ void chunk_body(const ChunkArgs& a, const PeerExt* px) {
  __shared__ StepSmem sm;
  extern __shared__ float4 stage_all[];   // replay warps: [warp][kBufs][3 tiles]; then step warps: [warp][kBufs][4 tiles]
  {
    const long long t_hi = a.t0 + a.n_steps;
    const long long base = t_hi - kWin + 1 > 0 ? t_hi - kWin + 1 : 0;
    if (threadIdx.x == 0) {
      sm.steps_done = 0;
      sm.dense_go = 0;
      sm.dense_arrive = 0;
      sm.quit = 0;
      sm.abort = 0;
      sm.win_base = base;
    }
    for (int i = threadIdx.x; i < kWin; i += THREADS) {
      const long long t = base + i;
      sm.alpha_w[i] = t <= t_hi ? a.alpha[t] : 0.f;
      sm.stepw_w[i] = (a.reg.stepw && t <= t_hi) ? a.reg.stepw[t] : 1.f;
    }
  }
  __syncthreads();
  // Register hand-over (setmaxnreg, per warpgroup of 4 warps): the replay loop needs few registers, the step warps
  // keep two rows / samples and their gathers in flight and were spilling their loads at the kernel-wide 64.
  constexpr int kRegsReplay = ChunkRegs<NV>::kReplay, kRegsStep = ChunkRegs<NV>::kStep;
  if (threadIdx.x >= kStepThreads) {
    if (kRegsReplay) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsReplay ? kRegsReplay : 24));
    if (a.mode == AR_ADAM_REPLAY) replay_role<NV, PEER>(a, px, sm, stage_all);
    else if (!PEER && a.mode == AR_ADAM_DENSE) dense_helper_role<NV>(a, sm, THREADS, stage_all);
    return;
  }
  if (kRegsStep) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsStep ? kRegsStep : 24));
  step_role<NV, PEER>(a, px, sm, THREADS, stage_all + (size_t)(THREADS / 32 - kStepWarps) * (StageCfg<NV>::kBufs * 96 * NV));
}

template <int NV, int THREADS, int REGS>
__global__ void __maxnreg__(REGS) chunk_kernel(const __grid_constant__ ChunkArgs a) {
  chunk_body<NV, THREADS, false>(a, nullptr);
}
// the same kernel over row-sharded tables and NVLink peer memory (AR_ADAM_REPLAY only), csrc/peer.inl
template <int NV, int THREADS, int REGS>
__global__ void __maxnreg__(REGS) peer_chunk_kernel(const __grid_constant__ ChunkArgs a, const __grid_constant__ PeerExt px) {
  chunk_body<NV, THREADS, true>(a, &px);
}

template <int NV> struct ChunkCfg {
  static constexpr int kThreads = NV == 1 ? AR_CHUNK_THREADS : 768;
  static constexpr int kRegs = NV == 1 ? AR_REGS_LAUNCH : 80;
};

template <int NV>
static int launch_chunk_nv(const ChunkArgs& a, const PeerExt* px, cudaStream_t st) {
  constexpr int T = ChunkCfg<NV>::kThreads, R = ChunkCfg<NV>::kRegs;
  static int max_ctas[2] = {-1, -1};
  const size_t dyn = ((size_t)(T / 32 - kStepWarps) * 96 * StageCfg<NV>::kBufs * NV +
                      (size_t)kStepWarps * (px ? StepStage<NV, true>::k4 : StepStage<NV, false>::k4)) * sizeof(float4);
  const int which = px ? 1 : 0;
  if (max_ctas[which] < 0) {
    int per_sm = 0;
    if (px) {
      AR_CUDA(cudaFuncSetAttribute(peer_chunk_kernel<NV, T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      AR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer_chunk_kernel<NV, T, R>, T, dyn));
    } else {
      AR_CUDA(cudaFuncSetAttribute(chunk_kernel<NV, T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      AR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chunk_kernel<NV, T, R>, T, dyn));
    }
    AR_REQUIRE(per_sm >= 1, "ar_train_steps: the step kernel does not fit on an SM");
    max_ctas[which] = std::min(num_sms(), kMaxCtas);
  }
  // peer mode leaves a few SMs to the side stream: the next chunk's planning (an NCCL all-gather, the selection and
  // the plan sorts, whose CTAs do not fit beside a step CTA) then overlaps this chunk instead of following it
  static const int reserve = getenv("AR_PEER_RESERVE_SMS") ? atoi(getenv("AR_PEER_RESERVE_SMS")) : 20;
  const int peer_ctas = std::max(1, max_ctas[which] - std::max(0, reserve));
  if (px) peer_chunk_kernel<NV, T, R><<<peer_ctas, T, dyn, st>>>(a, *px);
  else chunk_kernel<NV, T, R><<<max_ctas[which], T, dyn, st>>>(a);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

// px != null: peer mode (the caller fills everything of *px but cl1)
static int launch_chunk(const ar_train_ctx& x, int64_t epoch_step0, int64_t t0, int n_steps, cudaStream_t st,
                        PeerExt* px = nullptr) {
  const ChunkLayout l = chunk_layout(x.plan_u.n_slots, std::max(x.plan_u.batch_cap, x.plan_a.batch_cap), x.users.dim);
  char* ws = (char*)x.chunk_ws;
  ChunkArgs a{};
  a.tab[0] = x.users;
  a.tab[1] = x.anime;
  a.plan[0] = x.plan_u;
  a.plan[1] = x.plan_a;
  a.sched = x.sched;
  const int64_t base0 = epoch_step0 * (int64_t)x.batch;
  a.iu = px ? nullptr : x.iu + base0;      // peer mode works from the selection lists
  a.ia = px ? nullptr : x.ia + base0;
  a.label = px ? nullptr : x.label + base0;
  a.t0 = t0;
  a.n_steps = n_steps;
  a.batch = x.batch;
  a.mode = x.mode;
  a.depth = std::max(1, std::min((int)x.depth, AR_SCHED_MAX_DEPTH));
  a.l2x2 = (float)(2.0 * (double)x.l2);
  a.alpha = x.alpha;
  a.head = x.head;
  a.head_m = x.head_m;
  a.head_v = x.head_v;
  a.bn_moving = x.bn_moving;
  const int bc = std::max(x.plan_u.batch_cap, x.plan_a.batch_cap);
  float* st1 = (float*)(ws + l.stash1);
  a.uh[0] = x.uh;
  a.ah[0] = x.ah;
  a.c[0] = x.c;
  a.ru[0] = x.ru;
  a.ra[0] = x.ra;
  a.uh[1] = st1;
  a.ah[1] = st1 + (size_t)bc * x.users.dim;
  a.c[1] = st1 + (size_t)2 * bc * x.users.dim;
  a.ru[1] = a.c[1] + bc;
  a.ra[1] = a.ru[1] + bc;
  if (px) {
    px->cl1[0] = x.dy;                     // (sel_cap) scratch the staged kernels use for dy
    px->cl1[1] = a.ra[1] + bc;             // the fourth per-sample vector of the odd-parity stash
  }
  a.metrics = x.metrics;
  a.reg = reg_of(x);
  a.health = x.health;
  a.ctl = (ChunkCtl*)(ws + l.ctl);
  a.htick = (int32_t*)(ws + l.htick);
  a.fwd_part = (double*)(ws + l.fpart);
  a.hsum = (double*)(ws + l.hsum);
  a.mpart = (double*)(ws + l.mpart);
  a.stamps = (long long*)(ws + l.stamps);
  a.hpart = (float*)(ws + l.hpart);
  a.n_pieces = l.n_pieces;
  AR_CUDA(cudaMemsetAsync(ws, 0, l.zero_bytes, st));
  int rc = AR_OK;
  AR_DISPATCH_NV(x.users.dim, rc = launch_chunk_nv<NV>(a, px, st));
  return rc;
}

}  // namespace ar
