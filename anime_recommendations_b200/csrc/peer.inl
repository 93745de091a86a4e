// Multi-GPU training over NVLink PEER MEMORY: row-sharded tables, owner-computes, no collective library on
// the step's critical path (DESIGN.md "peer mode").  Included by train.cu.
//
// Same ownership rule as shard.inl (global row g lives on rank g % G at local index g / G), but the exchange
// is turned around: instead of shipping rows to the samples and partial gradients back (3 NCCL all-to-alls +
// 1 all-gather per step, ~45 us of launch/rendezvous latency each), every rank processes the samples of the
// GLOBAL batch that touch ITS rows and pulls the one thing it lacks -- the sample's row of the other table --
// straight out of the owner's HBM with 128-bit loads over NVLink (cudaIpc-mapped shards).
//
// Per chunk of steps (off the critical path): the chunk's (user, anime, label) triples of all ranks are
// all-gathered once; peer_select lists, per step and table, the samples whose row this rank owns (stable, so
// every sum below has a fixed order), and the ordinary dedup plan is built over those lists.
// Per step t:
//   look-ahead catch-up of my rows (side stream, as on one GPU)
//   peer_fwd   first CTA: flag 2t "my rows are current"; every CTA waits for all ranks' 2t, then warp per
//              listed sample: my row (local) + the other table's row (NVLink) -> 1/||.||, cosine c; the
//              normalised other row is stashed locally for the gradient; the user side appends (sample, c) to
//              my PUBLISHED list (local stores); last CTA: flag 2t+1 "my cosines are published"
//   peer_pull  waits for all ranks' 2t+1, reads every rank's published list (coalesced 8-byte NVLink loads),
//              scatters c into the local global-batch array and reduces (sum c, sum c^2) in list order
//   head_step  over the global batch, redundantly on every rank (SyncBN; the 4 head scalars stay bit-identical
//              because every rank sums the same lists in the same order)
//   rows_update  the single-GPU kernel on my rows: segment sums over the stash, catch-up-free Adam
// NVLink traffic per step and GPU: the pulled rows, ~2*B*dim*4 bytes (10 MB at B = 10000, dim = 128; measured
// ~610 GB/s, i.e. bandwidth-bound) plus 8*G*B bytes of cosines.  All of it is READS -- the only peer stores are
// the flag words, so no CTA needs a system-scope fence (a first version scattered c to every rank with 4-byte
// peer stores + a __threadfence_system per CTA: ~10 us slower per step).  Nothing is staged, packed or merged,
// and a row's gradient is summed by one warp in plan order.
//
// Flag barriers: rank r's arrival is the epoch number stored into word r of EVERY rank's flag array (relaxed
// system-scope stores behind at most one system fence); a waiter spins on its own array with relaxed
// system-scope loads and reads peer data only with L1-bypassing loads.  Epochs are derived from the optimizer
// step, so they only grow; a rank that waits longer than kPeerTimeoutNs raises a sticky error word instead of
// hanging the GPU.  Flag words: [0, 8) arrivals, [16, 24) published-list lengths, 32 error, 40/41 local
// ticket / last announced epoch.
namespace ar {

constexpr int kPeerMaxRanks = AR_PEER_MAX_RANKS;
constexpr int kPeerErrWord = 32;                       // flags[32]: sticky "a barrier timed out" (epoch that did)
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void st_release_sys(int32_t* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// peer rows: system-coherent loads that never allocate in this GPU's L1
__device__ __forceinline__ float4 ld4_sys(const float* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

// The training step splits the barrier and folds both halves into its kernels: the FIRST CTA of the forward to
// start signals epoch 2t ("my rows are current" -- everything before it in the stream is complete), its LAST
// CTA to finish signals 2t+1 ("my cosines are published"), and the first kernel that needs the other ranks'
// data waits in its prologue -- its CTAs are already resident when the flags land.  No barrier launches, no
// launch gaps around them (a separate signal kernel cost 5 us of gap before the forward, measured).
// AR_PEER_LOG=1: %globaltimer stamps of one step, written by the kernels themselves (no events, no
// perturbation): 0 fwd start, 1 fwd past its wait, 2 fwd last CTA done, 3 pull start, 4 pull past its wait,
// 5 pull done
constexpr int kLogStamps = 8;

// my arrival at `epoch`, to every rank: G relaxed system-scope stores behind at most ONE system-scope fence (a
// st.release.sys per rank repeats the fence G times, ~1 us each: measured as 4.4 us of minimum wait at G = 4).
// fence = false is for an arrival that only publishes what EARLIER kernels of the stream wrote: their stores
// reached L2 -- where the peers read them -- when those kernels completed, and this thread has written nothing.
__device__ __forceinline__ void st_relaxed_sys(int32_t* p, int v) {
  asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void peer_signal(int32_t* const* flags_peer, int G, int me, int epoch, bool fence) {
  if (fence) __threadfence_system();
  for (int r = 0; r < G; ++r) st_relaxed_sys(flags_peer[r] + me, epoch);
}

__device__ __forceinline__ int ld_relaxed_sys(const int32_t* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Called by every thread of a CTA; threads r < G wait for rank r's arrival.  The polls are relaxed system-scope
// loads; each waiting thread then re-reads its flag with ld.acquire.sys, which (with the __syncthreads) orders every
// later read of the CTA -- peer rows, published lists -- after the arrivals it has observed.
__device__ __forceinline__ void peer_wait(int32_t* mine, int G, int epoch) {
  if ((int)threadIdx.x < G && ld_relaxed_sys(mine + kPeerErrWord) == 0) {
    const unsigned long long t0 = global_ns();
    while (ld_relaxed_sys(mine + threadIdx.x) < epoch) {
      if (global_ns() - t0 > kPeerTimeoutNs) {
        st_release_sys(mine + kPeerErrWord, epoch);
        break;
      }
    }
    // acquire: one more load of the flag, now with acquire semantics (unlike a fence it does not have to drain this
    // thread's own stores first)
    int seen;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine + threadIdx.x) : "memory");
    (void)seen;
  }
  __syncthreads();
}

// Gather the published (sample, cosine) lists of every rank: CTA (b, r) takes pairs [b*kPullPairs, ...) of
// rank r's list, scatters c into the local global-batch array and writes one (sum c, sum c^2) partial -- in
// list order, so every rank computes bit-identical batch statistics.
constexpr int kPullThreads = 128, kPullPer = 8, kPullPairs = kPullThreads * kPullPer;
constexpr int kPeerCountWord = 16;     // flags[16 + r]: length of rank r's published list of the current step
struct PeerPullArgs {
  const float2* pub_peer[kPeerMaxRanks];   // every rank's published list: (.x = sample position as int bits, .y = c)
  int32_t* flags_peer[kPeerMaxRanks];
  float* c_all;                            // local (G*B)
  double* fwd_part;                        // [G * gridDim.x] x 2
  int G, me, epoch, ng;
  unsigned long long* log;
};
__device__ __forceinline__ float2 ld2_sys(const float2* p) {
  float2 r;
  asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
  return r;
}
__global__ void __launch_bounds__(kPullThreads) peer_pull_kernel(PeerPullArgs a) {
  __shared__ double red[2][kPullThreads / 32];
  const int r = blockIdx.y;
  const bool first = blockIdx.x == 0 && r == 0 && threadIdx.x == 0;
  if (a.log && first) a.log[3] = global_ns();
  peer_wait(a.flags_peer[a.me], a.G, a.epoch);
  if (a.log && first) a.log[4] = global_ns();
  const int cnt = ld_relaxed_sys(a.flags_peer[a.me] + kPeerCountWord + r);
  const float2* pub = a.pub_peer[r];
  const int i0 = blockIdx.x * kPullPairs + threadIdx.x;
  float2 v[kPullPer];
#pragma unroll
  for (int k = 0; k < kPullPer; ++k) {
    const int i = i0 + k * kPullThreads;
    v[k] = (i < cnt) ? ld2_sys(pub + i) : make_float2(__int_as_float(-1), 0.f);
  }
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int k = 0; k < kPullPer; ++k) {
    const int j = __float_as_int(v[k].x);
    if (j >= 0 && j < a.ng) {
      a.c_all[j] = v[k].y;
      const double x = (double)v[k].y;
      a0 += x;
      a1 += x * x;
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a0;
    red[1][threadIdx.x >> 5] = a1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b0 = 0.0, b1 = 0.0;
#pragma unroll
    for (int w = 0; w < kPullThreads / 32; ++w) {
      b0 += red[0][w];
      b1 += red[1][w];
    }
    const int slot = r * gridDim.x + blockIdx.x;
    a.fwd_part[2 * slot] = b0;
    a.fwd_part[2 * slot + 1] = b1;
    if (a.log && blockIdx.x == gridDim.x - 1 && r == a.G - 1) a.log[5] = global_ns();
  }
}

// ---------------------------------------------------------------------------------------------
// per chunk: which samples of the global batch touch my rows?  One CTA per (step, table).
struct PeerSelArgs {
  const int32_t* iu_all;   // [G][rank_stride] global user ids of the chunk, rank-major
  const int32_t* ia_all;
  const float* lab_all;
  int64_t rank_stride;
  int64_t n_local;         // samples of ONE rank in the chunk
  int B, G, me, cap;
  int32_t* key[2];         // [slot][cap] local row
  int32_t* samp[2];        // [slot][cap] position in the global batch (rank r's sample i at r*n + i)
  int32_t* oth[2];         // [slot][cap] global row of the other table
  int32_t* cnt[2];         // [slot]
  int32_t* max_count;      // [2]
  float* lab_step;         // [slot][G*B]
  float* lab_list[2];      // [slot][cap] label of the listed sample (may be null)
};
constexpr int kSelThreads = 1024;

__global__ void __launch_bounds__(kSelThreads, 1) peer_select_kernel(PeerSelArgs a) {
  __shared__ int warp_tot[kSelThreads / 32];
  __shared__ int tile_tot;
  const int slot = blockIdx.x, T = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t first = (int64_t)slot * a.B;
  const int n = (int)max((int64_t)0, min((int64_t)a.B, a.n_local - first));
  const int ng = n * a.G;
  int32_t* key = a.key[T] + (int64_t)slot * a.cap;
  int32_t* samp = a.samp[T] + (int64_t)slot * a.cap;
  int32_t* oth = a.oth[T] + (int64_t)slot * a.cap;
  float* lab = a.lab_step + (int64_t)slot * a.G * a.B;
  float* labl = a.lab_list[T] ? a.lab_list[T] + (int64_t)slot * a.cap : nullptr;
  int running = 0;
  for (int base = 0; base < ng; base += kSelThreads) {
    const int j = base + tid;
    bool mine = false;
    int k_own = 0, k_oth = 0;
    float lab_j = 0.f;
    if (j < ng) {
      const int r = j / n, i = j - r * n;
      const int64_t src = (int64_t)r * a.rank_stride + first + i;
      const int ku = a.iu_all[src], ka = a.ia_all[src];
      k_own = T ? ka : ku;
      k_oth = T ? ku : ka;
      mine = (k_own % a.G) == a.me;
      lab_j = a.lab_all[src];
      if (T == 0) lab[j] = lab_j;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
      const int w = warp_tot[lane];
      int incl = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      warp_tot[lane] = incl - w;
      if (lane == 31) tile_tot = incl;
    }
    __syncthreads();
    if (mine) {
      const int pos = running + warp_tot[wid] + __popc(bal & ((1u << lane) - 1u));
      if (pos < a.cap) {
        key[pos] = k_own / a.G;
        samp[pos] = j;
        oth[pos] = k_oth;
        if (labl) labl[pos] = lab_j;
      }
    }
    running += tile_tot;
    __syncthreads();
  }
  if (tid == 0) {
    a.cnt[T][slot] = min(running, a.cap);
    atomicMax(a.max_count + T, running);  // > cap = overflow, the host refuses to run the chunk
  }
}

// ---------------------------------------------------------------------------------------------
// per step: forward of the listed samples
struct PeerFwdArgs {
  const float* W_peer[2][kPeerMaxRanks];  // [table][rank] shard bases (own entry = local pointer)
  float2* pub;                            // my published (sample position, c) list of this step (local, peer-read)
  const int32_t* key[2];
  const int32_t* samp[2];
  const int32_t* oth[2];
  const int32_t* cnt[2];
  float* stash[2];   // [cap][dim] normalised row of the OTHER table per listed sample
  float* rinv[2];    // [cap] 1/||my row||
  int G, me, dim, blocks0;
  int32_t* flags_peer[kPeerMaxRanks];  // every rank's flag words
  int epoch_rows;    // wait for this epoch (all owners' rows current) before touching peer rows
  int epoch_c;       // the last CTA to finish signals this epoch (my cosines are in every c_all)
  unsigned long long* log;  // AR_PEER_LOG stamps of this step, or null
  int dbg;           // AR_PEER_DBG (timing experiments, WRONG results): 2 = read the other row from my own shard
                     // instead of the owner's
};
constexpr int kPeerTicketWord = 40;    // flags[40]: arrival counter of peer_fwd_kernel's CTAs (local)
constexpr int kPeerSignalWord = 41;    // flags[41]: last "rows current" epoch this rank has announced (local)

template <int NV>
__global__ void __launch_bounds__(kRowThreads) peer_fwd_kernel(PeerFwdArgs a) {
  const bool second = (int)blockIdx.x >= a.blocks0;
  const int blk = second ? blockIdx.x - a.blocks0 : blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int k = blk * kRowWarps + (threadIdx.x >> 5);
  const int cnt = second ? a.cnt[1][0] : a.cnt[0][0];
  int32_t* my_flags = a.flags_peer[a.me];
  if (a.log && blockIdx.x == 0 && threadIdx.x == 0) a.log[0] = global_ns();
  if (threadIdx.x == 0) {  // whichever CTA starts first announces that my rows are current
    if (atomicMax(my_flags + kPeerSignalWord, a.epoch_rows) < a.epoch_rows)
      peer_signal(a.flags_peer, a.G, a.me, a.epoch_rows, false);
  }
  if (blk * kRowWarps < cnt) peer_wait(my_flags, a.G, a.epoch_rows);  // CTA-uniform
  if (a.log && blockIdx.x == 0 && threadIdx.x == 0) a.log[1] = global_ns();
  if (k < cnt) {
  const int32_t* keyp = second ? a.key[1] : a.key[0];
  const int32_t* othp = second ? a.oth[1] : a.oth[0];
  const int dim = a.dim, d4 = dim >> 2;
  const int row = __ldg(keyp + k);
  const int og = __ldg(othp + k);
  const int owner = og % a.G, olocal = og / a.G;
  const float* own_base = second ? a.W_peer[1][a.me] : a.W_peer[0][a.me];
  const float* orow = (second ? a.W_peer[0][(a.dbg & 2) ? a.me : owner] : a.W_peer[1][(a.dbg & 2) ? a.me : owner]) + (size_t)olocal * dim;
  RowTile<NV> w, o;
  w.load(own_base + (size_t)row * dim, d4, lane);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int j = lane + 32 * v;
    o.x[v] = (j < d4) ? ld4_sys(orow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float sw = tile_dot<NV>(w, w);
  const float so = tile_dot<NV>(o, o);
  const float r_w = 1.0f / sqrtf(fmaxf(sw, kL2NormEps));   // same formulas as embed_fwd_kernel
  const float r_o = 1.0f / sqrtf(fmaxf(so, kL2NormEps));
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    w.x[v] = scale4(w.x[v], r_w);
    o.x[v] = scale4(o.x[v], r_o);
  }
  float* stash = second ? a.stash[1] : a.stash[0];
  o.store(stash + (size_t)k * dim, d4, lane);
  float* rinv = second ? a.rinv[1] : a.rinv[0];
  if (lane == 0) rinv[k] = r_w;
  if (!second) {  // the user side publishes the cosine (the anime side would compute the same bits)
    // identical summation on both sides is not needed for that: only this value is ever used
    const float cs = tile_dot<NV>(w, o);
    const int j = __ldg(a.samp[0] + k);
    if (lane == 0) a.pub[k] = make_float2(__int_as_float(j), cs);
  }
  }
  // arrival: once every CTA's stores are fenced, the last one tells every rank (threadFenceReduction pattern)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // device scope is enough: the last CTA's release below is cumulative over the tickets
    const int old = atomicAdd(my_flags + kPeerTicketWord, 1);
    if (old == (int)gridDim.x - 1) {
      my_flags[kPeerTicketWord] = 0;
      const int n_pub = a.cnt[0][0];  // length of my published list, into every rank's array ahead of the flag
      for (int r = 0; r < a.G; ++r) st_relaxed_sys(a.flags_peer[r] + kPeerCountWord + a.me, n_pub);
      peer_signal(a.flags_peer, a.G, a.me, a.epoch_c, true);
      if (a.log) a.log[2] = global_ns();
    }
  }
}

static int check_peer(const ar_peer_ctx* h) {
  AR_REQUIRE(h, "ar_train_steps_peer: null peer ctx");
  AR_REQUIRE(h->n_ranks >= 1 && h->n_ranks <= kPeerMaxRanks && h->rank >= 0 && h->rank < h->n_ranks,
             "ar_train_steps_peer: rank %d / n_ranks %d unsupported (1..%d)", h->rank, h->n_ranks, kPeerMaxRanks);
  for (int r = 0; r < h->n_ranks; ++r)
    AR_REQUIRE(h->W_peer[0][r] && h->W_peer[1][r] && h->pub_peer[r] && h->flags_peer[r],
               "ar_train_steps_peer: peer pointers of rank %d missing", r);
  for (int t = 0; t < 2; ++t)
    AR_REQUIRE(h->sel_key[t] && h->sel_samp[t] && h->sel_oth[t] && h->sel_cnt[t], "ar_train_steps_peer: null selection list");
  AR_REQUIRE(h->sel_cap > 0 && h->sel_cap <= AR_MAX_BATCH, "ar_train_steps_peer: sel_cap %d outside (0, %d]", h->sel_cap, AR_MAX_BATCH);
  AR_REQUIRE(h->max_count && h->label_step && h->c_all && h->dy_all && h->fwd_part_all && h->head_part_all,
             "ar_train_steps_peer: null buffer in peer ctx");
  return AR_OK;
}

// cudaIpc handles opened by this process (opening one twice is an error): handle bytes -> mapped base
struct PeerMapping {
  cudaIpcMemHandle_t h;
  void* base;
};
static std::vector<PeerMapping>& peer_mappings() {
  static std::vector<PeerMapping> v;
  return v;
}

}  // namespace ar

// ---- peer memory plumbing: export a device allocation to the other ranks' processes / map theirs ----
extern "C" int ar_peer_export(const void* dev_ptr, void* handle_out_host, int64_t* offset_out) {
  AR_REQUIRE(dev_ptr && handle_out_host && offset_out, "ar_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == AR_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  // the handle names the whole cudaMalloc allocation; find its base to report dev_ptr's offset in it
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  range_fn get_range = nullptr;
  cudaDriverEntryPointQueryResult qr;
  AR_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", (void**)&get_range, cudaEnableDefault, &qr));
  AR_REQUIRE(get_range && qr == cudaDriverEntryPointSuccess, "ar_peer_export: cuMemGetAddressRange unavailable");
  unsigned long long base = 0;
  size_t size = 0;
  const int drc = get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
  AR_REQUIRE(drc == 0 && base, "ar_peer_export: cuMemGetAddressRange failed (%d)", drc);
  cudaIpcMemHandle_t h;
  AR_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  memcpy(handle_out_host, &h, sizeof(h));
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return AR_OK;
}

extern "C" int ar_peer_open(const void* handle_host, int64_t offset, void** ptr_out) {
  AR_REQUIRE(handle_host && ptr_out && offset >= 0, "ar_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  for (const ar::PeerMapping& m : ar::peer_mappings())
    if (memcmp(&m.h, &h, sizeof(h)) == 0) {
      *ptr_out = (char*)m.base + offset;
      return AR_OK;
    }
  void* base = nullptr;
  AR_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  ar::peer_mappings().push_back({h, base});
  *ptr_out = (char*)base + offset;
  return AR_OK;
}

extern "C" int ar_peer_close_all(void) {
  for (const ar::PeerMapping& m : ar::peer_mappings()) cudaIpcCloseMemHandle(m.base);
  ar::peer_mappings().clear();
  return AR_OK;
}

extern "C" int ar_peer_plan(const int32_t* iu_all, const int32_t* ia_all, const float* label_all, int64_t rank_stride,
                            int64_t n_local, int32_t batch, int32_t n_steps, const ar_plan* plan_u,
                            const ar_plan* plan_a, const ar_peer_ctx* h, void* stream) {
  using namespace ar;
  AR_REQUIRE(iu_all && ia_all && label_all && plan_u && plan_a, "ar_peer_plan: null pointer");
  int rc = check_peer(h);
  if (rc) return rc;
  AR_REQUIRE(batch > 0 && n_local >= 0 && n_local <= rank_stride, "ar_peer_plan: bad sizes");
  AR_REQUIRE(n_steps >= 0 && n_steps <= plan_u->n_slots && n_steps <= plan_a->n_slots, "ar_peer_plan: n_steps exceeds the plans");
  AR_REQUIRE(plan_u->batch_cap == h->sel_cap && plan_a->batch_cap == h->sel_cap, "ar_peer_plan: plans must be sized sel_cap");
  cudaStream_t st = (cudaStream_t)stream;
  AR_CUDA(cudaMemsetAsync(h->max_count, 0, 2 * sizeof(int32_t), st));
  if (n_steps == 0) return AR_OK;
  PeerSelArgs a{};
  a.iu_all = iu_all; a.ia_all = ia_all; a.lab_all = label_all;
  a.rank_stride = rank_stride; a.n_local = n_local;
  a.B = batch; a.G = h->n_ranks; a.me = h->rank; a.cap = h->sel_cap;
  for (int t = 0; t < 2; ++t) {
    a.key[t] = h->sel_key[t]; a.samp[t] = h->sel_samp[t]; a.oth[t] = h->sel_oth[t]; a.cnt[t] = h->sel_cnt[t];
  }
  a.max_count = h->max_count;
  a.lab_step = h->label_step;
  a.lab_list[0] = h->sel_lab[0];
  a.lab_list[1] = h->sel_lab[1];
  peer_select_kernel<<<dim3(n_steps, 2), kSelThreads, 0, st>>>(a);
  AR_LAUNCH_CHECK();
  const ar_plan* plans[2] = {plan_u, plan_a};
  for (int t = 0; t < 2; ++t) {
    if ((rc = ar_plan_build_lists(h->sel_key[t], h->sel_cap, h->sel_cnt[t], n_steps, plans[t], stream))) return rc;
    if (plans[t]->in_prev && (rc = ar_plan_link(plans[t], n_steps, nullptr, nullptr, 1, stream))) return rc;
  }
  return AR_OK;
}

extern "C" int ar_train_steps_peer(const ar_train_ctx* ctx, const ar_peer_ctx* h, int64_t epoch_step0, int32_t slot0,
                                   int64_t t0, int32_t n_steps, int32_t count_hint, void* stream) {
  using namespace ar;
  int rc = check_ctx(ctx, slot0, n_steps);
  if (rc) return rc;
  if ((rc = check_peer(h))) return rc;
  const ar_train_ctx& x = *ctx;
  const int cap = h->sel_cap;
  AR_REQUIRE(x.plan_u.batch_cap == cap && x.plan_a.batch_cap == cap, "ar_train_steps_peer: plans must be sized sel_cap");
  AR_REQUIRE(count_hint > 0 && count_hint <= cap, "ar_train_steps_peer: count_hint %d outside (0, sel_cap]", count_hint);
  AR_REQUIRE(t0 + n_steps < (1ll << 29), "ar_train_steps_peer: optimizer step too large for the barrier epochs");
  cudaStream_t st = (cudaStream_t)stream;
  const int dim = x.users.dim, G = h->n_ranks, B = x.batch;
  // AR_ADAM_REPLAY with a replay schedule and the peer views of the readiness flags: the persistent kernel of the
  // single-GPU path, extended over peer memory (chunk.inl) -- one launch for the whole chunk
  static const bool staged = getenv("AR_PEER_STAGED") != nullptr;
  if (!staged && x.mode == AR_ADAM_REPLAY && x.sched.codes && x.chunk_ws && x.health && h->hdrin_peer[h->rank] &&
      h->sel_lab[0] && h->sel_lab[1]) {
    AR_REQUIRE(slot0 == 0, "ar_train_steps_peer: slot0 must be 0 (the step kernel indexes the plans by step)");
    if ((rc = check_ctx_single(ctx, n_steps))) return rc;
    AR_REQUIRE(t0 + n_steps < (1ll << 27), "ar_train_steps_peer: optimizer step too large for the row flags");
    int live = 0;
    for (int s = 0; s < n_steps; ++s)
      if ((epoch_step0 + s) * (int64_t)B < x.n_samples) live = s + 1;
    if (live == 0) return AR_OK;
    PeerExt px{};
    for (int r = 0; r < G; ++r) {
      AR_REQUIRE(h->rowflag_peer[0][r] && h->rowflag_peer[1][r] && h->pairs_peer[r] && h->hdrin_peer[r],
                 "ar_train_steps_peer: inboxes / row flags of rank %d missing", r);
      for (int k = 0; k < 2; ++k) {
        px.W_peer[k][r] = h->W_peer[k][r];
        px.rowflag_peer[k][r] = h->rowflag_peer[k][r];
      }
      px.pairs_peer[r] = reinterpret_cast<unsigned long long*>(h->pairs_peer[r]);
      px.hdrin_peer[r] = reinterpret_cast<unsigned long long*>(h->hdrin_peer[r]);
    }
    px.flags = h->flags_peer[h->rank];
    static const bool strict = getenv("AR_PEER_STRICT") != nullptr;
    px.strict = strict ? 1 : 0;
    for (int k = 0; k < 2; ++k) {
      px.key[k] = h->sel_key[k];
      px.oth[k] = h->sel_oth[k];
      px.cnt[k] = h->sel_cnt[k];
      px.lab[k] = h->sel_lab[k];
    }
    px.samp0 = h->sel_samp[0];
    px.label_step = h->label_step;
    px.G = G; px.me = h->rank; px.cap = cap; px.gb = G * B;
    return launch_chunk(x, epoch_step0, t0, live, st, &px);
  }
  static const bool no_overlap = getenv("AR_NO_LOOKAHEAD") != nullptr;
  // where the look-ahead catch-up of step s+1 starts: behind the forward of step s (default) -- the forward is
  // what the OTHER ranks wait for, so it gets the SMs to itself and the replay overlaps the pull, the head and
  // the row update -- or right behind update(s-1) as on one GPU (AR_PEER_AHEAD_EARLY=1).  Started early, the
  // low-priority replay CTAs fill the SMs as the forward drains and the next kernel of the main stream waits
  // for slots: a 13-18 us gap between forward and pull in the in-kernel timeline.  Measured on 2 GPUs with the
  // plan-order schedule PeerTrainSession uses (no classify pass: sharded replays are short and even): late
  // 72.0 us/step, early 80.5; with the longest-first schedule 78.5 / 78.3.
  static const bool ahead_early = getenv("AR_PEER_AHEAD_EARLY") != nullptr;
  const bool can_ahead = x.plan_u.in_prev && x.plan_a.in_prev;
  Lookahead* la = (x.mode == AR_ADAM_REPLAY && can_ahead && !no_overlap) ? lookahead() : nullptr;
  if (la) AR_CUDA(cudaEventRecord(la->ev_upd[1], st));
  float* c_all = h->c_all;
  static const bool do_log = getenv("AR_PEER_LOG") != nullptr;
  static const int dbg = getenv("AR_PEER_DBG") ? atoi(getenv("AR_PEER_DBG")) : 0;
  static unsigned long long* log_dev = nullptr;
  constexpr int kLogSteps = 1024;
  if (do_log && !log_dev) {
    AR_CUDA(cudaMalloc((void**)&log_dev, (size_t)kLogSteps * kLogStamps * 8));
    AR_CUDA(cudaMemset(log_dev, 0, (size_t)kLogSteps * kLogStamps * 8));
  }
  for (int s = 0; s < n_steps; ++s) {
    const int64_t e = epoch_step0 + s;
    const int64_t base = e * (int64_t)B;
    if (base >= x.n_samples) break;
    const int n = (int)std::min<int64_t>(B, x.n_samples - base);  // identical on every rank (caller's contract)
    const int slot = slot0 + s;
    const int64_t t = t0 + s + 1;
    const int ng = n * G;
    unsigned long long* log = (do_log && s < kLogSteps) ? log_dev + (size_t)s * kLogStamps : nullptr;
    const bool has_next = (s + 1 < n_steps) && ((e + 1) * (int64_t)B < x.n_samples);
    if (x.mode == AR_ADAM_REPLAY && (!la || s == 0)) {
      if ((rc = launch_catchup(&x.users, &x.plan_u, slot, &x.anime, &x.plan_a, slot, x.alpha, x.l2, t - 1, st, false, x.sched_ws, reg_of(x)))) return rc;
    }
    bool ahead = false;
    auto launch_ahead = [&](cudaEvent_t after) -> int {
      // as in run_steps: my rows of step s+1 that step s leaves alone are brought to step t on the side stream
      AR_CUDA(cudaStreamWaitEvent(la->st2, after, 0));
      int32_t* ws2 = x.sched_ws ? x.sched_ws + 3 * ((size_t)x.plan_u.batch_cap + x.plan_a.batch_cap) + 4 : nullptr;
      int r2 = launch_catchup(&x.users, &x.plan_u, slot + 1, &x.anime, &x.plan_a, slot + 1, x.alpha, x.l2, t, la->st2, true, ws2, reg_of(x));
      if (r2) return r2;
      AR_CUDA(cudaEventRecord(la->ev_ahead, la->st2));
      ahead = true;
      return AR_OK;
    };
    if (la && has_next && ahead_early && (rc = launch_ahead(la->ev_upd[(s + 1) & 1]))) return rc;
    PeerFwdArgs f{};
    for (int r = 0; r < G; ++r) {
      f.W_peer[0][r] = h->W_peer[0][r];
      f.W_peer[1][r] = h->W_peer[1][r];
      f.flags_peer[r] = h->flags_peer[r];
    }
    for (int k = 0; k < 2; ++k) {
      f.key[k] = h->sel_key[k] + (int64_t)slot * cap;
      f.samp[k] = h->sel_samp[k] + (int64_t)slot * cap;
      f.oth[k] = h->sel_oth[k] + (int64_t)slot * cap;
      f.cnt[k] = h->sel_cnt[k] + slot;
    }
    f.pub = reinterpret_cast<float2*>(h->pub_peer[h->rank]);
    f.stash[0] = x.ah; f.stash[1] = x.uh;
    f.rinv[0] = x.ru; f.rinv[1] = x.ra;
    f.G = G; f.me = h->rank; f.dim = dim;
    f.epoch_rows = (int)(2 * t);      // signalled by this kernel's first CTA: my rows of this step are current
    f.epoch_c = (int)(2 * t + 1);     // signalled by its last CTA: my cosines are in every c_all
    f.dbg = dbg;
    f.log = log;
    f.blocks0 = ceil_div(count_hint, kRowWarps);
    AR_DISPATCH_NV(dim, peer_fwd_kernel<NV><<<2 * f.blocks0, kRowThreads, 0, st>>>(f));
    AR_LAUNCH_CHECK();
    if (la && has_next && !ahead_early) {
      AR_CUDA(cudaEventRecord(la->ev_mid, st));
      if ((rc = launch_ahead(la->ev_mid))) return rc;
    }
    PeerPullArgs pl{};
    for (int r = 0; r < G; ++r) {
      pl.pub_peer[r] = reinterpret_cast<const float2*>(h->pub_peer[r]);
      pl.flags_peer[r] = h->flags_peer[r];
    }
    pl.c_all = c_all;
    pl.fwd_part = h->fwd_part_all;
    pl.G = G; pl.me = h->rank; pl.epoch = (int)(2 * t + 1); pl.ng = ng;
    pl.log = log;
    const int pull_blocks = ceil_div(count_hint, kPullPairs);
    const int nfp = G * pull_blocks;
    peer_pull_kernel<<<dim3(pull_blocks, G), kPullThreads, 0, st>>>(pl);
    AR_LAUNCH_CHECK();
    head_step_kernel<<<ceil_div(ng, kHeadThreads), kHeadThreads, 0, st>>>(
        c_all, h->label_step + (int64_t)slot * G * B, ng, nullptr, h->fwd_part_all,
        HeadIO{x.head, x.head_m, x.head_v, x.bn_moving, x.alpha, h->dy_all, x.stepc, x.ticket, x.metrics}, t, h->head_part_all, nfp);
    AR_LAUNCH_CHECK();
    UpdateArgs a{};
    fill_update(a, 0, &x.users, &x.plan_u, slot, x.ah, x.ru, count_hint);
    fill_update(a, 1, &x.anime, &x.plan_a, slot, x.uh, x.ra, count_hint);
    a.samp[0] = f.samp[0];
    a.samp[1] = f.samp[1];
    const RegAcc ss = reg_of(x);
    if ((rc = launch_update(a, true, c_all, h->dy_all, x.stepc, x.alpha, x.l2, t, 0, ss, st))) return rc;
    if (la) {
      AR_CUDA(cudaEventRecord(la->ev_upd[s & 1], st));
      if (ahead) AR_CUDA(cudaStreamWaitEvent(st, la->ev_ahead, 0));
    }
    if (x.mode == AR_ADAM_DENSE) {
      if ((rc = launch_flush(&x.users, x.alpha, x.l2, t, ss, st))) return rc;
      if ((rc = launch_flush(&x.anime, x.alpha, x.l2, t, ss, st))) return rc;
    }
  }
  if (do_log && n_steps >= 8) {
    AR_CUDA(cudaStreamSynchronize(st));
    const int ns = std::min(n_steps, kLogSteps);
    std::vector<unsigned long long> hl((size_t)ns * kLogStamps);
    AR_CUDA(cudaMemcpy(hl.data(), log_dev, hl.size() * 8, cudaMemcpyDeviceToHost));
    double acc[6] = {0, 0, 0, 0, 0, 0};
    int cntd = 0;
    for (int i = 2; i + 1 < ns; ++i) {
      const unsigned long long* a = &hl[(size_t)i * kLogStamps];
      const unsigned long long* nx = &hl[(size_t)(i + 1) * kLogStamps];
      if (!a[0] || !a[5] || !nx[0]) continue;
      for (int k = 0; k < 5; ++k) acc[k] += (double)(long long)(a[k + 1] - a[k]);
      acc[5] += (double)(long long)(nx[0] - a[5]);
      ++cntd;
    }
    if (const char* path = getenv("AR_PEER_LOG_DUMP")) {  // raw stamps of this call, one file per rank
      char name[512];
      snprintf(name, sizeof(name), "%s.rank%d.bin", path, h->rank);
      if (FILE* fp = fopen(name, "wb")) {
        fwrite(hl.data(), 8, hl.size(), fp);
        fclose(fp);
      }
    }
    if (cntd)
      fprintf(stderr, "[peer log rank %d] us over %d steps: fwd wait %.1f | fwd body %.1f | ->pull %.1f | pull wait %.1f | pull %.1f | head+update+gaps %.1f\n",
              h->rank, cntd, acc[0] / cntd / 1e3, acc[1] / cntd / 1e3, acc[2] / cntd / 1e3, acc[3] / cntd / 1e3,
              acc[4] / cntd / 1e3, acc[5] / cntd / 1e3);
  }
  return AR_OK;
}
