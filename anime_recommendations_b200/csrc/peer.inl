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
//   [flag barrier 2t]     every owner's rows are current
//   peer_fwd   warp per listed sample: my row (local) + the other table's row (NVLink) -> 1/||.||, cosine c;
//              the normalised other row is stashed locally for the gradient; the user side publishes c to
//              every rank's c_all (G 4-byte peer stores)
//   [flag barrier 2t+1]   c_all complete everywhere
//   head_step  over the global batch, redundantly on every rank (SyncBN; replicas of the 4 head scalars
//              stay bit-identical)
//   rows_update  the single-GPU kernel on my rows: segment sums over the stash, catch-up-free Adam
// NVLink traffic per step and GPU: the pulled rows, ~2*B*dim*4 bytes (10 MB at B = 10000, dim = 128), and
// G*B cosines out.  Nothing is staged, packed or merged, and a row's gradient is summed by one warp in plan
// order -- the result does not depend on which rank a sample came from.
//
// Flag barriers: rank r's arrival is the epoch number stored (st.release.sys) into word r of EVERY rank's flag
// array; a waiter spins (ld.acquire.sys) on its own array.  Epochs are derived from the optimizer step, so
// they only grow; a rank that waits longer than kPeerTimeoutNs raises a sticky error word instead of hanging
// the GPU.  ar_peer_barrier is the stand-alone form (one 32-thread kernel); the step uses the split form.
namespace ar {

constexpr int kPeerMaxRanks = AR_PEER_MAX_RANKS;
constexpr int kPeerErrWord = 32;                       // flags[32]: sticky "a barrier timed out" (epoch that did)
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void st_release_sys(int32_t* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
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

struct PeerFlags {
  int32_t* peer[kPeerMaxRanks];
  int G, me;
};

__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerFlags f, int epoch) {
  const int r = threadIdx.x;
  if (r >= f.G) return;
  int32_t* mine = f.peer[f.me];
  if (ld_acquire_sys(mine + kPeerErrWord) != 0) return;  // a barrier already timed out: do not stall again
  __threadfence_system();
  st_release_sys(f.peer[r] + f.me, epoch);
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(mine + r) < epoch) {
    if (global_ns() - t0 > kPeerTimeoutNs) {
      st_release_sys(mine + kPeerErrWord, epoch);
      break;
    }
  }
}

// The training step splits the barrier: arrival is signalled by a one-warp kernel (epoch 2t, "my rows are
// current") or by the last CTA of the forward (epoch 2t+1, "my cosines are published"), and the FIRST kernel
// that needs the other ranks' data waits in its prologue -- its CTAs are already resident when the flags land,
// which takes a launch gap and a separate spinning kernel off the critical path.
__global__ void __launch_bounds__(32) peer_signal_kernel(PeerFlags f, int epoch) {
  const int r = threadIdx.x;
  if (r < f.G) st_release_sys(f.peer[r] + f.me, epoch);
}

// called by every thread of a CTA; threads r < G wait for rank r's arrival
__device__ __forceinline__ void peer_wait(int32_t* mine, int G, int epoch) {
  if ((int)threadIdx.x < G && ld_acquire_sys(mine + kPeerErrWord) == 0) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {
      if (global_ns() - t0 > kPeerTimeoutNs) {
        st_release_sys(mine + kPeerErrWord, epoch);
        break;
      }
    }
  }
  __syncthreads();
}

// (sum c, sum c^2) per kPartSamples cosines, behind the "cosines published" flags.  One partial per CTA (the
// single-GPU forward emits one per 8 samples; at G*B samples every head CTA would re-read G times as many)
constexpr int kPartThreads = 128, kPartSamples = kPartThreads * 8;
__global__ void __launch_bounds__(kPartThreads)
c_partials_wait_kernel(const float* __restrict__ c, int n, double* __restrict__ fwd_part, int32_t* flags, int G, int epoch) {
  __shared__ double red[2][kPartThreads / 32];
  peer_wait(flags, G, epoch);
  const int i0 = (blockIdx.x * kPartThreads + threadIdx.x) * 8;
  double a0 = 0.0, a1 = 0.0;
  for (int i = i0; i < min(n, i0 + 8); ++i) {
    const double x = (double)__ldcg(c + i);
    a0 += x;
    a1 += x * x;
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
    for (int w = 0; w < kPartThreads / 32; ++w) {
      b0 += red[0][w];
      b1 += red[1][w];
    }
    fwd_part[2 * blockIdx.x] = b0;
    fwd_part[2 * blockIdx.x + 1] = b1;
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
  int running = 0;
  for (int base = 0; base < ng; base += kSelThreads) {
    const int j = base + tid;
    bool mine = false;
    int k_own = 0, k_oth = 0;
    if (j < ng) {
      const int r = j / n, i = j - r * n;
      const int64_t src = (int64_t)r * a.rank_stride + first + i;
      const int ku = a.iu_all[src], ka = a.ia_all[src];
      k_own = T ? ka : ku;
      k_oth = T ? ku : ka;
      mine = (k_own % a.G) == a.me;
      if (T == 0) lab[j] = a.lab_all[src];
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
  float* c_peer[kPeerMaxRanks];           // every rank's c_all
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
};
constexpr int kPeerTicketWord = 40;    // flags[40]: arrival counter of peer_fwd_kernel's CTAs (local)

template <int NV>
__global__ void __launch_bounds__(kRowThreads) peer_fwd_kernel(PeerFwdArgs a) {
  const bool second = (int)blockIdx.x >= a.blocks0;
  const int blk = second ? blockIdx.x - a.blocks0 : blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int k = blk * kRowWarps + (threadIdx.x >> 5);
  const int cnt = second ? a.cnt[1][0] : a.cnt[0][0];
  int32_t* my_flags = a.flags_peer[a.me];
  if (blk * kRowWarps < cnt) peer_wait(my_flags, a.G, a.epoch_rows);  // CTA-uniform
  if (k < cnt) {
  const int32_t* keyp = second ? a.key[1] : a.key[0];
  const int32_t* othp = second ? a.oth[1] : a.oth[0];
  const int dim = a.dim, d4 = dim >> 2;
  const int row = __ldg(keyp + k);
  const int og = __ldg(othp + k);
  const int owner = og % a.G, olocal = og / a.G;
  const float* own_base = second ? a.W_peer[1][a.me] : a.W_peer[0][a.me];
  const float* orow = (second ? a.W_peer[0][owner] : a.W_peer[1][owner]) + (size_t)olocal * dim;
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
    if (lane < a.G) a.c_peer[lane][j] = cs;
  }
  }
  // arrival: once every CTA's peer stores are fenced, the last one tells every rank (threadFenceReduction pattern)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int old = atomicAdd(my_flags + kPeerTicketWord, 1);
    if (old == (int)gridDim.x - 1) {
      my_flags[kPeerTicketWord] = 0;
      for (int r = 0; r < a.G; ++r) st_release_sys(a.flags_peer[r] + a.me, a.epoch_c);
    }
  }
}

static int check_peer(const ar_peer_ctx* h) {
  AR_REQUIRE(h, "ar_train_steps_peer: null peer ctx");
  AR_REQUIRE(h->n_ranks >= 1 && h->n_ranks <= kPeerMaxRanks && h->rank >= 0 && h->rank < h->n_ranks,
             "ar_train_steps_peer: rank %d / n_ranks %d unsupported (1..%d)", h->rank, h->n_ranks, kPeerMaxRanks);
  for (int r = 0; r < h->n_ranks; ++r)
    AR_REQUIRE(h->W_peer[0][r] && h->W_peer[1][r] && h->c_all_peer[r] && h->flags_peer[r],
               "ar_train_steps_peer: peer pointers of rank %d missing", r);
  for (int t = 0; t < 2; ++t)
    AR_REQUIRE(h->sel_key[t] && h->sel_samp[t] && h->sel_oth[t] && h->sel_cnt[t], "ar_train_steps_peer: null selection list");
  AR_REQUIRE(h->sel_cap > 0 && h->sel_cap <= AR_MAX_BATCH, "ar_train_steps_peer: sel_cap %d outside (0, %d]", h->sel_cap, AR_MAX_BATCH);
  AR_REQUIRE(h->max_count && h->label_step && h->dy_all && h->fwd_part_all && h->head_part_all,
             "ar_train_steps_peer: null buffer in peer ctx");
  return AR_OK;
}

static int peer_barrier(const ar_peer_ctx& h, int epoch, cudaStream_t st) {
  PeerFlags f{};
  for (int r = 0; r < h.n_ranks; ++r) f.peer[r] = h.flags_peer[r];
  f.G = h.n_ranks;
  f.me = h.rank;
  peer_barrier_kernel<<<1, 32, 0, st>>>(f, epoch);
  AR_LAUNCH_CHECK();
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

extern "C" int ar_peer_barrier(const ar_peer_ctx* h, int32_t epoch, void* stream) {
  using namespace ar;
  AR_REQUIRE(h, "ar_peer_barrier: null peer ctx");
  for (int r = 0; r < h->n_ranks; ++r) AR_REQUIRE(h->flags_peer[r], "ar_peer_barrier: flags of rank %d missing", r);
  return peer_barrier(*h, epoch, (cudaStream_t)stream);
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
  static const bool no_overlap = getenv("AR_NO_LOOKAHEAD") != nullptr;
  const bool can_ahead = x.plan_u.in_prev && x.plan_a.in_prev;
  Lookahead* la = (x.mode == AR_ADAM_REPLAY && can_ahead && !no_overlap) ? lookahead() : nullptr;
  if (la) AR_CUDA(cudaEventRecord(la->ev_upd[1], st));
  float* c_all = h->c_all_peer[h->rank];
  // AR_PEER_PROFILE=1: events around every launch, per-stage averages to stderr (developer aid; synchronises)
  static const bool prof = getenv("AR_PEER_PROFILE") != nullptr;
  // AR_PEER_UNFUSED=1: stand-alone barrier kernels instead of the split form (A/B measurement)
  static const bool unfused = getenv("AR_PEER_UNFUSED") != nullptr;
  StageTimer tm_store;
  tm_store.st = st;
  StageTimer* timer = prof ? &tm_store : nullptr;
  for (int s = 0; s < n_steps; ++s) {
    const int64_t e = epoch_step0 + s;
    const int64_t base = e * (int64_t)B;
    if (base >= x.n_samples) break;
    const int n = (int)std::min<int64_t>(B, x.n_samples - base);  // identical on every rank (caller's contract)
    const int slot = slot0 + s;
    const int64_t t = t0 + s + 1;
    const int ng = n * G;
    const bool has_next = (s + 1 < n_steps) && ((e + 1) * (int64_t)B < x.n_samples);
    if (x.mode == AR_ADAM_REPLAY && (!la || s == 0)) {
      if ((rc = launch_catchup(&x.users, &x.plan_u, slot, &x.anime, &x.plan_a, slot, x.alpha, x.l2, t - 1, st, false, x.sched_ws))) return rc;
    }
    bool ahead = false;
    if (la && has_next) {  // as in run_steps: my rows of step s+1 that step s leaves alone, on the side stream
      AR_CUDA(cudaStreamWaitEvent(la->st2, la->ev_upd[(s + 1) & 1], 0));
      int32_t* ws2 = x.sched_ws ? x.sched_ws + 3 * ((size_t)x.plan_u.batch_cap + x.plan_a.batch_cap) + 4 : nullptr;
      if ((rc = launch_catchup(&x.users, &x.plan_u, slot + 1, &x.anime, &x.plan_a, slot + 1, x.alpha, x.l2, t, la->st2, true, ws2)))
        return rc;
      AR_CUDA(cudaEventRecord(la->ev_ahead, la->st2));
      ahead = true;
    }
    AR_TICK(0);
    PeerFlags pf{};
    for (int r = 0; r < G; ++r) pf.peer[r] = h->flags_peer[r];
    pf.G = G;
    pf.me = h->rank;
    if (unfused) {
      if ((rc = peer_barrier(*h, (int)(2 * t), st))) return rc;
    } else {
      peer_signal_kernel<<<1, 32, 0, st>>>(pf, (int)(2 * t));  // my rows of this step are current
      AR_LAUNCH_CHECK();
    }
    AR_TICK(1);
    PeerFwdArgs f{};
    for (int r = 0; r < G; ++r) {
      f.W_peer[0][r] = h->W_peer[0][r];
      f.W_peer[1][r] = h->W_peer[1][r];
      f.c_peer[r] = h->c_all_peer[r];
    }
    for (int k = 0; k < 2; ++k) {
      f.key[k] = h->sel_key[k] + (int64_t)slot * cap;
      f.samp[k] = h->sel_samp[k] + (int64_t)slot * cap;
      f.oth[k] = h->sel_oth[k] + (int64_t)slot * cap;
      f.cnt[k] = h->sel_cnt[k] + slot;
    }
    f.stash[0] = x.ah; f.stash[1] = x.uh;
    f.rinv[0] = x.ru; f.rinv[1] = x.ra;
    f.G = G; f.me = h->rank; f.dim = dim;
    for (int r = 0; r < G; ++r) f.flags_peer[r] = h->flags_peer[r];
    f.epoch_rows = unfused ? 0 : (int)(2 * t);
    f.epoch_c = (int)(2 * t + 1);
    f.blocks0 = ceil_div(count_hint, kRowWarps);
    AR_DISPATCH_NV(dim, peer_fwd_kernel<NV><<<2 * f.blocks0, kRowThreads, 0, st>>>(f));
    AR_LAUNCH_CHECK();
    AR_TICK(2);
    if (unfused && (rc = peer_barrier(*h, (int)(2 * t + 1), st))) return rc;
    const int nfp = ceil_div(ng, kPartSamples);
    c_partials_wait_kernel<<<nfp, kPartThreads, 0, st>>>(c_all, ng, h->fwd_part_all, h->flags_peer[h->rank], G,
                                                         unfused ? 0 : (int)(2 * t + 1));
    AR_LAUNCH_CHECK();
    AR_TICK(3);
    head_step_kernel<<<ceil_div(ng, kHeadThreads), kHeadThreads, 0, st>>>(
        c_all, h->label_step + (int64_t)slot * G * B, ng, nullptr, h->fwd_part_all, x.head, x.head_m, x.head_v,
        x.bn_moving, x.alpha, t, h->dy_all, h->head_part_all, x.stepc, x.ticket, x.metrics + t * 4, nfp);
    AR_LAUNCH_CHECK();
    AR_TICK(4);
    UpdateArgs a{};
    fill_update(a, 0, &x.users, &x.plan_u, slot, x.ah, x.ru, count_hint);
    fill_update(a, 1, &x.anime, &x.plan_a, slot, x.uh, x.ra, count_hint);
    a.samp[0] = f.samp[0];
    a.samp[1] = f.samp[1];
    double* ss = (x.mode == AR_ADAM_DENSE && x.reg_sumsq) ? x.reg_sumsq + t * 32 : nullptr;
    if ((rc = launch_update(a, true, c_all, h->dy_all, x.stepc, x.alpha, x.l2, t, 0, ss, st))) return rc;
    AR_TICK(5);
    if (la) {
      AR_CUDA(cudaEventRecord(la->ev_upd[s & 1], st));
      if (ahead) AR_CUDA(cudaStreamWaitEvent(st, la->ev_ahead, 0));
    }
    if (x.mode == AR_ADAM_DENSE) {
      if ((rc = launch_flush(&x.users, x.alpha, x.l2, t, ss, st))) return rc;
      if ((rc = launch_flush(&x.anime, x.alpha, x.l2, t, ss, st))) return rc;
    }
  }
  if (timer && timer->ev.size() >= 12) {
    AR_CUDA(cudaStreamSynchronize(st));
    const size_t ns = timer->ev.size() / 6;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (size_t i = 1; i < ns; ++i) {  // skip the first step (exposed catch-up)
      float ms = 0.f;
      for (int k = 0; k < 5; ++k) {
        cudaEventElapsedTime(&ms, timer->ev[i * 6 + k], timer->ev[i * 6 + k + 1]);
        acc[k] += ms;
      }
      cudaEventElapsedTime(&ms, timer->ev[(i - 1) * 6 + 5], timer->ev[i * 6]);
      acc[5] += ms;
    }
    fprintf(stderr, "[peer rank %d] us/step over %zu steps: signal %.1f  wait+fwd %.1f  wait+c_partials %.1f  head %.1f  update %.1f  between-steps %.1f\n",
            h->rank, ns - 1, 1e3 * acc[0] / (ns - 1), 1e3 * acc[1] / (ns - 1), 1e3 * acc[2] / (ns - 1),
            1e3 * acc[3] / (ns - 1), 1e3 * acc[4] / (ns - 1), 1e3 * acc[5] / (ns - 1));
    for (cudaEvent_t ev : timer->ev) cudaEventDestroy(ev);
  }
  return AR_OK;
}
