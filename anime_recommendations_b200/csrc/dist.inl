// Multi-GPU training, replicated tables (BASELINE cfg2 / SURVEY §8e row 1).  Included by train.cu.
//
// One process per GPU; every rank holds the full tables + Adam state and trains on its own B samples
// of the global batch.  Per step:
//   catch-up(local rows) -> embed_fwd(local) -> [NCCL all-gather of c and labels] -> head_step over
//   the GLOBAL batch (identical on every rank: SyncBN statistics, identical head update, dy for all
//   samples) -> rows_update in EMIT mode: per local distinct row the partial sums (id, q, P[dim]) ->
//   [NCCL all-gather of the packed partial rows] -> rows_merge_update: for every row of the union, the
//   lowest rank holding it ("leader" entry) adds the partials of all ranks in rank order and applies
//   catch-up + Adam.  Every rank runs the same merge over the same gathered bytes in the same order,
//   so the replicas stay bit-identical without ever exchanging weights.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") -- the copy PyTorch already loaded -- so
// libanimerec.so has no link-time NCCL dependency and loads on a CPU-only box.
#include <dlfcn.h>
#include <nccl.h>

namespace ar {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // PyTorch's bundled copy, if loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
    return nullptr;
  }
#define AR_NCCL_SYM(name)                                              \
  api.name = (decltype(api.name))dlsym(h, "nccl" #name);               \
  if (!api.name) {                                                     \
    set_error("NCCL: symbol nccl" #name " not found");                 \
    return nullptr;                                                    \
  }
  AR_NCCL_SYM(GetUniqueId)
  AR_NCCL_SYM(CommInitRank)
  AR_NCCL_SYM(CommDestroy)
  AR_NCCL_SYM(AllGather)
  AR_NCCL_SYM(AllReduce)
  AR_NCCL_SYM(GroupStart)
  AR_NCCL_SYM(GroupEnd)
  AR_NCCL_SYM(Send)
  AR_NCCL_SYM(Recv)
  AR_NCCL_SYM(GetErrorString)
#undef AR_NCCL_SYM
  api.handle = h;
  return &api;
}

#define AR_NCCL(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      ar::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, nccl_api()->GetErrorString(r__)); \
      return AR_ERR_NCCL;                                                                      \
    }                                                                                          \
  } while (0)

// packed partial-row block of one rank: [ids_u B | q_u B | P_u B*D | ids_a B | q_a B | P_a B*D], 4-byte words
__host__ __device__ inline size_t pack_table_words(int B, int D) { return (size_t)B * (D + 2); }

struct MergeArgs {
  ar_table tab[2];
  const float* recv;   // n_ranks packed blocks
  int n_ranks;
  int B;
  int blocks_tab0;     // CTAs for the user table entries
};

// lower_bound in a sorted id list padded with INT_MAX
__device__ __forceinline__ int find_id(const int32_t* __restrict__ ids, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ids + mid) < key) lo = mid + 1; else hi = mid;
  }
  return (lo < n && __ldg(ids + lo) == key) ? lo : -1;
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads)
rows_merge_update_kernel(MergeArgs a, const float* __restrict__ alpha, float l2x2, int64_t t, int replay,
                         RegAcc reg) {
  const bool second = (int)blockIdx.x >= a.blocks_tab0;
  const int blk = second ? blockIdx.x - a.blocks_tab0 : blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int e = blk * kRowWarps + (threadIdx.x >> 5);
  const int B = a.B, G = a.n_ranks;
  if (e >= G * B) return;
  ar_table tb;
  tb.dim = a.tab[0].dim;
  tb.n_rows = second ? a.tab[1].n_rows : a.tab[0].n_rows;
  tb.W = second ? a.tab[1].W : a.tab[0].W;
  tb.m = second ? a.tab[1].m : a.tab[0].m;
  tb.v = second ? a.tab[1].v : a.tab[0].v;
  tb.last_step = second ? a.tab[1].last_step : a.tab[0].last_step;
  const int dim = tb.dim, d4 = dim >> 2;
  const size_t tw = pack_table_words(B, dim);
  const size_t toff = second ? tw : 0;
  const int p = e / B, j = e - p * B;
  const int32_t* my_ids = reinterpret_cast<const int32_t*>(a.recv + (size_t)p * 2 * tw + toff);
  const int id = __ldg(my_ids + j);
  if (id == 0x7fffffff) return;
  // lane r looks the row up in rank r's list (all ranks in parallel)
  int found = -1;
  if (lane < G) {
    if (lane == p) found = j;
    else found = find_id(reinterpret_cast<const int32_t*>(a.recv + (size_t)lane * 2 * tw + toff), B, id);
  }
  const unsigned have = __ballot_sync(0xffffffffu, found >= 0);
  if ((__ffs(have) - 1) != p) return;  // a lower rank leads this row
  RowTile<NV> acc;
  acc.zero();
  float q = 0.f;
  for (unsigned rem = have; rem; rem &= rem - 1) {  // ascending rank order: deterministic
    const int r = __ffs(rem) - 1;
    const int pos = __shfl_sync(0xffffffffu, found, r);
    const float* base = a.recv + (size_t)r * 2 * tw + toff;
    RowTile<NV> part;
    part.load(base + 2 * (size_t)B + (size_t)pos * dim, d4, lane);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc.x[k].x += part.x[k].x; acc.x[k].y += part.x[k].y; acc.x[k].z += part.x[k].z; acc.x[k].w += part.x[k].w;
    }
    q += __ldg(base + B + pos);
  }
  finish_row<NV>(tb, id, acc, q, -1.0f, alpha, l2x2, t, replay, reg, lane);
}

static int check_dist(const ar_train_ctx* ctx, const ar_dist_ctx* d) {
  AR_REQUIRE(d && d->comm, "ar_train_steps_dist: null dist ctx / communicator");
  AR_REQUIRE(d->n_ranks >= 1 && d->n_ranks <= 32 && d->rank >= 0 && d->rank < d->n_ranks,
             "ar_train_steps_dist: rank %d / n_ranks %d unsupported (1..32)", d->rank, d->n_ranks);
  AR_REQUIRE(d->c_all && d->label_all && d->dy_all && d->fwd_part_all && d->head_part_all && d->send && d->recv,
             "ar_train_steps_dist: null buffer in dist ctx");
  AR_REQUIRE((int64_t)ctx->batch * d->n_ranks <= (1 << 24), "ar_train_steps_dist: global batch too large");
  return AR_OK;
}

static int run_steps_dist(const ar_train_ctx& x, const ar_dist_ctx& d, int64_t epoch_step0, int32_t slot0,
                          int64_t t0, int32_t n_steps, cudaStream_t st) {
  NcclApi* nc = nccl_api();
  if (!nc) return AR_ERR_NCCL;
  ncclComm_t comm = (ncclComm_t)d.comm;
  const int dim = x.users.dim, G = d.n_ranks, B = x.batch;
  const size_t tw = pack_table_words(B, dim);
  const float l2x2 = (float)(2.0 * (double)x.l2);
  static const bool no_overlap = getenv("AR_NO_LOOKAHEAD") != nullptr;
  const bool can_ahead = x.plan_u.in_prev && x.plan_a.in_prev;  // flags built from ALL ranks' lists (ar_plan_link)
  Lookahead* la = (x.mode == AR_ADAM_REPLAY && can_ahead && !no_overlap) ? lookahead() : nullptr;
  if (la) AR_CUDA(cudaEventRecord(la->ev_upd[1], st));
  for (int s = 0; s < n_steps; ++s) {
    const int64_t e = epoch_step0 + s;
    const int64_t base = e * (int64_t)B;
    if (base >= x.n_samples) break;
    const int n = (int)std::min<int64_t>(B, x.n_samples - base);  // identical on every rank (caller's contract)
    const int slot = slot0 + s;
    const int64_t t = t0 + s + 1;
    const int ng = n * G;
    int rc;
    const bool has_next = (s + 1 < n_steps) && ((e + 1) * (int64_t)B < x.n_samples);
    if (x.mode == AR_ADAM_REPLAY && (!la || s == 0)) {
      if ((rc = launch_catchup(&x.users, &x.plan_u, slot, &x.anime, &x.plan_a, slot, x.alpha, x.l2, t - 1, st, false, x.sched_ws, reg_of(x)))) return rc;
    }
    bool ahead = false;
    if (la && has_next) {
      // look-ahead (see run_steps): my rows of step s+1 that NO rank touches in step s are brought to step t on
      // the side stream while step s runs; the rest are brought up to date by this step's merge (replay = 1)
      AR_CUDA(cudaStreamWaitEvent(la->st2, la->ev_upd[(s + 1) & 1], 0));
      int32_t* ws2 = x.sched_ws ? x.sched_ws + 3 * ((size_t)x.plan_u.batch_cap + x.plan_a.batch_cap) + 4 : nullptr;
      if ((rc = launch_catchup(&x.users, &x.plan_u, slot + 1, &x.anime, &x.plan_a, slot + 1, x.alpha, x.l2, t, la->st2, true, ws2, reg_of(x))))
        return rc;
      AR_CUDA(cudaEventRecord(la->ev_ahead, la->st2));
      ahead = true;
    }
    AR_DISPATCH_NV(dim, embed_fwd_kernel<NV><<<ceil_div(n, kRowWarps), kRowThreads, 0, st>>>(
                            x.users.W, x.anime.W, dim, x.iu + base, x.ia + base, n, nullptr, x.uh, x.ah, x.c, x.ru, x.ra, nullptr));
    AR_LAUNCH_CHECK();
    // SyncBN: the head sees the whole global batch (rank r's samples at [r*n, (r+1)*n))
    AR_NCCL(nc->GroupStart());
    AR_NCCL(nc->AllGather(x.c, d.c_all, (size_t)n, ncclFloat32, comm, st));
    AR_NCCL(nc->AllGather(x.label + base, d.label_all, (size_t)n, ncclFloat32, comm, st));
    AR_NCCL(nc->GroupEnd());
    c_partials_kernel<<<ceil_div(ceil_div(ng, kRowWarps), 128), 128, 0, st>>>(d.c_all, ng, d.fwd_part_all);
    AR_LAUNCH_CHECK();
    head_step_kernel<<<ceil_div(ng, kHeadThreads), kHeadThreads, 0, st>>>(
        d.c_all, d.label_all, ng, nullptr, d.fwd_part_all,
        HeadIO{x.head, x.head_m, x.head_v, x.bn_moving, x.alpha, d.dy_all, x.stepc, x.ticket, x.metrics}, t, d.head_part_all, 0);
    AR_LAUNCH_CHECK();
    // partial row gradients of the local samples -> packed send block
    UpdateArgs a{};
    fill_update(a, 0, &x.users, &x.plan_u, slot, x.ah, x.ru, B);
    fill_update(a, 1, &x.anime, &x.plan_a, slot, x.uh, x.ra, B);
    a.emit_cap = B;
    a.emit_ids[0] = reinterpret_cast<int32_t*>(d.send);
    a.emit_q[0] = d.send + B;
    a.emit_P[0] = d.send + 2 * (size_t)B;
    a.emit_ids[1] = reinterpret_cast<int32_t*>(d.send + tw);
    a.emit_q[1] = d.send + tw + B;
    a.emit_P[1] = d.send + tw + 2 * (size_t)B;
    if ((rc = launch_update(a, true, x.c, d.dy_all + (size_t)d.rank * n, x.stepc, x.alpha, x.l2, t, 0, RegAcc{}, st))) return rc;
    AR_NCCL(nc->AllGather(d.send, d.recv, 2 * tw, ncclFloat32, comm, st));
    MergeArgs m{};
    m.tab[0] = x.users;
    m.tab[1] = x.anime;
    m.recv = d.recv;
    m.n_ranks = G;
    m.B = B;
    m.blocks_tab0 = ceil_div((int64_t)G * B, kRowWarps);
    const RegAcc ss = reg_of(x);
    AR_DISPATCH_NV(dim, rows_merge_update_kernel<NV><<<2 * m.blocks_tab0, kRowThreads, 0, st>>>(
                            m, x.alpha, l2x2, t, x.mode == AR_ADAM_REPLAY ? 1 : 0, ss));
    AR_LAUNCH_CHECK();
    if (la) {
      AR_CUDA(cudaEventRecord(la->ev_upd[s & 1], st));
      if (ahead) AR_CUDA(cudaStreamWaitEvent(st, la->ev_ahead, 0));
    }
    if (x.mode == AR_ADAM_DENSE) {
      if ((rc = launch_flush(&x.users, x.alpha, x.l2, t, ss, st))) return rc;
      if ((rc = launch_flush(&x.anime, x.alpha, x.l2, t, ss, st))) return rc;
    }
  }
  return AR_OK;
}

}  // namespace ar

extern "C" int ar_nccl_unique_id(void* id_out_host) {
  AR_REQUIRE(id_out_host, "ar_nccl_unique_id: null pointer");
  ar::NcclApi* nc = ar::nccl_api();
  if (!nc) return AR_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  AR_NCCL(nc->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out_host)));
  return AR_OK;
}

extern "C" int ar_comm_init(const void* id_host, int32_t n_ranks, int32_t rank, void** comm_out) {
  AR_REQUIRE(id_host && comm_out, "ar_comm_init: null pointer");
  ar::NcclApi* nc = ar::nccl_api();
  if (!nc) return AR_ERR_NCCL;
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  ncclComm_t comm = nullptr;
  AR_NCCL(nc->CommInitRank(&comm, n_ranks, id, rank));
  *comm_out = comm;
  return AR_OK;
}

extern "C" int ar_comm_destroy(void* comm) {
  if (!comm) return AR_OK;
  ar::NcclApi* nc = ar::nccl_api();
  if (!nc) return AR_ERR_NCCL;
  AR_NCCL(nc->CommDestroy((ncclComm_t)comm));
  return AR_OK;
}

extern "C" int ar_train_steps_dist(const ar_train_ctx* ctx, const ar_dist_ctx* d, int64_t epoch_step0,
                                   int32_t slot0, int64_t t0, int32_t n_steps, void* stream) {
  int rc = ar::check_ctx(ctx, slot0, n_steps);
  if (rc) return rc;
  if ((rc = ar::check_dist(ctx, d))) return rc;
  return ar::run_steps_dist(*ctx, *d, epoch_step0, slot0, t0, n_steps, (cudaStream_t)stream);
}

// Sharded similarity helpers: all-gather equally sized per-rank buffers (partial top-k lists).
extern "C" int ar_allgather_bytes(void* comm, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
  AR_REQUIRE(comm && send && recv && bytes_per_rank >= 0, "ar_allgather_bytes: bad argument");
  ar::NcclApi* nc = ar::nccl_api();
  if (!nc) return AR_ERR_NCCL;
  AR_NCCL(nc->AllGather(send, recv, (size_t)bytes_per_rank, ncclUint8, (ncclComm_t)comm, (cudaStream_t)stream));
  return AR_OK;
}
