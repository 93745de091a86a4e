// Multi-GPU training, ROW-SHARDED tables (BASELINE cfg5 / SURVEY §8e row 2).  Included by train.cu.
//
// Tables too large to replicate (10 M x 256 with Adam state = 31 GB) live sharded: global row g belongs to
// rank g % G at local index g / G, together with its Adam slots and last_step.  One process per GPU; every
// rank trains on its own B samples of the global batch.  Per chunk of steps (off the critical path):
//   plan_build (local samples, GLOBAL row ids) -> shard_partition: split each step's distinct rows by
//   owner (ascending within an owner), remember where every sample's row will sit in the step's row cache
//   -> one NCCL all-to-all ships all request lists of the chunk to their owners.
// Per step:
//   owner: catch-up of the requested rows (the lowest requesting rank's entry leads, so a row is replayed
//          once) -> gather them into per-requester blocks -> [NCCL all-to-all of rows]
//   requester: embed_fwd on the received row cache (the same kernel, cache rows instead of table rows)
//          -> [NCCL all-gather of c, labels] -> head_step over the GLOBAL batch (SyncBN, identical everywhere)
//          -> rows_update in emit mode: per distinct row the partial sums (P[dim], q) written straight into
//             the per-owner send blocks -> [NCCL all-to-all of partial gradients back to the owners]
//   owner: shard_merge_update: per requested row the leading entry adds the partials of all ranks in rank
//          order, recomputes 1/||w|| from its own row and applies Adam.
// Per-step traffic per GPU is ~B*(2*dim + 4)*4 bytes each way per table whatever G is; results equal the
// single-GPU run on the concatenated batch to rounding (tests/test_gpu_dist.py).
namespace ar {

constexpr int kShardPlanThreads = 1024;
constexpr int kShardMaxRanks = 8;

struct ShardPlanArgs {
  ar_plan plan;
  int G, n_slots;
  int32_t* req_send;   // [G][n_slots][B]
  int32_t* emit_map;   // [n_slots][B]  plan segment -> cache row (owner * B + position)
  int32_t* cache_idx;  // [n_slots][B]  sample -> cache row
  int32_t* max_count;  // [1]
};

__global__ void __launch_bounds__(kShardPlanThreads, 1) shard_partition_kernel(ShardPlanArgs a) {
  __shared__ int warp_tot[kShardPlanThreads / 32];
  __shared__ int owner_tot[kShardMaxRanks];
  const int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int B = a.plan.batch_cap, G = a.G;
  const int32_t* uniq = a.plan.uniq + (int64_t)slot * B;
  const int32_t* off = a.plan.off + (int64_t)slot * (B + 1);
  const int32_t* order = a.plan.order + (int64_t)slot * B;
  const int n_uniq = a.plan.meta[(int64_t)slot * 4];
  int32_t* emap = a.emit_map + (int64_t)slot * B;
  const int per = (B + kShardPlanThreads - 1) / kShardPlanThreads;
  const int lo = tid * per;
  for (int o = 0; o < G; ++o) {
    int cnt = 0;
    for (int e = 0; e < per; ++e) {
      const int i = lo + e;
      if (i < n_uniq && (uniq[i] % G) == o) ++cnt;
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (tid < 32) {
      const int w = warp_tot[tid];
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, d);
        if (tid >= d) wi += t;
      }
      warp_tot[tid] = wi - w;
      if (tid == 31) owner_tot[o] = wi;
    }
    __syncthreads();
    int pos = warp_tot[wid] + incl - cnt;
    int32_t* dst = a.req_send + ((int64_t)o * a.n_slots + slot) * B;
    for (int e = 0; e < per; ++e) {
      const int i = lo + e;
      if (i < n_uniq && (uniq[i] % G) == o) {
        dst[pos] = uniq[i];
        emap[i] = o * B + pos;
        ++pos;
      }
    }
    __syncthreads();
    for (int i = owner_tot[o] + tid; i < B; i += kShardPlanThreads) dst[i] = 0x7fffffff;
    if (tid == 0) atomicMax(a.max_count, owner_tot[o]);
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  int32_t* cidx = a.cache_idx + (int64_t)slot * B;
  for (int s = tid; s < n_uniq; s += kShardPlanThreads) {
    const int row = emap[s];
    for (int j = off[s]; j < off[s + 1]; ++j) cidx[order[j]] = row;
  }
}

struct ShardServeArgs {
  ar_table tab;              // this rank's shard
  const int32_t* req;        // ids requested from me: rank p's list at req + p*req_stride, `cap` valid slots
  int64_t req_stride;
  int G, B, cap;
  float* rows_out;           // [G][B][dim]
  const float* grad;         // [G][B][dim+4] partial gradients received (merge only)
};

// lowest rank whose list holds `id` leads the row (so it is replayed / updated exactly once)
__device__ __forceinline__ bool shard_leads(const ShardServeArgs& a, int p, int id) {
  for (int r = 0; r < p; ++r)
    if (find_id(a.req + r * a.req_stride, a.cap, id) >= 0) return false;
  return true;
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads)
shard_catchup_kernel(ShardServeArgs a, const float* __restrict__ alpha, float l2x2, int64_t t_target, RegAcc reg) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (e >= a.G * a.cap) return;
  const int p = e / a.cap, j = e - p * a.cap;
  const int id = __ldg(a.req + p * a.req_stride + j);
  if (id == 0x7fffffff || !shard_leads(a, p, id)) return;
  const int row = id / a.G;
  const int64_t last = a.tab.last_step[row];
  if (last >= t_target) return;
  const int d4 = a.tab.dim >> 2;
  const size_t o = (size_t)row * a.tab.dim;
  RowTile<NV> w, m, v;
  w.load(a.tab.W + o, d4, lane);
  m.load(a.tab.m + o, d4, lane);
  v.load(a.tab.v + o, d4, lane);
  double regd = 0.0;
  replay_l2<NV>(w, m, v, alpha, reg.stepw, last, t_target, l2x2, lane, regd);
  w.store(a.tab.W + o, d4, lane);
  m.store(a.tab.m + o, d4, lane);
  v.store(a.tab.v + o, d4, lane);
  if (lane == 0) a.tab.last_step[row] = (int32_t)t_target;
  reg_commit(regd, reg, lane);
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads) shard_gather_kernel(ShardServeArgs a) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (e >= a.G * a.cap) return;
  const int p = e / a.cap, j = e - p * a.cap;
  const int id = __ldg(a.req + p * a.req_stride + j);
  if (id == 0x7fffffff) return;
  const int d4 = a.tab.dim >> 2;
  RowTile<NV> w;
  w.load(a.tab.W + (size_t)(id / a.G) * a.tab.dim, d4, lane);
  w.store(a.rows_out + ((size_t)p * a.B + j) * a.tab.dim, d4, lane);
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads)
shard_merge_update_kernel(ShardServeArgs a, const float* __restrict__ alpha, float l2x2, int64_t t,
                          RegAcc reg) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (e >= a.G * a.cap) return;
  const int p = e / a.cap, j = e - p * a.cap;
  const int id = __ldg(a.req + p * a.req_stride + j);
  if (id == 0x7fffffff) return;
  int found = -1;
  if (lane < a.G) found = (lane == p) ? j : find_id(a.req + lane * a.req_stride, a.cap, id);
  const unsigned have = __ballot_sync(0xffffffffu, found >= 0);
  if ((__ffs(have) - 1) != p) return;  // a lower rank's entry leads this row
  const int dim = a.tab.dim, d4 = dim >> 2, gs = dim + 4;
  RowTile<NV> acc;
  acc.zero();
  float q = 0.f;
  for (unsigned rem = have; rem; rem &= rem - 1) {  // ascending rank order: deterministic
    const int r = __ffs(rem) - 1;
    const int pos = __shfl_sync(0xffffffffu, found, r);
    const float* g = a.grad + ((size_t)r * a.B + pos) * gs;
    RowTile<NV> part;
    part.load(g, d4, lane);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc.x[k].x += part.x[k].x; acc.x[k].y += part.x[k].y; acc.x[k].z += part.x[k].z; acc.x[k].w += part.x[k].w;
    }
    q += __ldg(g + dim);
  }
  finish_row<NV>(a.tab, id / a.G, acc, q, -1.0f, alpha, l2x2, t, 0, reg, lane);
}

static int shard_alltoall(NcclApi* nc, ncclComm_t comm, int G, const void* send, void* recv, size_t stride_bytes,
                          size_t count_bytes, cudaStream_t st) {
  AR_NCCL(nc->GroupStart());
  for (int p = 0; p < G; ++p) {
    AR_NCCL(nc->Send((const char*)send + p * stride_bytes, count_bytes, ncclUint8, p, comm, st));
    AR_NCCL(nc->Recv((char*)recv + p * stride_bytes, count_bytes, ncclUint8, p, comm, st));
  }
  AR_NCCL(nc->GroupEnd());
  return AR_OK;
}

static int check_shard(const ar_train_ctx* ctx, const ar_shard_ctx* h) {
  AR_REQUIRE(h && h->comm, "ar_train_steps_sharded: null shard ctx / communicator");
  AR_REQUIRE(h->n_ranks >= 1 && h->n_ranks <= kShardMaxRanks && h->rank >= 0 && h->rank < h->n_ranks,
             "ar_train_steps_sharded: rank %d / n_ranks %d unsupported (1..%d)", h->rank, h->n_ranks, kShardMaxRanks);
  for (int t = 0; t < 2; ++t)
    AR_REQUIRE(h->req_send[t] && h->req_recv[t] && h->emit_map[t] && h->cache_idx[t] && h->rows_out[t] && h->rows_in[t] &&
                   h->grad_send[t] && h->grad_recv[t],
               "ar_train_steps_sharded: null buffer in shard ctx");
  AR_REQUIRE(h->max_count && h->c_all && h->label_all && h->dy_all && h->fwd_part_all && h->head_part_all,
             "ar_train_steps_sharded: null buffer in shard ctx");
  (void)ctx;
  return AR_OK;
}

}  // namespace ar

extern "C" int ar_shard_plan(const ar_plan* plan_u, const ar_plan* plan_a, int32_t n_steps, const ar_shard_ctx* h,
                             void* stream) {
  using namespace ar;
  AR_REQUIRE(plan_u && plan_a, "ar_shard_plan: null plan");
  int rc = check_shard(nullptr, h);
  if (rc) return rc;
  AR_REQUIRE(n_steps >= 0 && n_steps <= plan_u->n_slots && n_steps <= plan_a->n_slots, "ar_shard_plan: n_steps exceeds the plans");
  NcclApi* nc = nccl_api();
  if (!nc) return AR_ERR_NCCL;
  cudaStream_t st = (cudaStream_t)stream;
  AR_CUDA(cudaMemsetAsync(h->max_count, 0, 2 * sizeof(int32_t), st));
  if (n_steps == 0) return AR_OK;
  const ar_plan* plans[2] = {plan_u, plan_a};
  for (int t = 0; t < 2; ++t) {
    ShardPlanArgs a{};
    a.plan = *plans[t];
    a.G = h->n_ranks;
    a.n_slots = plans[t]->n_slots;
    a.req_send = h->req_send[t];
    a.emit_map = h->emit_map[t];
    a.cache_idx = h->cache_idx[t];
    a.max_count = h->max_count + t;
    shard_partition_kernel<<<n_steps, kShardPlanThreads, 0, st>>>(a);
    AR_LAUNCH_CHECK();
    const size_t blk = (size_t)plans[t]->n_slots * plans[t]->batch_cap * sizeof(int32_t);
    if ((rc = shard_alltoall(nc, (ncclComm_t)h->comm, h->n_ranks, h->req_send[t], h->req_recv[t], blk,
                             (size_t)n_steps * plans[t]->batch_cap * sizeof(int32_t), st)))
      return rc;
  }
  return AR_OK;
}

extern "C" int ar_train_steps_sharded(const ar_train_ctx* ctx, const ar_shard_ctx* h, int64_t epoch_step0, int32_t slot0,
                                      int64_t t0, int32_t n_steps, int32_t cap, void* stream) {
  using namespace ar;
  int rc = check_ctx(ctx, slot0, n_steps);
  if (rc) return rc;
  if ((rc = check_shard(ctx, h))) return rc;
  const ar_train_ctx& x = *ctx;
  AR_REQUIRE(cap > 0 && cap <= x.plan_u.batch_cap && cap <= x.plan_a.batch_cap, "ar_train_steps_sharded: cap %d outside (0, batch_cap]", cap);
  NcclApi* nc = nccl_api();
  if (!nc) return AR_ERR_NCCL;
  ncclComm_t comm = (ncclComm_t)h->comm;
  cudaStream_t st = (cudaStream_t)stream;
  const int dim = x.users.dim, G = h->n_ranks, B = x.batch, gs = dim + 4;
  const float l2x2 = (float)(2.0 * (double)x.l2);
  const ar_table* tabs[2] = {&x.users, &x.anime};
  const ar_plan* plans[2] = {&x.plan_u, &x.plan_a};
  for (int s = 0; s < n_steps; ++s) {
    const int64_t e = epoch_step0 + s;
    const int64_t base = e * (int64_t)B;
    if (base >= x.n_samples) break;
    const int n = (int)std::min<int64_t>(B, x.n_samples - base);  // identical on every rank (caller's contract)
    const int slot = slot0 + s;
    const int64_t t = t0 + s + 1;
    const int ng = n * G;
    ShardServeArgs sv[2];
    for (int k = 0; k < 2; ++k) {
      const int Bc = plans[k]->batch_cap;
      sv[k].tab = *tabs[k];
      sv[k].req = h->req_recv[k] + (int64_t)slot * Bc;
      sv[k].req_stride = (int64_t)plans[k]->n_slots * Bc;
      sv[k].G = G; sv[k].B = Bc; sv[k].cap = cap;
      sv[k].rows_out = h->rows_out[k];
      sv[k].grad = h->grad_recv[k];
      const int blocks = ceil_div((int64_t)G * cap, kRowWarps);
      if (x.mode == AR_ADAM_REPLAY) {
        AR_DISPATCH_NV(dim, shard_catchup_kernel<NV><<<blocks, kRowThreads, 0, st>>>(sv[k], x.alpha, l2x2, t - 1, reg_of(x)));
        AR_LAUNCH_CHECK();
      }
      AR_DISPATCH_NV(dim, shard_gather_kernel<NV><<<blocks, kRowThreads, 0, st>>>(sv[k]));
      AR_LAUNCH_CHECK();
    }
    for (int k = 0; k < 2; ++k)
      if ((rc = shard_alltoall(nc, comm, G, h->rows_out[k], h->rows_in[k], (size_t)plans[k]->batch_cap * dim * 4,
                               (size_t)cap * dim * 4, st)))
        return rc;
    const int32_t* cu = h->cache_idx[0] + (int64_t)slot * x.plan_u.batch_cap;
    const int32_t* ca = h->cache_idx[1] + (int64_t)slot * x.plan_a.batch_cap;
    AR_DISPATCH_NV(dim, embed_fwd_kernel<NV><<<ceil_div(n, kRowWarps), kRowThreads, 0, st>>>(
                            h->rows_in[0], h->rows_in[1], dim, cu, ca, n, nullptr, x.uh, x.ah, x.c, x.ru, x.ra, nullptr));
    AR_LAUNCH_CHECK();
    AR_NCCL(nc->GroupStart());
    AR_NCCL(nc->AllGather(x.c, h->c_all, (size_t)n, ncclFloat32, comm, st));
    AR_NCCL(nc->AllGather(x.label + base, h->label_all, (size_t)n, ncclFloat32, comm, st));
    AR_NCCL(nc->GroupEnd());
    c_partials_kernel<<<ceil_div(ceil_div(ng, kRowWarps), 128), 128, 0, st>>>(h->c_all, ng, h->fwd_part_all);
    AR_LAUNCH_CHECK();
    head_step_kernel<<<ceil_div(ng, kHeadThreads), kHeadThreads, 0, st>>>(
        h->c_all, h->label_all, ng, nullptr, h->fwd_part_all,
        HeadIO{x.head, x.head_m, x.head_v, x.bn_moving, x.alpha, h->dy_all, x.stepc, x.ticket, x.metrics}, t, h->head_part_all, 0);
    AR_LAUNCH_CHECK();
    UpdateArgs a{};
    fill_update(a, 0, &x.users, &x.plan_u, slot, x.ah, x.ru, B);
    fill_update(a, 1, &x.anime, &x.plan_a, slot, x.uh, x.ra, B);
    for (int k = 0; k < 2; ++k) {
      a.emit_P[k] = h->grad_send[k];
      a.emit_map[k] = h->emit_map[k] + (int64_t)slot * plans[k]->batch_cap;
    }
    a.emit_cap = B;
    a.emit_stride = gs;
    if ((rc = launch_update(a, true, x.c, h->dy_all + (size_t)h->rank * n, x.stepc, x.alpha, x.l2, t, 0, RegAcc{}, st))) return rc;
    for (int k = 0; k < 2; ++k)
      if ((rc = shard_alltoall(nc, comm, G, h->grad_send[k], h->grad_recv[k], (size_t)plans[k]->batch_cap * gs * 4,
                               (size_t)cap * gs * 4, st)))
        return rc;
    const RegAcc ss = reg_of(x);
    for (int k = 0; k < 2; ++k) {
      const int blocks = ceil_div((int64_t)G * cap, kRowWarps);
      AR_DISPATCH_NV(dim, shard_merge_update_kernel<NV><<<blocks, kRowThreads, 0, st>>>(sv[k], x.alpha, l2x2, t, ss));
      AR_LAUNCH_CHECK();
    }
    if (x.mode == AR_ADAM_DENSE) {
      if ((rc = launch_flush(&x.users, x.alpha, x.l2, t, ss, st))) return rc;
      if ((rc = launch_flush(&x.anime, x.alpha, x.l2, t, ss, st))) return rc;
    }
  }
  return AR_OK;
}
