/*
 * animerec.h -- C-ABI of libanimerec.so: the B200 (sm_100a) hot path of
 * Dyrutter/anime_recommendations.
 *
 * The reference has no FFI of its own (it is pure Python on TensorFlow/NumPy,
 * SURVEY.md §8b), so every entry point below names the reference call site whose
 * arithmetic it replaces.  Conventions:
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller
 *     (torch tensors on the Python side) unless it is marked "host";
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and
 *     never synchronise unless documented;
 *   - return 0 on success, a negative ar_status otherwise; ar_last_error() returns
 *     a thread-local message; nothing throws;
 *   - tables are row-major float32 (n_rows, dim), dim % 4 == 0, dim <= 512,
 *     16-byte aligned.
 */
#ifndef ANIMEREC_H_
#define ANIMEREC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  AR_OK = 0,
  AR_ERR_INVALID = -1,     /* bad argument (null pointer, dim, k, batch too large ...) */
  AR_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
  AR_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve */
  AR_ERR_NCCL = -4
} ar_status;

const char* ar_last_error(void);
/* version of this ABI; bumped on any signature change */
int ar_abi_version(void);
/* 0 when a device with compute capability 10.x is current, AR_ERR_UNSUPPORTED otherwise */
int ar_check_device(void);

/* ------------------------------------------------------------------ Half A: training
 * Replaces the TensorFlow train_step behind `model.fit` (neural_network.py:210-217) for the
 * model of neural_network.py:66-106.
 */

#define AR_MAX_BATCH 16384   /* per-step samples the in-shared-memory plan sort handles */
#define AR_HEAVY_LEN 64      /* rows hit by more than this many samples of one step take the CTA path */

/* One embedding table and its Adam slots (Keras: Embedding.embeddings + optimizer m, v). */
typedef struct {
  int32_t n_rows;
  int32_t dim;
  float* W;
  float* m;
  float* v;
  int32_t* last_step; /* (n_rows) optimizer step (1-based) already applied to the row; 0 = none */
} ar_table;

/* Per-step dedup plan of one table: samples grouped by row (stable, ascending row).
 * Arrays hold `n_slots` consecutive steps. */
typedef struct {
  int32_t batch_cap;  /* stride of order/uniq; off has stride batch_cap+1 */
  int32_t heavy_cap;  /* stride of heavy */
  int32_t n_slots;
  int32_t* order;     /* [slot][batch_cap]  sample id within the step, sorted by (row, sample) */
  int32_t* uniq;      /* [slot][batch_cap]  distinct rows ascending */
  int32_t* off;       /* [slot][batch_cap+1] segment starts into order; off[n_uniq] = n */
  int32_t* meta;      /* [slot][4]  n_uniq, n_heavy, n (samples in the step), 0 */
  int32_t* heavy;     /* [slot][heavy_cap] segment ids longer than AR_HEAVY_LEN */
  uint8_t* in_prev;   /* optional [slot][batch_cap]: 1 = the distinct row is also touched by step slot-1 (of any
                         rank); written by ar_plan_link, enables the look-ahead catch-up of AR_ADAM_REPLAY */
} ar_plan;

/* Build the plan of `n_steps` consecutive steps of one table.  Step s (0-based within the
 * call) covers idx[(step0+s)*batch : min(n_total, (step0+s+1)*batch)] and is written to slot s.
 * idx values must lie in [0, n_rows).  batch <= AR_MAX_BATCH.
 * Replaces TF's IndexedSlices -> UnsortedSegmentSum aggregation. */
int ar_plan_build(const int32_t* idx, int64_t n_total, int32_t batch, int64_t step0,
                  int32_t n_steps, const ar_plan* plan, void* stream);

/* List form: step s groups the `counts[s]` keys at keys[s*stride ...] (counts on the device, stride <=
 * AR_MAX_BATCH); the plan's sample ids are positions in that list.  Used by ar_peer_plan. */
int ar_plan_build_lists(const int32_t* keys, int32_t stride, const int32_t* counts, int32_t n_steps,
                        const ar_plan* plan, void* stream);

/* After ar_plan_build of `n_steps` slots: fill plan->in_prev.  uniq_all / meta_all: every rank's plan.uniq
 * ([n_ranks][n_slots][batch_cap]) and plan.meta ([n_ranks][n_slots][4]) of the same chunk, all-gathered, for
 * replicated multi-GPU training; null = this plan's own lists (single GPU). */
int ar_plan_link(const ar_plan* plan, int32_t n_steps, const int32_t* uniq_all, const int32_t* meta_all,
                 int32_t n_ranks, void* stream);

/* Replay schedule of AR_ADAM_REPLAY, built at plan time (ar_plan_sched) for the same slots as the two plans.
 * With gap = (this step) - (step of the row's previous touch or of the last full flush), a distinct row of slot
 * s with gap >= 2 becomes one ITEM of slot s (rows with gap == 1 are left current by the previous step's own
 * update): "replay the row's gap-1 missed pure-L2 Adam steps, up to global step t(s)-1".  The item may be
 * processed as soon as the row's previous touch is complete, and at most `depth` steps before its own step:
 *   k = min(gap-1, depth, s)      sublist of the item; it is released when step s-k-1 of the chunk is done
 *                                 (k == s: before the chunk's first step);
 * within a sublist items are ordered longest replay first (log2 buckets).  Rows with gap >= AR_SCHED_SPLIT_GAP
 * are split into ceil(dim/32) items of 32 consecutive elements each when the slot's capacity allows (a long
 * replay is a serial chain per element; splitting shortens its critical path).
 * Item code: bit 31 table (0 users, 1 anime) | bit 30 split | bits 29..26 part | bits 25..0 row. */
#define AR_SCHED_MAX_DEPTH 4
#define AR_SCHED_SUB (AR_SCHED_MAX_DEPTH + 2)  /* stride of sub / cursor */
#define AR_SCHED_SPLIT_GAP 256
#define AR_SCHED_MAX_ROWS (1 << 26)
#define AR_SCHED_PARTS 296   /* row-range parts of the plan-time walk (bounds has n_slots*(AR_SCHED_PARTS+1) entries) */
typedef struct {
  int32_t cap;        /* stride of codes / glen; >= plan_u.batch_cap + plan_a.batch_cap */
  int32_t n_slots;
  int32_t* codes;     /* [slot][cap]  item codes, sublist k at [sub[k], sub[k+1]) */
  int32_t* glen;      /* [slot][cap]  gap of the item's row */
  int32_t* sub;       /* [slot][AR_SCHED_SUB]  sublist starts; sub[k] for k > depth = number of items of the slot */
  int32_t* cursor;    /* [slot][AR_SCHED_SUB]  run-time state of the training kernel, zeroed by ar_plan_sched:
                         [k] items of sublist k handed out so far, [AR_SCHED_SUB-1] items completed */
  int32_t* gap_u;     /* [slot][plan_u.batch_cap] scratch: gap per distinct user row */
  int32_t* gap_a;     /* [slot][plan_a.batch_cap] */
  int32_t* bounds;    /* [n_slots][AR_SCHED_PARTS+1] scratch */
} ar_sched;

/* Build the schedule of `n_steps` planned slots whose global optimizer steps are t0+1 .. t0+n_steps.
 * seen_u / seen_a: (n_rows) int32, zero-initialised by the caller once and then only passed back: the global step
 * of every row's latest planned touch.  t_flush: global step every row was last brought to by ar_table_flush (0 =
 * never).  dim: embedding dimension (decides the split).  Chunks must be scheduled in training order, each
 * exactly once, and a schedule is consumed by exactly one ar_train_steps call. */
int ar_plan_sched(const ar_plan* plan_u, const ar_plan* plan_a, int32_t n_steps, int64_t t0, int64_t t_flush,
                  int32_t* seen_u, int32_t n_rows_u, int32_t* seen_a, int32_t n_rows_a, int32_t depth, int32_t dim,
                  const ar_sched* sched, void* stream);

typedef enum {
  AR_ADAM_REPLAY = 0,  /* reference-equivalent: missed dense steps of a row are replayed when it is next touched */
  AR_ADAM_DENSE = 1,   /* reference-literal: every row of both tables is updated every step */
  AR_ADAM_TOUCHED = 2  /* north-star literal: rows absent from the batch are left alone (NOT the reference's arithmetic) */
} ar_adam_mode;

/* Everything one training run needs; all pointers are device pointers. */
typedef struct {
  ar_table users;
  ar_table anime;
  float* head;         /* [4] Dense kernel w, Dense bias b, BN gamma, BN beta  (neural_network.py:97-99) */
  float* head_m;       /* [4] Adam m */
  float* head_v;       /* [4] Adam v */
  float* bn_moving;    /* [2] moving_mean, moving_variance */
  const float* alpha;  /* alpha[t] = lr_t*sqrt(1-b2^t)/(1-b1^t) for global 1-based step t (alpha[0] unused) */
  const int32_t* iu;   /* (n_samples) user row per sample, epoch visit order */
  const int32_t* ia;   /* (n_samples) anime row per sample */
  const float* label;  /* (n_samples) rating in [0,1] */
  int64_t n_samples;
  int32_t batch;
  float l2;            /* embeddings_regularizer factor (config l2_reg_factor) */
  int32_t mode;        /* ar_adam_mode */
  ar_plan plan_u;
  ar_plan plan_a;
  /* per-step scratch, each sized for `batch` samples */
  float* uh;           /* (batch, dim) normalised user rows of the step */
  float* ah;           /* (batch, dim) normalised anime rows */
  float* c;            /* (batch) cosine */
  float* ru;           /* (batch) 1/||u|| */
  float* ra;           /* (batch) 1/||a|| */
  float* dy;           /* (batch) dLoss/dy = (p - t)/n per sample */
  double* fwd_part;    /* (2*ceil(batch/8)) per-CTA (sum c, sum c^2) of the forward kernel */
  double* head_part;   /* (8*ceil(batch/256)) per-CTA partial sums of the head kernel */
  float* stepc;        /* (16) per-step scalars the head hands to the row update */
  uint32_t* ticket;    /* (1) zero-initialised arrival counter of the head kernel */
  /* per-step outputs, indexed by global step t (1-based): metrics[t*4 + {0: mean BCE, 1: mean
   * squared error, 2: n, 3: batch mean of z}] */
  float* metrics;
  /* L2-regulariser term of the reported loss (neural_network.py:73,78,85: `loss` = BCE + l2*(sum U^2 + sum A^2),
   * evaluated with the weights BEFORE each step).  Every (row, step) pair passes exactly once through a replay
   * step, a row update or a flush; each adds stepw[t] * ||row before step t||^2 to the fixed-point accumulator
   * reg_acc[0] (uint64, value * reg_scale; integer adds => order-independent, bit-reproducible).  After the
   * tables are flushed to step T:  reg_acc / reg_scale = sum_{t<=T} stepw[t] * (sum U_{t-1}^2 + sum A_{t-1}^2).
   * stepw == null or reg_acc == null: not accumulated. */
  unsigned long long* reg_acc;
  const float* stepw;  /* stepw[t], indexed like alpha: weight of global step t (samples in the step / batch) */
  float reg_scale;     /* power of two */
  /* optional (multi-GPU paths, AR_ADAM_REPLAY): 2 * (3 * (plan_u.batch_cap + plan_a.batch_cap) + 4) int32 of
   * scratch for the per-step classify launch of ar_train_steps_dist / _sharded / _peer; null = plan order */
  int32_t* sched_ws;
  /* single GPU (ar_train_steps): the plan-time replay schedule (AR_ADAM_REPLAY) and its look-ahead depth */
  ar_sched sched;
  int32_t depth;
  /* single GPU: device workspace of the persistent step kernel, ar_chunk_ws_info(...)[0] bytes, 256-byte aligned;
   * holds the grid-barrier state, per-step metric partials, the in-kernel timeline and the heavy-row partials */
  void* chunk_ws;
  /* (4) int32 device counters, zeroed by the caller: [0] rows a row update found behind schedule (must stay 0 in
   * AR_ADAM_REPLAY: a non-zero value means the replay schedule and the plans disagree), [1] waits that timed out
   * inside the step kernel (the run's results are invalid) */
  int32_t* health;
} ar_train_ctx;

/* Run `n_steps` consecutive training steps as ONE persistent kernel (one CTA per SM): per step a forward phase,
 * a grid barrier, the head (redundantly and bit-identically in every CTA), the row-update phase, a grid barrier;
 * in AR_ADAM_REPLAY the other warps of every CTA work through the replay schedule `depth` steps ahead of the
 * step warps; AR_ADAM_DENSE adds a third phase that moves every other row.
 * Epoch-local step e = epoch_step0 + s reads samples [e*batch, min(n_samples,(e+1)*batch)) and plan slot s; its
 * global Adam step is t = t0 + s + 1.  slot0 must be 0.  Replaces Keras Model.train_step x n_steps. */
int ar_train_steps(const ar_train_ctx* ctx, int64_t epoch_step0, int32_t slot0, int64_t t0,
                   int32_t n_steps, void* stream);

/* Layout of ar_train_ctx.chunk_ws for plans of `n_slots` slots, `batch_cap` samples per step and rows of `dim`
 * floats.  Writes to the HOST array out[8]: [0] total bytes, [1] byte offset of the timeline: int64
 * [n_slots][8] %globaltimer stamps (ns) written by CTA 0 for every step of the last ar_train_steps call:
 * 0 gate entered (waiting for the step's replay items), 1 forward begins, 2 first grid barrier passed, 3 head done,
 * 4 row update done (CTA 0's share), 5 second grid barrier passed, 6 dense phase done; [2] byte offset of the
 * statistics: uint64 [8]: 0 cycles replay warps spent inside items (summed over warps), 1 items, 2 replayed
 * element-steps, 3 kernel wall time (ns), 4 number of replay warps, 5 SM clock cycles of CTA 0 over the kernel. */
int ar_chunk_ws_info(int32_t n_slots, int32_t batch_cap, int32_t dim, int64_t* out_host);

/* ---- multi-GPU training, replicated tables (one process per GPU, NCCL over NVLink) ----
 * The reference's only data-parallel path is tf.distribute TPUStrategy (neural_network.py:142-147,
 * 173-182: replicated variables, implicit gradient all-reduce); this is its B200 counterpart.
 * NCCL is resolved with dlopen("libnccl.so.2") at run time. */
int ar_nccl_unique_id(void* id_out_host /* 128 bytes, host */);
int ar_comm_init(const void* id_host /* 128 bytes */, int32_t n_ranks, int32_t rank, void** comm_out);
int ar_comm_destroy(void* comm);

typedef struct {
  void* comm;            /* from ar_comm_init */
  int32_t n_ranks;
  int32_t rank;
  float* c_all;          /* (n_ranks*batch) cosines of the global batch, rank-major */
  float* label_all;      /* (n_ranks*batch) */
  float* dy_all;         /* (n_ranks*batch) */
  double* fwd_part_all;  /* (2*ceil(n_ranks*batch/8)) */
  double* head_part_all; /* (8*ceil(n_ranks*batch/256)) */
  float* send;           /* (2*batch*(dim+2)) packed partial row gradients of this rank:
                            [ids_u | q_u | P_u(batch,dim) | ids_a | q_a | P_a(batch,dim)], ids int32 */
  float* recv;           /* n_ranks such blocks */
} ar_dist_ctx;

/* ar_train_steps for rank `d->rank` of `d->n_ranks`: ctx holds this rank's samples (every rank must
 * pass the same n_samples and batch); global batch = n_ranks*batch; BatchNorm statistics and the head
 * update are computed over the global batch, row gradients are all-gathered and merged, so all
 * replicas stay bit-identical and equal to a single-GPU run on the concatenated batch. */
int ar_train_steps_dist(const ar_train_ctx* ctx, const ar_dist_ctx* d, int64_t epoch_step0,
                        int32_t slot0, int64_t t0, int32_t n_steps, void* stream);
/* ---- multi-GPU training, ROW-SHARDED tables (BASELINE cfg5: tables too large to replicate) ----
 * Global row g of a table lives on rank g % n_ranks at local index g / n_ranks, with its Adam slots; ctx->users /
 * ctx->anime describe THIS rank's shards; ctx->iu / ia hold GLOBAL row ids of this rank's samples and the plans
 * are built on them.  Per step the owners serve the requested rows (NCCL all-to-all), the requesters run the
 * forward on that row cache, BatchNorm/head run over the global batch, and the partial row gradients travel
 * back to the owners (all-to-all), which merge them and apply Adam.  Equal to a single-GPU run on the
 * concatenated batch up to rounding. */
typedef struct {
  void* comm;            /* from ar_comm_init */
  int32_t n_ranks;       /* <= 8 */
  int32_t rank;
  int32_t* req_send[2];  /* [n_ranks][n_slots][batch_cap] per table (0 users, 1 anime): ids asked of each owner */
  int32_t* req_recv[2];  /* same shape: ids each rank asks of me */
  int32_t* emit_map[2];  /* [n_slots][batch_cap] plan segment -> row of the step's row cache */
  int32_t* cache_idx[2]; /* [n_slots][batch_cap] sample -> row of the step's row cache */
  int32_t* max_count;    /* [2] largest per-owner request list of the planned chunk, per table */
  float* rows_out[2];    /* [n_ranks][batch_cap][dim] rows served to each requester */
  float* rows_in[2];     /* [n_ranks][batch_cap][dim] the step's row cache */
  float* grad_send[2];   /* [n_ranks][batch_cap][dim+4] partial gradients (P[dim], q, pad) for each owner */
  float* grad_recv[2];
  float* c_all;          /* as in ar_dist_ctx */
  float* label_all;
  float* dy_all;
  double* fwd_part_all;
  double* head_part_all;
} ar_shard_ctx;

/* After ar_plan_build of both tables for `n_steps` slots: split every step's distinct rows by owner, build the
 * cache maps and ship all request lists of the chunk to their owners (one all-to-all per table).  max_count is
 * written on the device; read it back and pass a `cap` >= both values to ar_train_steps_sharded. */
int ar_shard_plan(const ar_plan* plan_u, const ar_plan* plan_a, int32_t n_steps, const ar_shard_ctx* sh, void* stream);
int ar_train_steps_sharded(const ar_train_ctx* ctx, const ar_shard_ctx* sh, int64_t epoch_step0, int32_t slot0,
                           int64_t t0, int32_t n_steps, int32_t cap, void* stream);

/* ---- multi-GPU training over NVLink PEER MEMORY (row-sharded tables, owner computes) ----
 * Same ownership rule as the row-sharded path, no collective on the step's critical path: every rank works on
 * the samples of the global batch that touch ITS rows and loads the sample's row of the other table straight
 * from the owner's HBM (cudaIpc-mapped shards, 128-bit loads over NVLink) and the cosines every rank published;
 * the ranks meet at two flag barriers per step (spin on peer-written words) folded into the step's kernels.  One process per GPU on ONE node with all-to-all peer access. */
#define AR_PEER_MAX_RANKS 8
#define AR_PEER_HANDLE_BYTES 64
#define AR_PEER_FLAG_WORDS 64
/* handle (host, AR_PEER_HANDLE_BYTES) + byte offset naming dev_ptr inside its cudaMalloc allocation */
int ar_peer_export(const void* dev_ptr, void* handle_out_host, int64_t* offset_out);
/* map another process's allocation (each distinct handle is opened once per process) */
int ar_peer_open(const void* handle_host, int64_t offset, void** ptr_out);
int ar_peer_close_all(void);

typedef struct {
  int32_t n_ranks;                              /* <= AR_PEER_MAX_RANKS */
  int32_t rank;
  float* W_peer[2][AR_PEER_MAX_RANKS];          /* [0 users | 1 anime][rank]: base of that rank's shard; the own
                                                   entry is the local pointer (= ctx->users.W / ctx->anime.W) */
  float* pub_peer[AR_PEER_MAX_RANKS];           /* every rank's published (sample, cosine) list: sel_cap pairs of
                                                   (int32 position in the global batch, float c); two such lists
                                                   back to back (by step parity) for the persistent kernel */
  int32_t* flags_peer[AR_PEER_MAX_RANKS];       /* every rank's AR_PEER_FLAG_WORDS int32, zero-initialised before
                                                   any rank's first step; word 32 != 0: a barrier timed out */
  int32_t sel_cap;                              /* capacity of one selection list = batch_cap of both plans */
  int32_t* sel_key[2];                          /* [n_slots][sel_cap] local row of my table */
  int32_t* sel_samp[2];                         /* [n_slots][sel_cap] position in the global batch */
  int32_t* sel_oth[2];                          /* [n_slots][sel_cap] GLOBAL row of the other table */
  int32_t* sel_cnt[2];                          /* [n_slots] */
  int32_t* max_count;                           /* [2] longest list of the planned chunk per table (> sel_cap:
                                                   overflow, the chunk must not be run) */
  float* label_step;                            /* [n_slots][n_ranks*batch] labels in global-batch order */
  float* c_all;                                 /* (n_ranks*batch) local */
  float* dy_all;                                /* (n_ranks*batch) */
  double* fwd_part_all;                         /* (2 * n_ranks * ceil(sel_cap/1024)) */
  double* head_part_all;
  /* persistent peer kernel (AR_ADAM_REPLAY with a replay schedule; all of these set, else the staged kernels run).
   * Nothing here needs a fence on the step's critical path: cosines and batch sums travel as 64-bit words that carry
   * their own step tag (written with one store, valid when the tag matches), rows are guarded by per-row step words. */
  int32_t* rowflag_peer[2][AR_PEER_MAX_RANKS];  /* [table][rank] per row of that rank's shard: the optimizer step the
                                                   row is at, written (behind a system-scope fence) by its owner */
  uint64_t* pairs_peer[AR_PEER_MAX_RANKS];      /* [rank] that rank's inbox of (position, cosine) words:
                                                   [2 step parities][n_ranks senders][sel_cap], initialised to ~0 */
  uint64_t* hdrin_peer[AR_PEER_MAX_RANKS];      /* [rank] that rank's inbox of list headers: [2][n_ranks][8] words
                                                   (list length, sum c, sum c^2 as 32-bit halves + step tag) */
  float* sel_lab[2];                            /* [n_slots][sel_cap] label of every listed sample */
} ar_peer_ctx;

/* Plan a chunk: iu_all / ia_all / label_all hold every rank's samples of the chunk, rank r's at
 * [r*rank_stride, r*rank_stride + n_local) (GLOBAL row ids; all-gathered by the caller); step s of the chunk
 * is samples [s*batch, min(n_local, (s+1)*batch)) of every rank.  Builds the selection lists, label_step and
 * both plans (+ in_prev when the plans carry it). */
int ar_peer_plan(const int32_t* iu_all, const int32_t* ia_all, const float* label_all, int64_t rank_stride,
                 int64_t n_local, int32_t batch, int32_t n_steps, const ar_plan* plan_u, const ar_plan* plan_a,
                 const ar_peer_ctx* peer, void* stream);
/* ar_train_steps for this rank.  ctx->users / anime: my shards; ctx->uh / ah: (sel_cap, dim) and ctx->c / ru /
 * ra: (sel_cap) scratch; ctx->iu / ia / label are not read (n_samples and batch are).  count_hint >= both
 * max_count values (grid sizing).  Equal to a single-GPU run on the concatenated batch up to rounding. */
int ar_train_steps_peer(const ar_train_ctx* ctx, const ar_peer_ctx* peer, int64_t epoch_step0, int32_t slot0,
                        int64_t t0, int32_t n_steps, int32_t count_hint, void* stream);

/* NCCL all-gather of equally sized byte buffers (sharded top-k lists before ar_topk_merge). */
int ar_allgather_bytes(void* comm, const void* send, void* recv, int64_t bytes_per_rank, void* stream);

/* Bring every row of the table to optimizer step t_target by replaying its missed pure-L2 steps
 * (no-op per row when last_step >= t_target).  Used at epoch end / before validation, saving and
 * similarity in AR_ADAM_REPLAY, and as the "all other rows" half of a dense step.  reg_acc / stepw / reg_scale:
 * the regulariser accumulator of ar_train_ctx (the replayed steps' share of the reported loss), or null. */
int ar_table_flush(const ar_table* tab, const float* alpha, float l2, int64_t t_target,
                   unsigned long long* reg_acc, const float* stepw, float reg_scale, void* stream);

/* Individual stages (the kernels ar_train_steps chains), exported for unit parity tests. */
int ar_embed_fwd(const float* U, const float* A, int32_t dim, const int32_t* iu, const int32_t* ia,
                 int32_t n, float* uh, float* ah, float* c, float* ru, float* ra, void* stream);
int ar_head_step(const float* c, const float* label, int32_t n, float* head, float* head_m,
                 float* head_v, float* bn_moving, const float* alpha, int64_t t, float* dc,
                 float* metrics_row, void* stream);
int ar_rows_catchup(const ar_table* tab, const ar_plan* plan, int32_t slot, const float* alpha,
                    float l2, int64_t t, void* stream);
int ar_rows_update(const ar_table* tab, const ar_plan* plan, int32_t slot, const float* other_hat,
                   const float* c, const float* dy, const float* stepc, const float* rinv,
                   const float* alpha, float l2, int64_t t, int32_t replay, void* stream);

/* Inference forward, Keras `model.predict([users, animes])` (model_recs.py:394): BN uses the
 * moving statistics.  out: (n) float32 probabilities. */
int ar_predict(const float* U, const float* A, int32_t dim, const float* head, const float* bn_moving,
               const int32_t* iu, const int32_t* ia, int64_t n, float* out, void* stream);
/* Validation pass (Keras test_step): adds sum of BCE-from-logits and of squared error over the n
 * samples to sums[0], sums[1] (double, caller zeroes them). */
int ar_eval_sums(const float* U, const float* A, int32_t dim, const float* head, const float* bn_moving,
                 const int32_t* iu, const int32_t* ia, const float* label, int64_t n, double* sums,
                 void* stream);
/* out[0] += sum of squares of the table (L2 regulariser term, neural_network.py:73). */
int ar_sumsq(const float* W, int64_t n_elems, double* out, void* stream);

/* Measurement aid (bench.py): launches blocks x threads (<= 256) threads that each run `iters` x 8 independent
 * sqrt.approx -> add -> rcp.approx chains (the special-function work of one Adam element-step) and write one float
 * to scratch[blocks*threads].  2*8*blocks*threads*iters MUFU ops per launch; the caller times it. */
int ar_bench_sfu(float* scratch, int32_t blocks, int32_t threads, int32_t iters, void* stream);

/* ------------------------------------------------------------------ user_recs: collaborative aggregation
 * Ratings as a CSR by user: indptr (n_users+1) int64, anime_idx / rating aligned with it. */

/* fav_flag[j] = 1 iff rating j is at or above its user's `percentile`-th percentile (np.percentile, linear
 * interpolation, evaluated in double) -- the "favourites" of user_recs.py:359-361 / 380-382 (percentile = 80) and
 * similar_users.py:216 (75).  thr_out (n_users) doubles: the thresholds, or null. */
int ar_user_favourites(const int64_t* indptr, const float* rating, int32_t n_users, double percentile,
                       uint8_t* fav_flag, double* thr_out, void* stream);

/* For every query user q (row index into the CSR): count how many of q's similar users sim_users[q][0..k_sim)
 * (row indices; < 0 = padding) hold each anime among their favourites, drop q's own favourites, and emit the
 * n_recs anime with the highest counts, ties by lower anime index (user_recs.py:761-774 value_counts; the
 * reference's order among equal counts is pandas-version dependent).  out_idx / out_cnt: [n_query][n_recs],
 * idx = -1 past the last recommendation. */
int ar_user_recs(const int64_t* indptr, const int32_t* anime_idx, const uint8_t* fav_flag, int32_t n_anime,
                 const int32_t* query_users, int32_t n_query, const int32_t* sim_users, int32_t k_sim,
                 int32_t n_recs, int32_t* out_idx, int32_t* out_cnt, void* stream);

/* ------------------------------------------------------------------ Half B: cosine top-k */

/* out = W / ||W||_2 per row, the reference's get_weights (similar_anime.py:159-170).  Only used to
 * export normalised tables; the top-k kernels fuse the normalisation. */
int ar_rownorm(const float* W, int64_t n_rows, int32_t dim, float* out, void* stream);

/* Single-query cosine top-k: dists = np.dot(Wn, Wn[q]); rank (similar_anime.py:404-468,
 * similar_users.py:293-312).  cand_mask: optional bitmask, bit r set = row r is a candidate
 * ((n_rows+31)/32 words).  exclude: row to drop (-1 none).  Results sorted by (score desc, row asc);
 * unfilled slots hold idx -1 / score -inf.  workspace: >= ar_topk_query_workspace(n_rows,k) bytes. */
int64_t ar_topk_query_workspace(int64_t n_rows, int32_t k);
int ar_cosine_topk_query(const float* W, int64_t n_rows, int32_t dim, int64_t q,
                         const uint32_t* cand_mask, int64_t exclude, int32_t k, int32_t* out_idx,
                         float* out_score, void* workspace, void* stream);

/* n_queries single-query top-k's in one call: query i is row q0+i, excluded from its own list; out_idx / out_score
 * are [n_queries][k].  Exact fp32; used for embedding sizes above 128, which the tensor-core pass is not built for. */
int ar_cosine_topk_queries(const float* W, int64_t n_rows, int32_t dim, int64_t q0, int64_t n_queries,
                           const uint32_t* cand_mask, int32_t k, int32_t* out_idx, float* out_score,
                           void* workspace, void* stream);

/* Merge `n_lists` partial top-k lists per query (lists[l][query][k], global row ids) into one.
 * lists_sorted != 0: every list is sorted best first (the output order of the top-k entry points), which lets
 * the kernel drop everything below the largest k_out-th entry of any list before merging. */
int ar_topk_merge(const int32_t* idx, const float* score, int32_t n_lists, int64_t n_queries,
                  int32_t k_in, int32_t k_out, int32_t lists_sorted, int32_t* out_idx, float* out_score,
                  void* stream);

/* Exact fp32 re-rank.  Candidates: n_lists lists of up to list_cap row ids per query, layout
 * [list][query][list_cap] (the output layout of ar_cosine_topk_allpairs / the input layout of ar_topk_merge);
 * cand_cnt[list][query] = number of valid entries (null: every entry >= 0 is valid).  Query rows are
 * Wq[q0 + i], candidate rows index Wc.  Keeps the best k by fp32 cosine (duplicates collapse).
 * Optional certification: cand_thr[list][query] = the selection-score bound of everything list `list` left
 * out (ar_cosine_topk_allpairs' out_thr); with eps (+ q_eps[i], optional per query) bounding
 * |selection score - fp32 score|, certified[i] = 1 iff the fp32 top-k of query i is provably the true top-k
 * over ALL rows the lists were drawn from (k-th score >= max_list thr + eps), else 0. */
int ar_cosine_rerank(const float* Wq, int64_t q0, int64_t n_queries, const float* Wc, int32_t dim,
                     const int32_t* cand, const int32_t* cand_cnt, const float* cand_thr, int32_t n_lists,
                     int32_t list_cap, int32_t k, float eps, const float* q_eps, int32_t* out_idx,
                     float* out_score, uint8_t* certified, void* stream);

/* Many-query cosine candidates on the tensor cores (tcgen05.mma, bf16 operands, fp32 accumulation in
 * TMEM, TMA-fed): for query rows [q0, q0+n_q) of Qn and candidate rows [c0, c0+n_c) of Cn -- both
 * ROW-NORMALISED bf16 (q_rows_total|c_rows_total, dim) tables made by ar_rownorm_bf16, dim == 128 --
 * list per query row and candidate chunk every admissible candidate whose bf16-operand score exceeds a
 * running threshold that rises to the (kprime+1)-th (kprime = 16 or 24) best score seen.
 *   exclude_self     drop candidate id == query id (Qn and Cn are the same table); self_ids (optional,
 *                    [n_q]) gives the id to drop per query row when the query table is a gathered subset
 *   watched          optional bit rows [n_q][watched_stride] over candidate ids, set bit = drop
 *                    (model_recs.py:144-155: anime the user already rated)
 *   thr_init         optional [n_q] starting threshold (only scores above it are ever listed)
 *   n_chunks         from ar_allpairs_chunks(n_q, n_c): the candidate range is split so every SM has work
 *   out_idx/out_score [n_chunks][n_q][cap], cap = ar_allpairs_list_cap(kprime); unsorted, -1 / -inf = empty
 *   out_cnt          [n_chunks][n_q] valid entries (kprime <= cnt <= cap once kprime candidates were seen)
 *   out_thr          [n_chunks][n_q] every admissible candidate of the chunk NOT listed scored <= this
 *   dump_scores      optional [n_q][n_c] raw scores (tests only)
 * Replaces the reference's np.dot + np.argsort per query (similar_anime.py:404-409, similar_users.py:
 * 293-296) looped over all rows, and model.predict over (user x all anime) (model_recs.py:394-396). */
/* resid (optional, [n_rows]): ||w_hat - bf16(w_hat)||_2 per row; |bf16 score - fp32 cosine| of a pair is
 * bounded by resid_q + resid_c + resid_q*resid_c (+ fp32 accumulation error). */
int ar_rownorm_bf16(const float* W, int64_t n_rows, int32_t dim, void* out_bf16, float* resid, void* stream);
int32_t ar_allpairs_chunks(int64_t n_q, int64_t n_c);
int32_t ar_allpairs_list_cap(int32_t kprime);
int ar_cosine_topk_allpairs(const void* Qn_bf16, int64_t q_rows_total, int64_t q0, int64_t n_q,
                            const void* Cn_bf16, int64_t c_rows_total, int64_t c0, int64_t n_c,
                            int32_t dim, int32_t kprime, int32_t exclude_self, const int32_t* self_ids,
                            const uint32_t* watched, int64_t watched_stride, const float* thr_init,
                            int32_t n_chunks, int32_t* out_idx, float* out_score, int32_t* out_cnt,
                            float* out_thr, float* dump_scores, void* stream);

/* OR the CSR (indptr int64[n_rows+1], idx int32) into bit rows out[n_rows][stride_words] (caller
 * initialises `out`): the "already watched" mask of model_recs.py:144-155 for ar_cosine_topk_allpairs. */
int ar_bits_from_csr(const int64_t* indptr, const int32_t* idx, int64_t n_rows, int64_t stride_words,
                     int64_t n_bits, uint32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANIMEREC_H_ */
