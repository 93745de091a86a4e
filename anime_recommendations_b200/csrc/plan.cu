// Per-step dedup plan: group the samples of each training step by embedding row.
//
// Replaces TF's IndexedSlices -> UnsortedSegmentSum gradient aggregation with a
// sort-based, atomic-free plan: one CTA sorts one step's (row, sample) pairs entirely in shared
// memory (bitonic network over 64-bit packed keys), then emits the distinct rows, the segment
// offsets and the list of "heavy" rows.  Many steps are planned per launch (one CTA each), so
// the cost is amortised off the per-step critical path.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

namespace ar {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int kPlanThreads = 256;   // 256 x <=64 registers = 16 K: a plan CTA fits on an SM beside the persistent step kernel
constexpr int kPlanWarps = kPlanThreads / 32;
constexpr int kRadixBits = 4, kRadix = 1 << kRadixBits;

// One CTA sorts one step's samples by row, STABLY (samples of a row keep their order), with an LSD radix sort of
// the sample permutation in shared memory: 4-bit digits, every thread owns a contiguous run of positions (blocked
// arrangement => stable), per-thread digit counts -> one block-wide exclusive scan in digit-major order -> scatter.
// ceil(bits(max row)/4) passes of O(n) instead of the 105 compare-exchange stages of a 16 K bitonic network (the
// round-1 kernel: 400 us per step-plan at 512 threads, 9 % of a training chunk; this one: see profiles/).
// Dynamic shared memory: rows[batch] u32 | perm A[batch] u16 | perm B[batch] u16 | counts[16][512] u16.
__global__ void __launch_bounds__(kPlanThreads, 4)
plan_build_kernel(const int32_t* __restrict__ idx, int64_t n_total, int batch, int64_t step0,
                  ar_plan plan, const int32_t* __restrict__ counts) {
  extern __shared__ unsigned char smem_raw[];
  uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);
  uint16_t* perm_a = reinterpret_cast<uint16_t*>(rows + batch);
  uint16_t* perm_b = perm_a + batch + (batch & 1);
  uint16_t* cnt = perm_b + batch + (batch & 1);            // [kRadix][kPlanThreads]
  __shared__ int warp_tot[kPlanWarps];
  __shared__ int n_heavy_s, n_uniq_s;
  __shared__ uint32_t max_row_s;
  const int slot = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int64_t base = (step0 + slot) * (int64_t)batch;
  int n = 0;
  if (counts) {  // list form (ar_plan_build_lists): slot's keys at idx[slot*batch ...], counts[slot] of them
    base = (int64_t)slot * batch;
    n = min(batch, counts[slot]);
  } else if (base < n_total) {
    n = (int)min((int64_t)batch, n_total - base);
  }
  if (tid == 0) {
    n_heavy_s = 0;
    max_row_s = 0u;
  }
  __syncthreads();
  uint32_t mx = 0u;
  for (int i = tid; i < n; i += kPlanThreads) {
    const uint32_t r = (uint32_t)idx[base + i];
    rows[i] = r;
    perm_a[i] = (uint16_t)i;
    mx = max(mx, r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(&max_row_s, mx);
  __syncthreads();
  const int bits = 32 - __clz(max_row_s | 1u);
  const int per = (n + kPlanThreads - 1) / kPlanThreads;     // positions per thread
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  uint16_t* pin = perm_a;
  uint16_t* pout = perm_b;
  for (int shift = 0; shift < bits; shift += kRadixBits) {
    int c[kRadix];
#pragma unroll
    for (int d = 0; d < kRadix; ++d) c[d] = 0;
    for (int i = lo; i < hi; ++i) {
      const int dgt = (rows[pin[i]] >> shift) & (kRadix - 1);
#pragma unroll
      for (int d = 0; d < kRadix; ++d) c[d] += (dgt == d) ? 1 : 0;
    }
#pragma unroll
    for (int d = 0; d < kRadix; ++d) cnt[d * kPlanThreads + tid] = (uint16_t)c[d];
    __syncthreads();
    // exclusive scan of the kRadix*kPlanThreads counters in digit-major order: thread t owns 16 consecutive ones
    int own[kRadix];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < kRadix; ++j) {
      own[j] = cnt[tid * kRadix + j];
      sum += own[j];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (tid < 32) {
      const int w = tid < kPlanWarps ? warp_tot[tid] : 0;
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (tid >= o) wi += t;
      }
      if (tid < kPlanWarps) warp_tot[tid] = wi - w;  // exclusive
    }
    __syncthreads();
    int run = warp_tot[wid] + incl - sum;
#pragma unroll
    for (int j = 0; j < kRadix; ++j) {
      cnt[tid * kRadix + j] = (uint16_t)run;     // n <= 16384 fits 16 bits
      run += own[j];
    }
    __syncthreads();
    int off[kRadix];
#pragma unroll
    for (int d = 0; d < kRadix; ++d) off[d] = cnt[d * kPlanThreads + tid];
    for (int i = lo; i < hi; ++i) {
      const uint16_t item = pin[i];
      const int dgt = (rows[item] >> shift) & (kRadix - 1);
      int pos = 0;
#pragma unroll
      for (int d = 0; d < kRadix; ++d)
        if (dgt == d) pos = off[d]++;
      pout[pos] = item;
    }
    __syncthreads();
    uint16_t* t = pin;
    pin = pout;
    pout = t;
  }
  // pin: samples sorted by (row, sample)

  int32_t* order = plan.order + (int64_t)slot * plan.batch_cap;
  int32_t* uniq = plan.uniq + (int64_t)slot * plan.batch_cap;
  int32_t* off_o = plan.off + (int64_t)slot * (plan.batch_cap + 1);
  int32_t* meta = plan.meta + (int64_t)slot * 4;
  int32_t* heavy = plan.heavy + (int64_t)slot * plan.heavy_cap;

  // segment heads over the same blocked arrangement
  int cnt_heads = 0;
  for (int i = lo; i < hi; ++i) {
    const uint32_t r = rows[pin[i]];
    cnt_heads += (i == 0 || rows[pin[i - 1]] != r) ? 1 : 0;
  }
  int incl = cnt_heads;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (tid < 32) {
    const int w = tid < kPlanWarps ? warp_tot[tid] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += t;
    }
    if (tid < kPlanWarps) warp_tot[tid] = wi - w;  // exclusive
  }
  __syncthreads();
  int seg = warp_tot[wid] + incl - cnt_heads;
  for (int i = lo; i < hi; ++i) {
    const int item = pin[i];
    const uint32_t r = rows[item];
    order[i] = item;
    if (i == 0 || rows[pin[i - 1]] != r) {
      uniq[seg] = (int32_t)r;
      off_o[seg] = i;
      ++seg;
    }
  }
  if (tid == kPlanThreads - 1) {
    n_uniq_s = seg;  // the last thread's running count == total
    off_o[seg] = n;
  }
  __syncthreads();
  const int n_uniq = n_uniq_s;
  __threadfence_block();
  for (int sg = tid; sg < n_uniq; sg += kPlanThreads) {
    const int len = off_o[sg + 1] - off_o[sg];
    if (len > AR_HEAVY_LEN) {
      const int pos = atomicAdd(&n_heavy_s, 1);
      if (pos < plan.heavy_cap) heavy[pos] = sg;
    }
  }
  __syncthreads();
  if (tid == 0) {
    meta[0] = n_uniq;
    meta[1] = min(n_heavy_s, plan.heavy_cap);
    meta[2] = n;
    meta[3] = 0;
  }
}

// in_prev[slot][seg] = 1 iff distinct row `seg` of step `slot` is also a distinct row of step slot-1 in any of
// the n_lists lists (this rank's own plan, or every rank's all-gathered plans).  One CTA per slot >= 1.
__global__ void __launch_bounds__(256)
plan_link_kernel(ar_plan plan, const int32_t* __restrict__ uniq_lists, const int32_t* __restrict__ meta_lists,
                 int n_lists, int64_t uniq_stride, int64_t meta_stride) {
  const int slot = blockIdx.x + 1;
  const int B = plan.batch_cap;
  const int32_t* uniq = plan.uniq + (int64_t)slot * B;
  const int n = plan.meta[(int64_t)slot * 4];
  uint8_t* out = plan.in_prev + (int64_t)slot * B;
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    const int row = uniq[s];
    bool hit = false;
    for (int l = 0; l < n_lists && !hit; ++l) {
      const int32_t* prev = uniq_lists + l * uniq_stride + (int64_t)(slot - 1) * B;
      const int np = meta_lists[l * meta_stride + (int64_t)(slot - 1) * 4];
      int lo = 0, hi = np;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (prev[mid] < row) lo = mid + 1; else hi = mid;
      }
      hit = lo < np && prev[lo] == row;
    }
    out[s] = hit ? 1 : 0;
  }
}


// ---------------------------------------------------------------------------------------------
// Replay schedule (AR_ADAM_REPLAY): how long ago was every distinct row of every planned step last touched?
// The answer depends only on the index stream, so it is computed here, at plan time and off the step's critical
// path, instead of by a per-step classify launch.  `seen[row]` = global step (1-based) of the row's latest
// PLANNED touch; it is carried from chunk to chunk by the caller.
//
// plan_bounds_kernel: the walk below splits the row range into gridDim.x contiguous parts; bounds[s][j] = first
// segment of step s whose row is >= part j's first row (lower bound in the step's ascending distinct-row list).
__global__ void __launch_bounds__(256)
plan_bounds_kernel(ar_plan plan, int n_parts, int rows_per_part, int32_t* __restrict__ bounds) {
  const int slot = blockIdx.x;
  const int32_t* uniq = plan.uniq + (int64_t)slot * plan.batch_cap;
  const int n = plan.meta[(int64_t)slot * 4];
  for (int j = threadIdx.x; j <= n_parts; j += blockDim.x) {
    const int64_t first = (int64_t)j * rows_per_part;
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)uniq[mid] < first) lo = mid + 1; else hi = mid;
    }
    bounds[(int64_t)slot * (n_parts + 1) + j] = lo;
  }
}

// plan_gap_kernel: CTA j owns rows [j*rows_per_part, (j+1)*rows_per_part) and walks the steps in order with its
// slice of `seen` in shared memory: gap[s][seg] = t(s) - max(seen[row], t_flush), then seen[row] = t(s).
__global__ void __launch_bounds__(128)
plan_gap_kernel(ar_plan plan, int n_steps, int64_t t0, int64_t t_flush, int32_t* __restrict__ seen, int n_rows,
                int rows_per_part, const int32_t* __restrict__ bounds, int32_t* __restrict__ gap) {
  extern __shared__ int32_t seen_s[];
  const int part = blockIdx.x, n_parts = gridDim.x;
  const int64_t first = (int64_t)part * rows_per_part;
  const int mine = (int)max((int64_t)0, min((int64_t)rows_per_part, (int64_t)n_rows - first));
  for (int i = threadIdx.x; i < mine; i += blockDim.x) seen_s[i] = seen[first + i];
  __syncthreads();
  const int32_t tf = (int32_t)t_flush;
  for (int s = 0; s < n_steps; ++s) {
    const int lo = bounds[(int64_t)s * (n_parts + 1) + part], hi = bounds[(int64_t)s * (n_parts + 1) + part + 1];
    const int32_t t = (int32_t)(t0 + s + 1);
    const int32_t* uniq = plan.uniq + (int64_t)s * plan.batch_cap;
    int32_t* g = gap + (int64_t)s * plan.batch_cap;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {   // rows of one step are distinct: no conflicts
      const int r = uniq[i] - (int)first;
      g[i] = t - max(seen_s[r], tf);
      seen_s[r] = t;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < mine; i += blockDim.x) seen[first + i] = seen_s[i];
}

// plan_sched_kernel: one CTA per step.  Every distinct row with gap >= 2 becomes an item (animerec.h: ar_sched):
// sublist k = min(gap-1, depth, slot), longest replay first inside a sublist (log2 buckets: counting sort over
// (k, clz(gap)) bins).  Rows with gap >= AR_SCHED_SPLIT_GAP become `nparts` items of 32 elements each when the
// slot's capacity allows.  Also zeroes the slot's run-time cursors.
constexpr int kSchedBins = (AR_SCHED_MAX_DEPTH + 1) * 32;
constexpr int kLongClz = 23;   // clz(gap) <= 23  <=>  gap >= 256 = AR_SCHED_SPLIT_GAP
static_assert((1 << (31 - kLongClz)) == AR_SCHED_SPLIT_GAP, "split threshold and its clz bucket disagree");
__global__ void __launch_bounds__(512)
plan_sched_kernel(ar_plan pu, ar_plan pa, const int32_t* __restrict__ gap_u, const int32_t* __restrict__ gap_a,
                  int depth, int nparts, ar_sched sc) {
  __shared__ int cnt[kSchedBins], base[kSchedBins];
  __shared__ int split_s;
  const int slot = blockIdx.x;
  const int nu = pu.meta[(int64_t)slot * 4], na = pa.meta[(int64_t)slot * 4];
  for (int i = threadIdx.x; i < kSchedBins; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  auto bin_of = [&](int g) { return g <= 1 ? -1 : min(min(g - 1, depth), slot) * 32 + __clz(g); };
  for (int i = threadIdx.x; i < nu + na; i += blockDim.x) {
    const int g = i < nu ? gap_u[(int64_t)slot * pu.batch_cap + i] : gap_a[(int64_t)slot * pa.batch_cap + (i - nu)];
    const int b = bin_of(g);
    if (b >= 0) atomicAdd(&cnt[b], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int rows = 0, n_long = 0;
    for (int b = 0; b < kSchedBins; ++b) {
      rows += cnt[b];
      if ((b & 31) <= kLongClz) n_long += cnt[b];
    }
    const bool split = nparts > 1 && (int64_t)rows + (int64_t)n_long * (nparts - 1) <= (int64_t)sc.cap;
    split_s = split ? 1 : 0;
    int32_t* sub = sc.sub + (int64_t)slot * AR_SCHED_SUB;
    int32_t* cur = sc.cursor + (int64_t)slot * AR_SCHED_SUB;
    int acc = 0;
    for (int b = 0; b < kSchedBins; ++b) {
      if ((b & 31) == 0) sub[b >> 5] = acc;
      base[b] = acc;
      acc += cnt[b] * ((split && (b & 31) <= kLongClz) ? nparts : 1);
    }
    sub[AR_SCHED_MAX_DEPTH + 1] = acc;
    for (int j = 0; j < AR_SCHED_SUB; ++j) cur[j] = 0;
  }
  __syncthreads();
  const bool split = split_s != 0;
  int32_t* codes = sc.codes + (int64_t)slot * sc.cap;
  int32_t* glen = sc.glen + (int64_t)slot * sc.cap;
  for (int i = threadIdx.x; i < nu + na; i += blockDim.x) {
    const bool second = i >= nu;
    const int g = second ? gap_a[(int64_t)slot * pa.batch_cap + (i - nu)] : gap_u[(int64_t)slot * pu.batch_cap + i];
    const int b = bin_of(g);
    if (b < 0) continue;
    const int row = second ? pa.uniq[(int64_t)slot * pa.batch_cap + (i - nu)] : pu.uniq[(int64_t)slot * pu.batch_cap + i];
    const int code = row | (second ? (int)0x80000000 : 0);
    const int pos = atomicAdd(&cnt[b], -1) - 1;
    if (split && (b & 31) <= kLongClz) {
      for (int p = 0; p < nparts; ++p) {
        codes[base[b] + pos * nparts + p] = code | (1 << 30) | (p << 26);
        glen[base[b] + pos * nparts + p] = g;
      }
    } else {
      codes[base[b] + pos] = code;
      glen[base[b] + pos] = g;
    }
  }
}

}  // namespace ar

extern "C" int ar_plan_link(const ar_plan* plan, int32_t n_steps, const int32_t* uniq_all, const int32_t* meta_all,
                            int32_t n_ranks, void* stream) {
  AR_REQUIRE(plan && plan->in_prev, "ar_plan_link: plan without in_prev");
  AR_REQUIRE(n_steps >= 0 && n_steps <= plan->n_slots, "ar_plan_link: n_steps %d > plan.n_slots %d", n_steps, plan->n_slots);
  AR_REQUIRE((uniq_all == nullptr) == (meta_all == nullptr) && n_ranks >= 1, "ar_plan_link: bad list arguments");
  AR_CUDA(cudaMemsetAsync(plan->in_prev, 0, (size_t)plan->batch_cap, (cudaStream_t)stream));  // slot 0: nothing known
  if (n_steps <= 1) return AR_OK;
  const int32_t* ul = uniq_all ? uniq_all : plan->uniq;
  const int32_t* ml = meta_all ? meta_all : plan->meta;
  const int nl = uniq_all ? n_ranks : 1;
  ar::plan_link_kernel<<<n_steps - 1, 256, 0, (cudaStream_t)stream>>>(*plan, ul, ml, nl, (int64_t)plan->n_slots * plan->batch_cap,
                                                                       (int64_t)plan->n_slots * 4);
  AR_LAUNCH_CHECK();
  return AR_OK;
}


extern "C" int ar_plan_sched(const ar_plan* plan_u, const ar_plan* plan_a, int32_t n_steps, int64_t t0, int64_t t_flush,
                             int32_t* seen_u, int32_t n_rows_u, int32_t* seen_a, int32_t n_rows_a, int32_t depth,
                             int32_t dim, const ar_sched* sched, void* stream) {
  AR_REQUIRE(plan_u && plan_a && sched && seen_u && seen_a, "ar_plan_sched: null pointer");
  AR_REQUIRE(sched->codes && sched->glen && sched->sub && sched->cursor && sched->gap_u && sched->gap_a && sched->bounds,
             "ar_plan_sched: null buffer in sched");
  AR_REQUIRE(n_steps >= 0 && n_steps <= plan_u->n_slots && n_steps <= plan_a->n_slots && n_steps <= sched->n_slots,
             "ar_plan_sched: n_steps %d exceeds the plans / schedule", n_steps);
  AR_REQUIRE(sched->cap >= plan_u->batch_cap + plan_a->batch_cap, "ar_plan_sched: sched.cap too small");
  AR_REQUIRE(depth >= 1 && depth <= AR_SCHED_MAX_DEPTH, "ar_plan_sched: depth %d outside [1, %d]", depth, AR_SCHED_MAX_DEPTH);
  AR_REQUIRE(dim > 0 && dim <= 512, "ar_plan_sched: dim %d unsupported", dim);
  AR_REQUIRE(n_rows_u <= AR_SCHED_MAX_ROWS && n_rows_a <= AR_SCHED_MAX_ROWS, "ar_plan_sched: more than %d rows", AR_SCHED_MAX_ROWS);
  AR_REQUIRE(t0 + n_steps < (1LL << 27), "ar_plan_sched: step counter exceeds 2^27 (the top bits of last_step count split-row parts)");
  if (n_steps == 0) return AR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    AR_CUDA(cudaFuncSetAttribute(ar::plan_gap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const ar_plan* plans[2] = {plan_u, plan_a};
  int32_t* seen[2] = {seen_u, seen_a};
  const int n_rows[2] = {n_rows_u, n_rows_a};
  int32_t* gaps[2] = {sched->gap_u, sched->gap_a};
  for (int w = 0; w < 2; ++w) {
    // parts: enough CTAs to fill the GPU, slices that fit shared memory
    int n_parts = AR_SCHED_PARTS;
    int rows_per_part = ar::ceil_div(std::max(1, n_rows[w]), n_parts);
    AR_REQUIRE((size_t)rows_per_part * 4 <= 200 * 1024, "ar_plan_sched: table of %d rows too large for %d parts", n_rows[w], n_parts);
    ar::plan_bounds_kernel<<<n_steps, 256, 0, st>>>(*plans[w], n_parts, rows_per_part, sched->bounds);
    AR_LAUNCH_CHECK();
    ar::plan_gap_kernel<<<n_parts, 128, (size_t)rows_per_part * 4, st>>>(*plans[w], n_steps, t0, t_flush, seen[w], n_rows[w],
                                                                        rows_per_part, sched->bounds, gaps[w]);
    AR_LAUNCH_CHECK();
  }
  ar::plan_sched_kernel<<<n_steps, 512, 0, st>>>(*plan_u, *plan_a, sched->gap_u, sched->gap_a, depth, ar::ceil_div(dim, 32), *sched);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" const char* ar_last_error(void) { return ar::g_err; }
extern "C" int ar_abi_version(void) { return 15; }

extern "C" int ar_check_device(void) {
  int dev = 0;
  AR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  AR_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) {
    ar::set_error("libanimerec is built for sm_100a only; current device is sm_%d%d", p.major, p.minor);
    return AR_ERR_UNSUPPORTED;
  }
  return AR_OK;
}

static int plan_build_impl(const int32_t* idx, int64_t n_total, int32_t batch, int64_t step0, int32_t n_steps,
                           const ar_plan* plan, const int32_t* counts, void* stream);

extern "C" int ar_plan_build(const int32_t* idx, int64_t n_total, int32_t batch, int64_t step0,
                             int32_t n_steps, const ar_plan* plan, void* stream) {
  return plan_build_impl(idx, n_total, batch, step0, n_steps, plan, nullptr, stream);
}

extern "C" int ar_plan_build_lists(const int32_t* keys, int32_t stride, const int32_t* counts, int32_t n_steps,
                                   const ar_plan* plan, void* stream) {
  AR_REQUIRE(counts, "ar_plan_build_lists: null counts");
  return plan_build_impl(keys, 0, stride, 0, n_steps, plan, counts, stream);
}

static int plan_build_impl(const int32_t* idx, int64_t n_total, int32_t batch, int64_t step0, int32_t n_steps,
                           const ar_plan* plan, const int32_t* counts, void* stream) {
  AR_REQUIRE(idx && plan, "ar_plan_build: null pointer");
  AR_REQUIRE(batch > 0 && batch <= AR_MAX_BATCH, "ar_plan_build: batch %d outside (0, %d]", batch, AR_MAX_BATCH);
  AR_REQUIRE(plan->batch_cap >= batch, "ar_plan_build: plan.batch_cap %d < batch %d", plan->batch_cap, batch);
  AR_REQUIRE(n_steps >= 0 && n_steps <= plan->n_slots, "ar_plan_build: n_steps %d > plan.n_slots %d", n_steps, plan->n_slots);
  AR_REQUIRE(plan->heavy_cap >= batch / AR_HEAVY_LEN + 1, "ar_plan_build: heavy_cap too small");
  if (n_steps == 0) return AR_OK;
  const size_t be = (size_t)batch + (batch & 1);
  const size_t smem = (size_t)batch * 4 + 2 * be * 2 + (size_t)ar::kRadix * ar::kPlanThreads * 2;
  static bool attr_set = false;
  if (!attr_set) {
    AR_CUDA(cudaFuncSetAttribute(ar::plan_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AR_MAX_BATCH * 8 + ar::kRadix * ar::kPlanThreads * 2));
    attr_set = true;
  }
  ar::plan_build_kernel<<<n_steps, ar::kPlanThreads, smem, (cudaStream_t)stream>>>(idx, n_total, batch, step0, *plan, counts);
  AR_LAUNCH_CHECK();
  return AR_OK;
}
