// Half A: the training step of the neural_network.py embedding model on sm_100a.
//
// Single GPU (ar_train_steps): one persistent kernel per chunk of steps, chunk.inl.
// The multi-GPU paths (dist.inl, shard.inl, peer.inl) chain the stage kernels of this file per step:
//   rows_classify / rows_catchup  (AR_ADAM_REPLAY) bring the step's distinct rows to optimizer step t-1 by replaying
//                      their missed pure-L2 Adam steps in registers
//   embed_fwd          warp per sample: gather both rows (128-bit loads), l2-normalise, dot
//   head_step          Dense(1) + BatchNorm(train) + sigmoid + BCE, dLoss/dy per sample, Adam on the 4 head scalars,
//                      moving statistics, metrics -- ticketed last CTA, fixed summation order
//   rows_update        warp per distinct row: atomic-free segment reduction of the row gradient over the plan's
//                      sorted samples + L2 term + Adam, one RMW of (W, m, v)
// AR_ADAM_DENSE appends a flush of every other row to step t (the reference-literal dense Adam).
//
// Arithmetic follows oracle/train.py (the restatement of neural_network.py:66-106 under
// Keras-2.12 semantics); citations there.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ar {

constexpr int kRowThreads = 256;               // 8 warps per CTA, one row / sample per warp
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kHeadThreads = 256;

// ---------------------------------------------------------------------------------------------
// Row register tile: a warp owns one row, lane l holds float4 #(l + 32*k), k < NV.
template <int NV>
struct RowTile {
  float4 x[NV];
  __device__ __forceinline__ void load(const float* row, int d4, int lane) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int j = lane + 32 * k;
      x[k] = (j < d4) ? ld4(row + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ void store(float* row, int d4, int lane) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int j = lane + 32 * k;
      if (j < d4) st4(row + 4 * j, x[k]);
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

template <int NV>
__device__ __forceinline__ float tile_dot(const RowTile<NV>& a, const RowTile<NV>& b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) s += dot4(a.x[k], b.x[k]);
  return warp_sum(s);
}

// L2-regulariser accumulator of the reported loss (animerec.h: ar_train_ctx.reg_acc).
struct RegAcc {
  unsigned long long* acc;
  const float* stepw;
  float scale;
};
// One fixed-point add per row visit: integer addition is associative, so the total does not depend on the order
// in which warps arrive.
__device__ __forceinline__ void reg_commit(double regd, const RegAcc& reg, int lane) {
  if (!reg.acc) return;
  regd = warp_sum(regd);
  if (lane == 0 && regd != 0.0) atomicAdd(reg.acc, (unsigned long long)__double2ll_rn(regd * (double)reg.scale));
}

// Replay pure-L2 Adam steps (from, to] of one row held in registers.  regd += sum over the replayed
// steps t of stepw[t] * (this lane's share of ||w before step t||^2).
template <int NV>
__device__ __forceinline__ void replay_l2(RowTile<NV>& w, RowTile<NV>& m, RowTile<NV>& v,
                                          const float* __restrict__ alpha, const float* __restrict__ stepw,
                                          int64_t from, int64_t to, float l2x2, int lane, double& regd) {
  // both loops stay ROLLED: unrolled by the compiler the body was 20 KB of code per instantiation for no gain
#pragma unroll 1
  for (int64_t t0 = from + 1; t0 <= to; t0 += 32) {
    int64_t tl = t0 + lane;
    float a_l = (tl <= to) ? __ldg(alpha + tl) : 0.f;
    float w_l = (stepw && tl <= to) ? __ldg(stepw + tl) : 0.f;
    int cnt = (int)min((int64_t)32, to - t0 + 1);
    float accf = 0.f;
#pragma unroll 1
    for (int s = 0; s < cnt; ++s) {
      float a = __shfl_sync(0xffffffffu, a_l, s);
      float sw = __shfl_sync(0xffffffffu, w_l, s);
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        p = fmaf(w.x[k].x, w.x[k].x, p);
        p = fmaf(w.x[k].y, w.x[k].y, p);
        p = fmaf(w.x[k].z, w.x[k].z, p);
        p = fmaf(w.x[k].w, w.x[k].w, p);
        float4 g = make_float4(__fmul_rn(l2x2, w.x[k].x), __fmul_rn(l2x2, w.x[k].y),
                               __fmul_rn(l2x2, w.x[k].z), __fmul_rn(l2x2, w.x[k].w));
        adam4(w.x[k], m.x[k], v.x[k], g, a);
      }
      accf = fmaf(sw, p, accf);
    }
    regd += (double)accf;
  }
}

// ---------------------------------------------------------------------------------------------
// 1. catch-up of the step's distinct rows (both tables in one launch)
struct CatchupArgs {
  ar_table tab[2];
  const int32_t* uniq[2];
  const int32_t* meta[2];
  int blocks0;  // CTAs assigned to table 0
  // look-ahead catch-up: distinct rows flagged here are ALSO touched by the previous step (of any rank) and
  // are skipped -- that step's own update brings them up to date, and may be running concurrently.  The
  // flags come from ar_plan_link (per chunk, off the critical path), so the check is one byte load
  const uint8_t* skip_flag[2];
  // longest-first schedule (ar_train_ctx.sched_ws): three buckets of (table << 31 | row) by replay length,
  // bucket b at sched + b*cap, their fill counts at sched + 3*cap
  int32_t* sched;
  int cap;
  RegAcc reg;
};
constexpr int kLongReplay = 128, kMidReplay = 32;

// Which rows need how much replay?  Thread per distinct row of the step: rows the skip list covers or that
// are already current drop out, the rest go to the long / mid / short bucket (warp-aggregated append).
// Replay lengths are geometric (mean n_rows / distinct-per-step, tail ~10x that); launching the catch-up in
// bucket order starts the long rows first, so they no longer form the kernel's tail.
__global__ void __launch_bounds__(256)
rows_classify_kernel(CatchupArgs a, int64_t t_target, int n0_cap, int n1_cap) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const bool second = g >= n0_cap;
  const int seg = second ? g - n0_cap : g;
  int bucket = -1, row = 0;
  if (seg < (second ? n1_cap : n0_cap) && seg < (second ? a.meta[1] : a.meta[0])[0]) {
    row = (second ? a.uniq[1] : a.uniq[0])[seg];
    const uint8_t* __restrict__ flag = second ? a.skip_flag[1] : a.skip_flag[0];
    if (!(flag && flag[seg])) {
      const int64_t len = t_target - (int64_t)(second ? a.tab[1].last_step : a.tab[0].last_step)[row];
      if (len > 0) bucket = len > kLongReplay ? 0 : (len > kMidReplay ? 1 : 2);
    }
  }
  int32_t* counts = a.sched + 3 * (size_t)a.cap;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const unsigned m = __ballot_sync(0xffffffffu, bucket == b);
    if (!m) continue;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counts + b, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (bucket == b) a.sched[(size_t)b * a.cap + base + __popc(m & ((1u << lane) - 1))] = row | (second ? (int)0x80000000 : 0);
  }
}

// One row per CTA, ONE ELEMENT PER LANE (ceil(dim/32) warps).  A row's replay is a serial chain per element
// (sqrt -> add -> rcp -> fma, ~65 cycles a step) and the longest row of a step (gap ~ 10x the mean) sets the
// kernel's tail: with four elements per lane ptxas issues the four chains back to back (4x the latency per
// step: 62 us per launch measured, the tail of one 350-step row), whatever the source order.  One element per
// lane leaves no intra-warp ILP to lose; the SFU is kept busy by the ~50 resident warps of other rows instead.
// Per element-step: 1 SHFL (alpha) + 1 FFMA (regulariser term) + 8 FP32 + 2 MUFU.
__device__ __forceinline__ void replay_l2_lane(float& w, float& m, float& v, const float* __restrict__ alpha,
                                               const float* __restrict__ stepw, int64_t from, int64_t to,
                                               float l2x2, int lane, double& regd) {
#pragma unroll 1
  for (int64_t t0 = from + 1; t0 <= to; t0 += 32) {
    const int64_t tl = t0 + lane;
    const float a_l = (tl <= to) ? __ldg(alpha + tl) : 0.f;
    const float w_l = (stepw && tl <= to) ? __ldg(stepw + tl) : 1.f;
    const int cnt = (int)min((int64_t)32, to - t0 + 1);
    float accf = 0.f;
    if (__all_sync(0xffffffffu, w_l == 1.f)) {  // every step of the block has full weight (all but an epoch's last)
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

__global__ void __launch_bounds__(512)
rows_catchup_kernel(CatchupArgs a, const float* __restrict__ alpha, float l2x2, int64_t t_target) {
  __shared__ double red[16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int unit = blockIdx.x;
  bool second;
  int row;
  if (a.sched) {  // bucket order: long rows first
    const int32_t* counts = a.sched + 3 * (size_t)a.cap;
    const int n0 = counts[0], n1 = counts[1], n2 = counts[2];
    int b = unit, code;
    if (b < n0) code = a.sched[b];
    else if ((b -= n0) < n1) code = a.sched[(size_t)a.cap + b];
    else if ((b -= n1) < n2) code = a.sched[2 * (size_t)a.cap + b];
    else return;
    second = code < 0;
    row = code & 0x7fffffff;
  } else {
    second = unit >= a.blocks0;
    const int seg = second ? unit - a.blocks0 : unit;
    const int32_t* meta = second ? a.meta[1] : a.meta[0];
    if (seg >= (second ? a.cap - a.blocks0 : a.blocks0) || seg >= meta[0]) return;
    row = (second ? a.uniq[1] : a.uniq[0])[seg];
    const uint8_t* __restrict__ flag = second ? a.skip_flag[1] : a.skip_flag[0];
    if (flag && flag[seg]) return;
  }
  const int dim = a.tab[0].dim;
  float* __restrict__ W = second ? a.tab[1].W : a.tab[0].W;
  float* __restrict__ M = second ? a.tab[1].m : a.tab[0].m;
  float* __restrict__ V = second ? a.tab[1].v : a.tab[0].v;
  int32_t* __restrict__ last_step = second ? a.tab[1].last_step : a.tab[0].last_step;
  const int64_t last = last_step[row];
  if (last >= t_target) return;  // uniform over the CTA
  const int e = threadIdx.x;
  const bool live = e < dim;
  const size_t o = (size_t)row * dim + e;
  float w = 0.f, m = 0.f, v = 0.f;
  if (live) {
    w = W[o];
    m = M[o];
    v = V[o];
  }
  double regd = 0.0;
  replay_l2_lane(w, m, v, alpha, a.reg.stepw, last, t_target, l2x2, lane, regd);
  if (live) {
    W[o] = w;
    M[o] = m;
    V[o] = v;
  }
  if (a.reg.acc) {
    regd = warp_sum(regd);
    if (lane == 0) red[wid] = regd;
  }
  __syncthreads();  // every warp of the row has read last_step[row]
  if (threadIdx.x == 0) {
    last_step[row] = (int32_t)t_target;
    if (a.reg.acc) {
      double tot = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
      if (tot != 0.0) atomicAdd(a.reg.acc, (unsigned long long)__double2ll_rn(tot * (double)a.reg.scale));
    }
  }
}

// whole-table flush: every row to t_target
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
table_flush_kernel(ar_table tb, const float* __restrict__ alpha, float l2x2, int64_t t_target, RegAcc reg) {
  const int lane = threadIdx.x & 31;
  const int d4 = tb.dim >> 2;
  double regd = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5); row < tb.n_rows;
       row += (int64_t)gridDim.x * kRowWarps) {
    const int64_t last = tb.last_step[row];
    if (last >= t_target) continue;
    const size_t o = (size_t)row * tb.dim;
    RowTile<NV> w, m, v;
    w.load(tb.W + o, d4, lane);
    m.load(tb.m + o, d4, lane);
    v.load(tb.v + o, d4, lane);
    replay_l2<NV>(w, m, v, alpha, reg.stepw, last, t_target, l2x2, lane, regd);
    w.store(tb.W + o, d4, lane);
    m.store(tb.m + o, d4, lane);
    v.store(tb.v + o, d4, lane);
    if (lane == 0) tb.last_step[row] = (int32_t)t_target;
  }
  reg_commit(regd, reg, lane);
}

// ---------------------------------------------------------------------------------------------
// 2. forward: gather + l2_normalize + dot (neural_network.py:75-96; oracle forward())
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
embed_fwd_kernel(const float* __restrict__ U, const float* __restrict__ A, int dim,
                 const int32_t* __restrict__ iu, const int32_t* __restrict__ ia, int n,
                 const int32_t* __restrict__ meta_n, float* __restrict__ uh, float* __restrict__ ah,
                 float* __restrict__ c, float* __restrict__ ru, float* __restrict__ ra,
                 double* __restrict__ fwd_part) {
  __shared__ float cs_s[kRowWarps];
  if (meta_n) n = min(n, meta_n[2]);
  const int s = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (fwd_part) {  // (sum c, sum c^2) of this CTA's samples, summed in a fixed order by one thread
    if (lane == 0) cs_s[threadIdx.x >> 5] = 0.f;
  }
  if (s < n) {
  const int d4 = dim >> 2;
  RowTile<NV> u, a;
  u.load(U + (size_t)iu[s] * dim, d4, lane);
  a.load(A + (size_t)ia[s] * dim, d4, lane);
  const float su = tile_dot<NV>(u, u);
  const float sa = tile_dot<NV>(a, a);
  const float r_u = 1.0f / sqrtf(fmaxf(su, kL2NormEps));
  const float r_a = 1.0f / sqrtf(fmaxf(sa, kL2NormEps));
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    u.x[k] = scale4(u.x[k], r_u);
    a.x[k] = scale4(a.x[k], r_a);
  }
  const float cs = tile_dot<NV>(u, a);
  if (uh) u.store(uh + (size_t)s * dim, d4, lane);
  if (ah) a.store(ah + (size_t)s * dim, d4, lane);
  if (lane == 0) {
    c[s] = cs;
    if (ru) ru[s] = r_u;
    if (ra) ra[s] = r_a;
    cs_s[threadIdx.x >> 5] = cs;
  }
  }
  if (fwd_part) {
    __syncthreads();
    if (threadIdx.x == 0) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int i = 0; i < kRowWarps; ++i) {
        const double x = (double)cs_s[i];
        a0 += x;
        a1 += x * x;
      }
      fwd_part[2 * blockIdx.x] = a0;
      fwd_part[2 * blockIdx.x + 1] = a1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 3. head: Dense(1) -> BatchNorm(train) -> sigmoid -> BCE, backward, Adam on (w,b,gamma,beta)
//    (neural_network.py:97-104; oracle forward()/head_backward()).  One CTA; sums in double.
template <int NVAL>
__device__ __forceinline__ void block_sum(double (&v)[NVAL], double* smem /* [NVAL][32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NVAL; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect smem reuse between consecutive calls
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NVAL; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NVAL; ++i) {
    double x = (lane < nw) ? smem[i * 32 + lane] : 0.0;
    v[i] = warp_sum(x);
  }
}

// numerically stable sigmoid without a branch: e = exp(-|y|) <= 1;  y >= 0: 1/(1+e),  y < 0: e/(1+e)
__device__ __forceinline__ float sigmoidf_(float y) {
  const float e = expf(-fabsf(y));
  const float r = 1.0f / (1.0f + e);
  return y >= 0.f ? r : e * r;
}
__device__ __forceinline__ float bce_logits(float y, float t) {
  return fmaxf(y, 0.f) - y * t + log1pf(expf(-fabsf(y)));
}

// Per-step scalars the head hands to the row-update kernel: dc_s = K_COEF*(dy_s - K_S1N - zh_s*K_S2N),
// zh_s = ((w*c_s + b) - mu)*inv   (oracle head_backward(): dz, dc)
enum { K_COEF = 0, K_S1N, K_S2N, K_MU, K_INV, K_W, K_B, K_GAMMA, K_BETA, K_FN, K_RN, K_STEPC };  // 16 floats reserved
constexpr int kHeadSums = 7;  // sum bce, sq err, dy, dy*zh, zh, dy*c, zh*c

__device__ __forceinline__ float dc_of(float dy, float c, const float* __restrict__ k) {
  const float zh = ((k[K_W] * c + k[K_B]) - k[K_MU]) * k[K_INV];
  return k[K_COEF] * (dy - k[K_S1N] - zh * k[K_S2N]);
}
// dLoss/dy of one sample from its cosine and label: the SAME expression sequence wherever it is evaluated (the
// head's sums and the row update's per-sample factor must see identical bits)
__device__ __forceinline__ float dy_of(float c, float tg, float w, float b, float mu, float inv, float gamma,
                                       float beta, float rn, float* zh_out) {
  const float zh = ((w * c + b) - mu) * inv;
  const float y = gamma * zh + beta;
  *zh_out = zh;
  return (sigmoidf_(y) - tg) * rn;      // rn = 1/n: (p - t)/n up to one rounding
}
__device__ __forceinline__ float dc_of_label(float c, float tg, const float* __restrict__ k) {
  float zh;
  const float dy = dy_of(c, tg, k[K_W], k[K_B], k[K_MU], k[K_INV], k[K_GAMMA], k[K_BETA], k[K_RN], &zh);
  return k[K_COEF] * (dy - k[K_S1N] - zh * k[K_S2N]);
}

// The head's state and outputs (device pointers), shared by head_step_kernel and the fused forward.
struct HeadIO {
  float* head;
  float* head_m;
  float* head_v;
  float* bn_moving;
  const float* alpha;
  float* dy;
  float* stepc;
  unsigned int* ticket;
  float* metrics;   // base of the per-step metrics table (row t is written), or null
};
struct HeadScalars {
  float w, b, gamma, beta, mu, var, inv, fn, rn;
};
// batch statistics of z = w*c + b from (sum c, sum c^2)
__device__ __forceinline__ HeadScalars head_scalars(const float* __restrict__ head, double sum_c, double sum_c2, int n) {
  HeadScalars h;
  h.w = head[0]; h.b = head[1]; h.gamma = head[2]; h.beta = head[3];
  h.fn = (float)n;
  h.rn = 1.0f / h.fn;
  const double mean_c = sum_c / n;
  const double var_c = fmax(sum_c2 / n - mean_c * mean_c, 0.0);
  h.mu = (float)((double)h.w * mean_c + (double)h.b);
  h.var = (float)((double)h.w * (double)h.w * var_c);
  h.inv = 1.0f / sqrtf(h.var + kBnEps);
  return h;
}
// one sample: Dense(1) -> BN(train) -> sigmoid -> BCE; returns dLoss/dy and adds the 7 head sums
__device__ __forceinline__ float head_sample(float ci, float tg, const HeadScalars& h, double (&s1)[kHeadSums]) {
  const float zh = ((h.w * ci + h.b) - h.mu) * h.inv;
  const float y = h.gamma * zh + h.beta;
  const float p = sigmoidf_(y);
  const float dy = (p - tg) * h.rn;
  s1[0] += (double)bce_logits(y, tg);
  s1[1] += (double)((tg - p) * (tg - p));
  s1[2] += (double)dy;
  s1[3] += (double)dy * (double)zh;
  s1[4] += (double)zh;
  s1[5] += (double)dy * (double)ci;
  s1[6] += (double)zh * (double)ci;
  return dy;
}
// one thread: Adam on (w, b, gamma, beta), moving statistics, the scalars of the row update, metrics
__device__ __forceinline__ void head_finalize(const HeadScalars& h, const double (&s2)[kHeadSums], double sum_c, int n,
                                              const HeadIO& io, int64_t t) {
  const double S1 = s2[2], S2 = s2[3], Szh = s2[4], Sdyc = s2[5], Szhc = s2[6], Sc = sum_c;
  const double ig = (double)h.inv * (double)h.gamma;
  // oracle head_backward(): dgamma = sum dy*zh, dbeta = sum dy, dz_i = inv*gamma*(dy_i - S1/n - zh_i*S2/n)
  const float g[4] = {(float)(ig * (Sdyc - S1 / n * Sc - S2 / n * Szhc)),  // dw = sum dz*c
                      (float)(-ig * (S2 / n) * Szh),                       // db = sum dz (0 up to rounding)
                      (float)S2, (float)S1};
  const float a = io.alpha[t];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float th = io.head[k], m = io.head_m[k], v = io.head_v[k];
    adam1(th, m, v, g[k], a);
    io.head[k] = th;
    io.head_m[k] = m;
    io.head_v[k] = v;
  }
  const float mm = io.bn_moving[0], mv = io.bn_moving[1];
  io.bn_moving[0] = mm - (mm - h.mu) * kBnOneMinusMomentum;
  io.bn_moving[1] = mv - (mv - h.var) * kBnOneMinusMomentum;
  io.stepc[K_COEF] = (float)((double)h.w * ig);
  io.stepc[K_S1N] = (float)(S1 / n);
  io.stepc[K_S2N] = (float)(S2 / n);
  io.stepc[K_MU] = h.mu;
  io.stepc[K_INV] = h.inv;
  io.stepc[K_W] = h.w;
  io.stepc[K_B] = h.b;
  io.stepc[K_GAMMA] = h.gamma;
  io.stepc[K_BETA] = h.beta;
  io.stepc[K_FN] = h.fn;
  io.stepc[K_RN] = h.rn;
  if (io.metrics) {
    float* row = io.metrics + t * 4;
    row[0] = (float)(s2[0] / n);
    row[1] = (float)(s2[1] / n);
    row[2] = h.fn;
    row[3] = h.mu;
  }
  *io.ticket = 0u;  // ready for the next step
}

__global__ void __launch_bounds__(kHeadThreads)
head_step_kernel(const float* __restrict__ c, const float* __restrict__ label, int n,
                 const int32_t* __restrict__ meta_n, const double* __restrict__ fwd_part, HeadIO io, int64_t t,
                 double* __restrict__ head_part, int nfp_in) {
  __shared__ double red[kHeadSums * 32];
  __shared__ int is_last;
  if (meta_n) n = min(n, meta_n[2]);
  if (n <= 0) return;
  const int nblk = (n + kHeadThreads - 1) / kHeadThreads;
  if ((int)blockIdx.x >= nblk) return;
  const int tid = threadIdx.x;

  // batch statistics from the forward's per-CTA partials (fixed summation order in every CTA)
  const int nfp = nfp_in > 0 ? nfp_in : (n + kRowWarps - 1) / kRowWarps;  // peer mode pre-reduces per 1024 samples
  double s0[2] = {0.0, 0.0};
  for (int i = tid; i < nfp; i += kHeadThreads) {
    s0[0] += __ldcg(fwd_part + 2 * i);
    s0[1] += __ldcg(fwd_part + 2 * i + 1);
  }
  block_sum<2>(s0, red);
  const HeadScalars h = head_scalars(io.head, s0[0], s0[1], n);

  double s1[kHeadSums];
#pragma unroll
  for (int i = 0; i < kHeadSums; ++i) s1[i] = 0.0;
  const int i = blockIdx.x * kHeadThreads + tid;
  if (i < n) io.dy[i] = head_sample(c[i], label[i], h, s1);
  block_sum<kHeadSums>(s1, red);
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < kHeadSums; ++k) head_part[blockIdx.x * 8 + k] = s1[k];
    __threadfence();
    const unsigned int old = atomicAdd(io.ticket, 1u);
    is_last = (old == (unsigned int)(nblk - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s2[kHeadSums];
#pragma unroll
  for (int k = 0; k < kHeadSums; ++k) s2[k] = 0.0;
  for (int j = tid; j < nblk; j += kHeadThreads) {
#pragma unroll
    for (int k = 0; k < kHeadSums; ++k) s2[k] += __ldcg(head_part + j * 8 + k);
  }
  block_sum<kHeadSums>(s2, red);
  if (tid == 0) head_finalize(h, s2, s0[0], n, io, t);
}

// helpers behind the stand-alone ar_head_step entry point (unit tests): partials from c, dc from dy
__global__ void c_partials_kernel(const float* __restrict__ c, int n, double* __restrict__ fwd_part) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk * kRowWarps >= n) return;
  double a0 = 0.0, a1 = 0.0;
  for (int i = blk * kRowWarps; i < min(n, (blk + 1) * kRowWarps); ++i) {
    const double x = (double)c[i];
    a0 += x;
    a1 += x * x;
  }
  fwd_part[2 * blk] = a0;
  fwd_part[2 * blk + 1] = a1;
}
__global__ void dc_from_dy_kernel(const float* __restrict__ dy, const float* __restrict__ c, int n,
                                  const float* __restrict__ stepc, float* __restrict__ dc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dc[i] = dc_of(dy[i], c[i], stepc);
}

// ---------------------------------------------------------------------------------------------
// 4. row update: segment-reduce the row gradient over the plan's sorted samples, add the L2 term,
//    apply Adam step t.   g_r = r * (sum_s dc_s*o_s - (sum_s dc_s*c_s) * (W_r*r)) + 2*l2*W_r
//    where o_s is the normalised row of the OTHER table for sample s and r = 1/||W_r||
//    (= oracle du/da summed per row, factored).
struct UpdateArgs {
  ar_table tab[2];
  const int32_t* order[2];
  const int32_t* uniq[2];
  const int32_t* off[2];
  const int32_t* meta[2];
  const int32_t* heavy[2];
  const float* other[2];  // (batch, dim) normalised rows of the other table
  const float* rinv[2];   // (batch) 1/||row|| per sample of THIS table
  int blocks_norm[2];     // warp-per-row CTAs per table
  int blocks_heavy[2];    // CTA-per-heavy-row CTAs per table
  // multi-GPU: instead of updating, emit (row id, q, P[dim]) per segment for the gradient exchange
  int32_t* emit_ids[2];
  float* emit_q[2];
  float* emit_P[2];
  int emit_cap;           // entries per table in the emit buffers; ids beyond n_uniq are set to INT_MAX
  // row-sharded tables: segment s goes to row emit_map[s] of emit_P, rows are emit_stride (> dim) floats
  // apart and carry q in column `dim`; ids are not emitted (the owners already hold the request lists)
  const int32_t* emit_map[2];
  int emit_stride;
  // peer mode: the plan's sample ids index this rank's selection list; samp maps them to the position in the
  // GLOBAL batch that c / dy are indexed by (`other` and `rinv` stay indexed by the plan's own sample id)
  const int32_t* samp[2];
  RegAcc reg;
};

template <int NV>
__device__ __forceinline__ void finish_row(const ar_table& tb, int row, RowTile<NV>& acc, float q,
                                           float rinv, const float* __restrict__ alpha, float l2x2,
                                           int64_t t, int replay, const RegAcc& reg, int lane) {
  const int d4 = tb.dim >> 2;
  const size_t o = (size_t)row * tb.dim;
  RowTile<NV> w, m, v;
  w.load(tb.W + o, d4, lane);
  m.load(tb.m + o, d4, lane);
  v.load(tb.v + o, d4, lane);
  double regd = 0.0;
  if (replay) {
    const int64_t last = tb.last_step[row];
    if (last < t - 1) replay_l2<NV>(w, m, v, alpha, reg.stepw, last, t - 1, l2x2, lane, regd);
  }
  if (reg.acc || rinv < 0.f) {
    const float ss = tile_dot<NV>(w, w);
    if (reg.acc && lane == 0) regd += (double)(reg.stepw[t] * ss);
    if (rinv < 0.f) rinv = 1.0f / sqrtf(fmaxf(ss, kL2NormEps));  // same formula as embed_fwd
  }
  const float a = alpha[t];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float4 wk = w.x[k], ak = acc.x[k], g;
    g.x = __fadd_rn(rinv * (ak.x - q * (wk.x * rinv)), __fmul_rn(l2x2, wk.x));
    g.y = __fadd_rn(rinv * (ak.y - q * (wk.y * rinv)), __fmul_rn(l2x2, wk.y));
    g.z = __fadd_rn(rinv * (ak.z - q * (wk.z * rinv)), __fmul_rn(l2x2, wk.z));
    g.w = __fadd_rn(rinv * (ak.w - q * (wk.w * rinv)), __fmul_rn(l2x2, wk.w));
    adam4(w.x[k], m.x[k], v.x[k], g, a);
  }
  w.store(tb.W + o, d4, lane);
  m.store(tb.m + o, d4, lane);
  v.store(tb.v + o, d4, lane);
  if (lane == 0) tb.last_step[row] = (int32_t)t;
  reg_commit(regd, reg, lane);
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads)
rows_update_kernel(UpdateArgs a, const float* __restrict__ c, const float* __restrict__ dy,
                   const float* __restrict__ stepc, const float* __restrict__ alpha, float l2x2, int64_t t,
                   int replay) {
  extern __shared__ float red[];  // heavy path: [kRowWarps][dim] + [kRowWarps]
  int b = blockIdx.x;
  int which, heavy_path;
  if (b < a.blocks_norm[0]) { which = 0; heavy_path = 0; }
  else if ((b -= a.blocks_norm[0]) < a.blocks_norm[1]) { which = 1; heavy_path = 0; }
  else if ((b -= a.blocks_norm[1]) < a.blocks_heavy[0]) { which = 0; heavy_path = 1; }
  else { b -= a.blocks_heavy[0]; which = 1; heavy_path = 1; }

  ar_table tb;
  tb.dim = a.tab[0].dim;
  tb.n_rows = which ? a.tab[1].n_rows : a.tab[0].n_rows;
  tb.W = which ? a.tab[1].W : a.tab[0].W;
  tb.m = which ? a.tab[1].m : a.tab[0].m;
  tb.v = which ? a.tab[1].v : a.tab[0].v;
  tb.last_step = which ? a.tab[1].last_step : a.tab[0].last_step;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int dim = tb.dim, d4 = dim >> 2;
  const int32_t* __restrict__ order = which ? a.order[1] : a.order[0];
  const int32_t* __restrict__ off = which ? a.off[1] : a.off[0];
  const int32_t* __restrict__ uniq = which ? a.uniq[1] : a.uniq[0];
  const int32_t* __restrict__ meta = which ? a.meta[1] : a.meta[0];
  const int32_t* __restrict__ heavy = which ? a.heavy[1] : a.heavy[0];
  const float* __restrict__ other = which ? a.other[1] : a.other[0];
  const float* __restrict__ rinv = which ? a.rinv[1] : a.rinv[0];
  int32_t* __restrict__ emit_ids = which ? a.emit_ids[1] : a.emit_ids[0];
  float* __restrict__ emit_q = which ? a.emit_q[1] : a.emit_q[0];
  float* __restrict__ emit_P = which ? a.emit_P[1] : a.emit_P[0];
  const int32_t* __restrict__ emit_map = which ? a.emit_map[1] : a.emit_map[0];
  const int32_t* __restrict__ samp = which ? a.samp[1] : a.samp[0];
  float kk[K_STEPC];
#pragma unroll
  for (int i = 0; i < K_STEPC; ++i) kk[i] = stepc[i];

  if (!heavy_path) {
    const int seg = b * kRowWarps + wid;
    if (seg >= meta[0]) {
      if (emit_ids && seg < a.emit_cap && lane == 0) emit_ids[seg] = 0x7fffffff;
      return;
    }
    const int beg = off[seg], end = off[seg + 1];
    if (end - beg > AR_HEAVY_LEN) return;  // CTA path handles it
    RowTile<NV> acc;
    acc.zero();
    float q = 0.f;
    int j = beg;
    for (; j + 1 < end; j += 2) {  // two samples in flight
      const int s0 = order[j], s1 = order[j + 1];
      RowTile<NV> o0, o1;
      o0.load(other + (size_t)s0 * dim, d4, lane);
      o1.load(other + (size_t)s1 * dim, d4, lane);
      const int g0 = samp ? samp[s0] : s0, g1 = samp ? samp[s1] : s1;
      const float c0 = c[g0], c1 = c[g1];
      const float d0 = dc_of(dy[g0], c0, kk);
      const float d1 = dc_of(dy[g1], c1, kk);
      q = fmaf(d0, c0, q);
      q = fmaf(d1, c1, q);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        acc.x[k] = fma4(d0, o0.x[k], acc.x[k]);
        acc.x[k] = fma4(d1, o1.x[k], acc.x[k]);
      }
    }
    if (j < end) {
      const int s0 = order[j];
      RowTile<NV> o0;
      o0.load(other + (size_t)s0 * dim, d4, lane);
      const int g0 = samp ? samp[s0] : s0;
      const float c0 = c[g0];
      const float d0 = dc_of(dy[g0], c0, kk);
      q = fmaf(d0, c0, q);
#pragma unroll
      for (int k = 0; k < NV; ++k) acc.x[k] = fma4(d0, o0.x[k], acc.x[k]);
    }
    if (emit_map) {
      float* dst = emit_P + (size_t)emit_map[seg] * a.emit_stride;
      acc.store(dst, d4, lane);
      if (lane == 0) dst[dim] = q;
      return;
    }
    if (emit_ids) {
      acc.store(emit_P + (size_t)seg * dim, d4, lane);
      if (lane == 0) {
        emit_ids[seg] = uniq[seg];
        emit_q[seg] = q;
      }
      return;
    }
    finish_row<NV>(tb, uniq[seg], acc, q, rinv[order[beg]], alpha, l2x2, t, replay, a.reg, lane);
    return;
  }

  // heavy row: the whole CTA reduces one long segment; warp w takes samples beg+w, beg+w+8, ...
  if (b >= meta[1]) return;
  const int seg = heavy[b];
  const int beg = off[seg], end = off[seg + 1];
  RowTile<NV> acc;
  acc.zero();
  float q = 0.f;
  for (int j = beg + wid; j < end; j += kRowWarps) {
    const int s0 = order[j];
    RowTile<NV> o0;
    o0.load(other + (size_t)s0 * dim, d4, lane);
    const int g0 = samp ? samp[s0] : s0;
    const float c0 = c[g0];
    const float d0 = dc_of(dy[g0], c0, kk);
    q = fmaf(d0, c0, q);
#pragma unroll
    for (int k = 0; k < NV; ++k) acc.x[k] = fma4(d0, o0.x[k], acc.x[k]);
  }
  acc.store(red + (size_t)wid * dim, d4, lane);
  float* qred = red + (size_t)kRowWarps * dim;
  if (lane == 0) qred[wid] = q;
  __syncthreads();
  if (wid != 0) return;
  acc.zero();
  q = 0.f;
  for (int w8 = 0; w8 < kRowWarps; ++w8) {  // fixed order: deterministic
    RowTile<NV> p;
    p.load(red + (size_t)w8 * dim, d4, lane);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc.x[k].x += p.x[k].x; acc.x[k].y += p.x[k].y; acc.x[k].z += p.x[k].z; acc.x[k].w += p.x[k].w;
    }
    q += qred[w8];
  }
  if (emit_map) {
    float* dst = emit_P + (size_t)emit_map[seg] * a.emit_stride;
    acc.store(dst, d4, lane);
    if (lane == 0) dst[dim] = q;
    return;
  }
  if (emit_ids) {
    acc.store(emit_P + (size_t)seg * dim, d4, lane);
    if (lane == 0) {
      emit_ids[seg] = uniq[seg];
      emit_q[seg] = q;
    }
    return;
  }
  finish_row<NV>(tb, uniq[seg], acc, q, rinv[order[beg]], alpha, l2x2, t, replay, a.reg, lane);
}

// ---------------------------------------------------------------------------------------------
// inference forward / validation sums (Keras predict / test_step; oracle predict()/evaluate())
template <int NV, bool EVAL>
__global__ void __launch_bounds__(kRowThreads)
predict_kernel(const float* __restrict__ U, const float* __restrict__ A, int dim,
               const float* __restrict__ head, const float* __restrict__ bn_moving,
               const int32_t* __restrict__ iu, const int32_t* __restrict__ ia,
               const float* __restrict__ label, int64_t n, float* __restrict__ out,
               double* __restrict__ sums) {
  __shared__ double red[2 * kRowWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int d4 = dim >> 2;
  const float w = head[0], b = head[1], gamma = head[2], beta = head[3];
  const float mu = bn_moving[0];
  const float inv = 1.0f / sqrtf(bn_moving[1] + kBnEps);
  double sb = 0.0, sm = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * kRowWarps + wid; s < n; s += (int64_t)gridDim.x * kRowWarps) {
    RowTile<NV> u, a;
    u.load(U + (size_t)iu[s] * dim, d4, lane);
    a.load(A + (size_t)ia[s] * dim, d4, lane);
    const float su = tile_dot<NV>(u, u);
    const float sa = tile_dot<NV>(a, a);
    const float r_u = 1.0f / sqrtf(fmaxf(su, kL2NormEps));
    const float r_a = 1.0f / sqrtf(fmaxf(sa, kL2NormEps));
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      u.x[k] = scale4(u.x[k], r_u);
      a.x[k] = scale4(a.x[k], r_a);
    }
    const float cs = tile_dot<NV>(u, a);
    const float y = gamma * (((w * cs + b) - mu) * inv) + beta;
    const float p = sigmoidf_(y);
    if (EVAL) {
      const float tg = label[s];
      sb += (double)bce_logits(y, tg);
      sm += (double)((tg - p) * (tg - p));
    } else if (lane == 0) {
      out[s] = p;
    }
  }
  if (EVAL) {
    if (lane == 0) { red[wid] = sb; red[kRowWarps + wid] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {  // per-CTA partials; reduce_partials_kernel adds them in CTA order
      double a0 = 0.0, a1 = 0.0;
      for (int i = 0; i < kRowWarps; ++i) { a0 += red[i]; a1 += red[kRowWarps + i]; }
      sums[2 * blockIdx.x] = a0;
      sums[2 * blockIdx.x + 1] = a1;
    }
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ W, int64_t n4, double* out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = ld4_nc(W + 4 * i);
    s += (double)(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += red[i];
    out[blockIdx.x] = a;
  }
}

// out[j] += sum_i part[i*width + j] in a fixed order (one CTA): the deterministic second half of the reductions above
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ part, int n, int width,
                                                              double* __restrict__ out) {
  __shared__ double red[2 * 32];
  double v[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int j = 0; j < width; ++j) v[j] += part[(size_t)i * width + j];
  block_sum<2>(v, red);
  if (threadIdx.x == 0)
    for (int j = 0; j < width; ++j) out[j] += v[j];
}

// ---------------------------------------------------------------------------------------------
static int nv_of(int dim) { return (dim + 127) / 128; }
static bool dim_ok(int dim) { return dim > 0 && dim <= 512 && (dim % 4) == 0; }
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

#define AR_DISPATCH_NV(dim, ...)                     \
  switch (nv_of(dim)) {                              \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
    default: { constexpr int NV = 4; __VA_ARGS__; } break; \
  }

static int catch_threads(int dim) { return ((dim + 31) / 32) * 32; }   // one element per lane

static RegAcc reg_of(const ar_train_ctx& x) {
  RegAcc r{};
  if (x.reg_acc && x.stepw) {
    r.acc = x.reg_acc;
    r.stepw = x.stepw;
    r.scale = x.reg_scale > 0.f ? x.reg_scale : 1.f;
  }
  return r;
}

static int launch_catchup(const ar_table* t0, const ar_plan* p0, int slot0, const ar_table* t1,
                          const ar_plan* p1, int slot1, const float* alpha, float l2, int64_t t_target,
                          cudaStream_t st, bool skip_prev = false, int32_t* sched_ws = nullptr, RegAcc reg = RegAcc{}) {
  CatchupArgs a{};
  a.reg = reg;
  if (skip_prev) {  // look-ahead: leave the rows the previous step also touches to that step's update
    a.skip_flag[0] = p0->in_prev + (int64_t)slot0 * p0->batch_cap;
    if (t1) a.skip_flag[1] = p1->in_prev + (int64_t)slot1 * p1->batch_cap;
  }
  a.tab[0] = *t0;
  a.uniq[0] = p0->uniq + (int64_t)slot0 * p0->batch_cap;
  a.meta[0] = p0->meta + (int64_t)slot0 * 4;
  a.blocks0 = p0->batch_cap;
  int units = a.blocks0;
  if (t1) {
    a.tab[1] = *t1;
    a.uniq[1] = p1->uniq + (int64_t)slot1 * p1->batch_cap;
    a.meta[1] = p1->meta + (int64_t)slot1 * 4;
    units += p1->batch_cap;
  }
  a.cap = units;
  const float l2x2 = (float)(2.0 * (double)l2);
  if (sched_ws) {  // longest-first: classify into buckets, then replay in bucket order
    a.sched = sched_ws;
    AR_CUDA(cudaMemsetAsync(sched_ws + 3 * (size_t)units, 0, 4 * sizeof(int32_t), st));
    rows_classify_kernel<<<ceil_div(units, 256), 256, 0, st>>>(a, t_target, a.blocks0, t1 ? p1->batch_cap : 0);
    AR_LAUNCH_CHECK();
  }
  // (capping the resident catch-up CTAs per SM with dynamic shared memory, to leave warp slots for the step's
  // own kernels, was measured: 8..28 KB per CTA cost 0..10% of the step -- the replay wants the occupancy)
  rows_catchup_kernel<<<units, catch_threads(t0->dim), 0, st>>>(a, alpha, l2x2, t_target);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

static void fill_update(UpdateArgs& a, int w, const ar_table* tab, const ar_plan* p, int slot,
                        const float* other, const float* rinv, int batch_hint) {
  a.tab[w] = *tab;
  a.order[w] = p->order + (int64_t)slot * p->batch_cap;
  a.uniq[w] = p->uniq + (int64_t)slot * p->batch_cap;
  a.off[w] = p->off + (int64_t)slot * (p->batch_cap + 1);
  a.meta[w] = p->meta + (int64_t)slot * 4;
  a.heavy[w] = p->heavy + (int64_t)slot * p->heavy_cap;
  a.other[w] = other;
  a.rinv[w] = rinv;
  int b = batch_hint > 0 ? batch_hint : p->batch_cap;
  a.blocks_norm[w] = ceil_div(b, kRowWarps);
  a.blocks_heavy[w] = b / AR_HEAVY_LEN;  // at most this many segments can exceed AR_HEAVY_LEN
}

static int launch_update(UpdateArgs& a, bool two, const float* c, const float* dy, const float* stepc,
                         const float* alpha, float l2, int64_t t, int replay, RegAcc reg, cudaStream_t st) {
  if (!two) { a.blocks_norm[1] = 0; a.blocks_heavy[1] = 0; a.tab[1] = a.tab[0]; }
  a.reg = reg;
  int blocks = a.blocks_norm[0] + a.blocks_norm[1] + a.blocks_heavy[0] + a.blocks_heavy[1];
  const int dim = a.tab[0].dim;
  size_t smem = (size_t)kRowWarps * dim * sizeof(float) + kRowWarps * sizeof(float);
  const float l2x2 = (float)(2.0 * (double)l2);
  AR_DISPATCH_NV(dim, rows_update_kernel<NV><<<blocks, kRowThreads, smem, st>>>(a, c, dy, stepc, alpha, l2x2, t, replay));
  AR_LAUNCH_CHECK();
  return AR_OK;
}

static int launch_flush(const ar_table* tab, const float* alpha, float l2, int64_t t_target,
                        RegAcc reg, cudaStream_t st) {
  const float l2x2 = (float)(2.0 * (double)l2);
  int blocks = (int)std::min<int64_t>(ceil_div(tab->n_rows, kRowWarps), (int64_t)num_sms() * 8);
  if (blocks <= 0) return AR_OK;
  AR_DISPATCH_NV(tab->dim, table_flush_kernel<NV><<<blocks, kRowThreads, 0, st>>>(*tab, alpha, l2x2, t_target, reg));
  AR_LAUNCH_CHECK();
  return AR_OK;
}

}  // namespace ar

using namespace ar;

extern "C" int ar_table_flush(const ar_table* tab, const float* alpha, float l2, int64_t t_target,
                              unsigned long long* reg_acc, const float* stepw, float reg_scale, void* stream) {
  AR_REQUIRE(tab && alpha, "ar_table_flush: null pointer");
  AR_REQUIRE(dim_ok(tab->dim), "ar_table_flush: dim %d unsupported", tab->dim);
  RegAcc reg{};
  if (reg_acc && stepw) {
    reg.acc = reg_acc;
    reg.stepw = stepw;
    reg.scale = reg_scale > 0.f ? reg_scale : 1.f;
  }
  return launch_flush(tab, alpha, l2, t_target, reg, (cudaStream_t)stream);
}

extern "C" int ar_embed_fwd(const float* U, const float* A, int32_t dim, const int32_t* iu, const int32_t* ia,
                            int32_t n, float* uh, float* ah, float* c, float* ru, float* ra, void* stream) {
  AR_REQUIRE(U && A && iu && ia && c, "ar_embed_fwd: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_embed_fwd: dim %d unsupported", dim);
  if (n <= 0) return AR_OK;
  AR_DISPATCH_NV(dim, embed_fwd_kernel<NV><<<ceil_div(n, kRowWarps), kRowThreads, 0, (cudaStream_t)stream>>>(
                          U, A, dim, iu, ia, n, nullptr, uh, ah, c, ru, ra, nullptr));
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_head_step(const float* c, const float* label, int32_t n, float* head, float* head_m,
                            float* head_v, float* bn_moving, const float* alpha, int64_t t, float* dc,
                            float* metrics_row, void* stream) {
  AR_REQUIRE(c && label && head && head_m && head_v && bn_moving && alpha && dc, "ar_head_step: null pointer");
  AR_REQUIRE(n > 0, "ar_head_step: n must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const int nfp = ceil_div(n, kRowWarps), nhb = ceil_div(n, kHeadThreads);
  // stand-alone call: temporary scratch on the stream (ar_train_steps uses the caller's ctx buffers)
  const size_t bytes = (size_t)nfp * 16 + (size_t)nhb * 64 + (size_t)n * 4 + K_STEPC * 4 + 16;
  char* buf = nullptr;
  AR_CUDA(cudaMallocAsync((void**)&buf, bytes, st));
  double* fwd_part = (double*)buf;
  double* head_part = fwd_part + 2 * (size_t)nfp;
  float* dy = (float*)(head_part + 8 * (size_t)nhb);
  float* stepc = dy + n;
  unsigned int* ticket = (unsigned int*)(stepc + K_STEPC);
  AR_CUDA(cudaMemsetAsync(ticket, 0, 4, st));
  c_partials_kernel<<<ceil_div(nfp, 128), 128, 0, st>>>(c, n, fwd_part);
  AR_LAUNCH_CHECK();
  HeadIO io{head, head_m, head_v, bn_moving, alpha, dy, stepc, ticket, metrics_row ? metrics_row - t * 4 : nullptr};
  head_step_kernel<<<nhb, kHeadThreads, 0, st>>>(c, label, n, nullptr, fwd_part, io, t, head_part, 0);
  AR_LAUNCH_CHECK();
  dc_from_dy_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dy, c, n, stepc, dc);
  AR_LAUNCH_CHECK();
  AR_CUDA(cudaFreeAsync(buf, st));
  return AR_OK;
}

extern "C" int ar_rows_catchup(const ar_table* tab, const ar_plan* plan, int32_t slot, const float* alpha,
                               float l2, int64_t t, void* stream) {
  AR_REQUIRE(tab && plan && alpha, "ar_rows_catchup: null pointer");
  AR_REQUIRE(dim_ok(tab->dim), "ar_rows_catchup: dim %d unsupported", tab->dim);
  AR_REQUIRE(slot >= 0 && slot < plan->n_slots, "ar_rows_catchup: slot out of range");
  return launch_catchup(tab, plan, slot, nullptr, nullptr, 0, alpha, l2, t - 1, (cudaStream_t)stream);
}

extern "C" int ar_rows_update(const ar_table* tab, const ar_plan* plan, int32_t slot, const float* other_hat,
                              const float* c, const float* dy, const float* stepc, const float* rinv,
                              const float* alpha, float l2, int64_t t, int32_t replay, void* stream) {
  AR_REQUIRE(tab && plan && other_hat && c && dy && stepc && rinv && alpha, "ar_rows_update: null pointer");
  AR_REQUIRE(dim_ok(tab->dim), "ar_rows_update: dim %d unsupported", tab->dim);
  AR_REQUIRE(slot >= 0 && slot < plan->n_slots, "ar_rows_update: slot out of range");
  UpdateArgs a{};
  fill_update(a, 0, tab, plan, slot, other_hat, rinv, 0);
  return launch_update(a, false, c, dy, stepc, alpha, l2, t, replay, RegAcc{}, (cudaStream_t)stream);
}

namespace ar {
// Side stream + events of the look-ahead catch-up (one set per device, created on first use).
struct Lookahead {
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_upd[2] = {nullptr, nullptr};  // "row update of step s is complete" (recorded on the main stream)
  cudaEvent_t ev_ahead = nullptr;              // "look-ahead catch-up for the next step is complete" (side stream)
  cudaEvent_t ev_mid = nullptr;                // peer mode: "forward of step s is queued" (main stream)
  bool ok = false;
};
static Lookahead* lookahead() {
  static Lookahead la[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Lookahead& l = la[dev];
  if (!l.ok) {
    // lowest priority: when SM slots free up, the step's own (latency-bound) kernels get them first and the
    // SFU-bound replay of the NEXT steps fills the gaps
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (cudaStreamCreateWithPriority(&l.st2, cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
    for (int i = 0; i < 2; ++i)
      if (cudaEventCreateWithFlags(&l.ev_upd[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&l.ev_ahead, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&l.ev_mid, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    l.ok = true;
  }
  return &l;
}

static int check_ctx(const ar_train_ctx* ctx, int32_t slot0, int32_t n_steps) {
  AR_REQUIRE(ctx, "ar_train_steps: null ctx");
  const ar_train_ctx& x = *ctx;
  AR_REQUIRE(x.users.W && x.anime.W && x.head && x.alpha && x.iu && x.ia && x.label, "ar_train_steps: null pointer in ctx");
  AR_REQUIRE(x.users.dim == x.anime.dim && dim_ok(x.users.dim), "ar_train_steps: dim %d/%d unsupported", x.users.dim, x.anime.dim);
  AR_REQUIRE(x.batch > 0 && x.batch <= AR_MAX_BATCH, "ar_train_steps: batch %d outside (0,%d]", x.batch, AR_MAX_BATCH);
  AR_REQUIRE(x.uh && x.ah && x.c && x.ru && x.ra && x.dy && x.fwd_part && x.head_part && x.stepc && x.ticket && x.metrics,
             "ar_train_steps: null scratch");
  AR_REQUIRE(slot0 >= 0 && slot0 + n_steps <= x.plan_u.n_slots && slot0 + n_steps <= x.plan_a.n_slots,
             "ar_train_steps: plan slots [%d,%d) exceed plan size", slot0, slot0 + n_steps);
  AR_REQUIRE(x.mode >= AR_ADAM_REPLAY && x.mode <= AR_ADAM_TOUCHED, "ar_train_steps: bad mode %d", x.mode);
  return AR_OK;
}
}  // namespace ar

#include "chunk.inl"

namespace ar {
// the single-GPU entry point additionally needs the workspace and, in replay mode, the schedule
static int check_ctx_single(const ar_train_ctx* ctx, int32_t n_steps) {
  const ar_train_ctx& x = *ctx;
  AR_REQUIRE(x.chunk_ws && ((uintptr_t)x.chunk_ws & 255) == 0, "ar_train_steps: chunk_ws null or not 256-byte aligned");
  AR_REQUIRE(x.health, "ar_train_steps: null health");
  if (x.mode == AR_ADAM_REPLAY) {
    AR_REQUIRE(x.sched.codes && x.sched.glen && x.sched.sub && x.sched.cursor,
               "ar_train_steps: AR_ADAM_REPLAY needs the replay schedule (ar_plan_sched)");
    AR_REQUIRE(x.sched.n_slots >= n_steps && x.sched.cap >= x.plan_u.batch_cap + x.plan_a.batch_cap, "ar_train_steps: schedule too small");
    AR_REQUIRE(x.depth >= 1 && x.depth <= AR_SCHED_MAX_DEPTH, "ar_train_steps: depth %d outside [1,%d]", x.depth, AR_SCHED_MAX_DEPTH);
  }
  return AR_OK;
}
}  // namespace ar

extern "C" int ar_train_steps(const ar_train_ctx* ctx, int64_t epoch_step0, int32_t slot0, int64_t t0,
                              int32_t n_steps, void* stream) {
  int rc = check_ctx(ctx, slot0, n_steps);
  if (rc) return rc;
  if ((rc = check_ctx_single(ctx, n_steps))) return rc;
  AR_REQUIRE(slot0 == 0, "ar_train_steps: slot0 must be 0 (the step kernel indexes the plans by step)");
  const ar_train_ctx& x = *ctx;
  int live = 0;   // steps that actually hold samples
  for (int s = 0; s < n_steps; ++s)
    if ((epoch_step0 + s) * (int64_t)x.batch < x.n_samples) live = s + 1;
  if (live == 0) return AR_OK;
  return launch_chunk(x, epoch_step0, t0, live, (cudaStream_t)stream);
}

extern "C" int ar_chunk_ws_info(int32_t n_slots, int32_t batch_cap, int32_t dim, int64_t* out_host) {
  AR_REQUIRE(out_host && n_slots > 0 && batch_cap > 0 && dim_ok(dim), "ar_chunk_ws_info: bad arguments");
  const ChunkLayout l = chunk_layout(n_slots, batch_cap, dim);
  for (int i = 0; i < 8; ++i) out_host[i] = 0;
  out_host[0] = (int64_t)l.total;
  out_host[1] = (int64_t)l.stamps;
  out_host[2] = (int64_t)(l.ctl + offsetof(ChunkCtl, stats));
  out_host[3] = kStamps;
  return AR_OK;
}

#include "dist.inl"
#include "shard.inl"
#include "peer.inl"

extern "C" int ar_predict(const float* U, const float* A, int32_t dim, const float* head, const float* bn_moving,
                          const int32_t* iu, const int32_t* ia, int64_t n, float* out, void* stream) {
  AR_REQUIRE(U && A && head && bn_moving && iu && ia && out, "ar_predict: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_predict: dim %d unsupported", dim);
  if (n <= 0) return AR_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, kRowWarps), (int64_t)num_sms() * 8);
  AR_DISPATCH_NV(dim, predict_kernel<NV, false><<<blocks, kRowThreads, 0, (cudaStream_t)stream>>>(
                          U, A, dim, head, bn_moving, iu, ia, nullptr, n, out, nullptr));
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_eval_sums(const float* U, const float* A, int32_t dim, const float* head, const float* bn_moving,
                            const int32_t* iu, const int32_t* ia, const float* label, int64_t n, double* sums,
                            void* stream) {
  AR_REQUIRE(U && A && head && bn_moving && iu && ia && label && sums, "ar_eval_sums: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_eval_sums: dim %d unsupported", dim);
  if (n <= 0) return AR_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, kRowWarps), (int64_t)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  AR_CUDA(cudaMallocAsync((void**)&part, (size_t)blocks * 2 * sizeof(double), st));
  AR_DISPATCH_NV(dim, predict_kernel<NV, true><<<blocks, kRowThreads, 0, st>>>(
                          U, A, dim, head, bn_moving, iu, ia, label, n, nullptr, part));
  AR_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 256, 0, st>>>(part, blocks, 2, sums);
  AR_LAUNCH_CHECK();
  AR_CUDA(cudaFreeAsync(part, st));
  return AR_OK;
}

namespace ar {
// MUFU throughput probe: 8 independent sqrt -> add -> reciprocal chains per thread (the two special-function ops
// of one Adam element-step, adam1()), nothing else competing for the pipe.
__global__ void __launch_bounds__(256) sfu_probe_kernel(float* __restrict__ out, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = rcp_approx(__fadd_rn(sqrt_approx(x[i]), kAdamEps));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace ar

extern "C" int ar_bench_sfu(float* scratch, int32_t blocks, int32_t threads, int32_t iters, void* stream) {
  AR_REQUIRE(scratch && blocks > 0 && threads > 0 && threads <= 256 && iters > 0, "ar_bench_sfu: bad arguments");
  ar::sfu_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_sumsq(const float* W, int64_t n_elems, double* out, void* stream) {
  AR_REQUIRE(W && out, "ar_sumsq: null pointer");
  AR_REQUIRE(n_elems % 4 == 0, "ar_sumsq: n_elems must be a multiple of 4");
  if (n_elems == 0) return AR_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n_elems / 4, 256), (int64_t)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  AR_CUDA(cudaMallocAsync((void**)&part, (size_t)blocks * sizeof(double), st));
  sumsq_kernel<<<blocks, 256, 0, st>>>(W, n_elems / 4, part);
  AR_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 256, 0, st>>>(part, blocks, 1, out);
  AR_LAUNCH_CHECK();
  AR_CUDA(cudaFreeAsync(part, st));
  return AR_OK;
}
