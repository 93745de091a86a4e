// Half B, memory-bound part: row normalisation, single-query cosine GEMV fused with a warp
// top-k (similar_anime.py:404-468, similar_users.py:293-312), partial-list merge and the exact
// fp32 re-rank used after the tensor-core candidate pass.
//
// Ranking order everywhere: higher score first, equal scores -> lower row index first; NaN scores
// (zero-norm rows) never rank (oracle/similarity.py rank_desc).
#include <algorithm>
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.cuh"

namespace ar {

constexpr int kTopkThreads = 256;
constexpr int kTopkWarps = kTopkThreads / 32;
constexpr int kMaxK = 32;
constexpr int kMergeThreads = 1024;
constexpr int kMergeWarps = kMergeThreads / 32;

__device__ __forceinline__ bool better(float as, int ai, float bs, int bi) {
  return (as > bs) || (as == bs && ai < bi);
}

// Sorted top-32 list distributed over a warp: lane l holds the l-th best (score, idx).
struct WarpList {
  float s;
  int i;
  __device__ __forceinline__ void init() {
    s = -CUDART_INF_F;
    i = 0x7fffffff;
  }
  // warp-uniform candidate; dedup = ignore a row id the list already holds
  __device__ __forceinline__ void insert(float cs, int ci, int lane, bool dedup = false) {
    if (dedup && __ballot_sync(0xffffffffu, i == ci)) return;
    const unsigned worse = __ballot_sync(0xffffffffu, better(cs, ci, s, i));
    if (!worse) return;
    const int pos = __ffs(worse) - 1;
    const float ps = __shfl_up_sync(0xffffffffu, s, 1);
    const int pi = __shfl_up_sync(0xffffffffu, i, 1);
    if (lane > pos) {
      s = ps;
      i = pi;
    } else if (lane == pos) {
      s = cs;
      i = ci;
    }
  }
  // every lane offers its own candidate (valid = has one); k-th entry is the admission threshold
  __device__ __forceinline__ void insert_lanes(float cs, int ci, bool valid, int k, int lane) {
    float ts = __shfl_sync(0xffffffffu, s, k - 1);
    int ti = __shfl_sync(0xffffffffu, i, k - 1);
    unsigned pend = __ballot_sync(0xffffffffu, valid && better(cs, ci, ts, ti));
    while (pend) {
      const int src = __ffs(pend) - 1;
      const float bs = __shfl_sync(0xffffffffu, cs, src);
      const int bi = __shfl_sync(0xffffffffu, ci, src);
      insert(bs, bi, lane);
      ts = __shfl_sync(0xffffffffu, s, k - 1);
      ti = __shfl_sync(0xffffffffu, i, k - 1);
      pend &= pend - 1;
      pend &= __ballot_sync(0xffffffffu, valid && better(cs, ci, ts, ti));
    }
  }
};

// Merge the lists of all warps of a CTA through shared memory; result in warp 0's list.
// sm_s / sm_i: [nwarps][32]
__device__ __forceinline__ void cta_merge(WarpList& wl, float* sm_s, int* sm_i, int k, int nwarps) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  sm_s[wid * 32 + lane] = wl.s;
  sm_i[wid * 32 + lane] = wl.i;
  __syncthreads();
  if (wid == 0) {
    for (int w = 1; w < nwarps; ++w) {
      const float cs = sm_s[w * 32 + lane];
      const int ci = sm_i[w * 32 + lane];
      wl.insert_lanes(cs, ci, lane < k && ci != 0x7fffffff, k, lane);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// resid (optional, with out_bf): ||w_hat - bf16(w_hat)||_2 per row, the rounding residual that bounds the
// bf16-operand score error of the tensor-core pass (|q~.c~ - q^.c^| <= resid_q + resid_c + resid_q*resid_c).
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ W, int64_t n, int dim,
                                                      float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf,
                                                      float* __restrict__ resid) {
  const int lane = threadIdx.x & 31;
  const int d4 = dim >> 2;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
    const float* src = W + row * dim;
    float ss = 0.f;
    for (int j = lane; j < d4; j += 32) {
      float4 x = ld4_nc(src + 4 * j);
      ss += dot4(x, x);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    float rs = 0.f;
    for (int j = lane; j < d4; j += 32) {
      float4 x = ld4(src + 4 * j);
      x.x /= nrm; x.y /= nrm; x.z /= nrm; x.w /= nrm;
      if (out) st4(out + row * dim + 4 * j, x);
      if (out_bf) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(x.z, x.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out_bf + row * dim + 4 * j) = pk;
        const float2 fl = __bfloat1622float2(lo), fh = __bfloat1622float2(hi);
        const float e0 = x.x - fl.x, e1 = x.y - fl.y, e2 = x.z - fh.x, e3 = x.w - fh.y;
        rs += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
    }
    if (resid) {
      rs = warp_sum(rs);
      if (lane == 0) resid[row] = sqrtf(rs) * 1.0000005f;  // rounded up: used as an upper bound
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Single-query GEMV + top-k.  8 lanes per row, 4 rows per warp iteration, 2 iterations in flight.
// NQ = float4 per lane = ceil(dim/4/8).
template <int NQ>
__global__ void __launch_bounds__(kTopkThreads)
query_topk_kernel(const float* __restrict__ W, int64_t n, int dim, int64_t q,
                  const uint32_t* __restrict__ cand_mask, int64_t exclude, int k,
                  int* __restrict__ part_idx, float* __restrict__ part_score) {
  __shared__ float sm_s[kTopkWarps * 32];
  __shared__ int sm_i[kTopkWarps * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  const int d4 = dim >> 2;

  // normalised query row chunk owned by this lane: q_hat = W[q] / ||W[q]||
  float4 qv[NQ];
  float qs = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int j = sub + 8 * i;
    qv[i] = (j < d4) ? ld4(W + q * dim + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    qs += dot4(qv[i], qv[i]);
  }
  qs += __shfl_xor_sync(0xffffffffu, qs, 1);
  qs += __shfl_xor_sync(0xffffffffu, qs, 2);
  qs += __shfl_xor_sync(0xffffffffu, qs, 4);
  const float qn = sqrtf(qs);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    qv[i].x /= qn; qv[i].y /= qn; qv[i].z /= qn; qv[i].w /= qn;
  }

  WarpList wl;
  wl.init();
  const int64_t warp_global = (int64_t)blockIdx.x * kTopkWarps + wid;
  const int64_t n_warps = (int64_t)gridDim.x * kTopkWarps;
  // software pipeline: the loads of the NEXT 4 rows are in flight while the current 4 are reduced and ranked
  auto load_rows = [&](int64_t r0, float4 (&x)[NQ]) {
    const int64_t row = r0 + grp;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int j = sub + 8 * i;
      x[i] = (row < n && j < d4) ? ld4_nc(W + row * dim + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 xa[NQ], xb[NQ];
  int64_t r0 = warp_global * 4;
  if (r0 < n) load_rows(r0, xa);
  float thr = -CUDART_INF_F;  // k-th best held by this warp (list admission bound), refreshed after inserts
  auto rank_rows = [&](int64_t rbase, const float4 (&x)[NQ]) {
    const int64_t row = rbase + grp;
    float dot = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      dot += dot4(x[i], qv[i]);
      ss += dot4(x[i], x[i]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const float sc = dot / sqrtf(ss);  // 0/0 -> NaN for zero rows, rejected below
    bool ok = (row < n) && (sc == sc) && (row != exclude);
    if (ok && cand_mask) ok = (cand_mask[row >> 5] >> (row & 31)) & 1u;
    // one candidate per 8-lane group: lanes 0, 8, 16, 24 offer theirs; a score below the warp's k-th best
    // cannot enter (ties go to the lower row id, which a later row never has), so most iterations stop here
    if (__any_sync(0xffffffffu, ok && sub == 0 && sc >= thr)) {
      wl.insert_lanes(sc, (int)row, ok && sub == 0, k, lane);
      thr = __shfl_sync(0xffffffffu, wl.s, k - 1);
    }
  };
  for (; r0 < n; r0 += n_warps * 8) {
    const int64_t r1 = r0 + n_warps * 4;
    if (r1 < n) load_rows(r1, xb);
    rank_rows(r0, xa);
    if (r1 >= n) break;
    const int64_t r2 = r1 + n_warps * 4;
    if (r2 < n) load_rows(r2, xa);
    rank_rows(r1, xb);
  }
  // CTA list = the k best of the 8 warp lists: every held entry counts how many of the others beat it (score desc,
  // row asc) and writes itself to that slot -- ~250 compares per thread instead of warp 0 inserting 7 lists serially
  sm_s[wid * 32 + lane] = wl.s;
  sm_i[wid * 32 + lane] = wl.i;
  const bool held = lane < k && wl.i != 0x7fffffff;
  const int m = __syncthreads_count(held);
  if (held) {
    int rank = 0;
    for (int w = 0; w < kTopkWarps; ++w)
      for (int l = 0; l < k; ++l) {
        const int oi = sm_i[w * 32 + l];
        const float os = sm_s[w * 32 + l];
        rank += (oi != 0x7fffffff && (os > wl.s || (os == wl.s && oi < wl.i))) ? 1 : 0;
      }
    if (rank < k) {
      part_idx[(int64_t)blockIdx.x * k + rank] = wl.i;
      part_score[(int64_t)blockIdx.x * k + rank] = wl.s;
    }
  }
  for (int r = m + (int)threadIdx.x; r < k; r += kTopkThreads) {
    part_idx[(int64_t)blockIdx.x * k + r] = -1;
    part_score[(int64_t)blockIdx.x * k + r] = -CUDART_INF_F;
  }
}

// Merge n_lists partial lists per query.  idx/score layout: [list][query][k_in], each list sorted best first.
// One CTA per query.  The k_out-th best of the union is at least the largest k_out-th entry of any single
// list, so everything below that bound T is dropped before the (serial, shuffle-heavy) list insertion:
// merging the 592 per-CTA lists of a single-query scan inserts ~k candidates instead of thousands.
__global__ void __launch_bounds__(kMergeThreads)
topk_merge_kernel(const int* __restrict__ idx, const float* __restrict__ score, int n_lists,
                  int64_t n_queries, int k_in, int k_out, int lists_sorted, int* __restrict__ out_idx,
                  float* __restrict__ out_score) {
  __shared__ float sm_s[kMergeWarps * 32];
  __shared__ int sm_i[kMergeWarps * 32];
  __shared__ float bound_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t qy = blockIdx.x;
  float tmax = -CUDART_INF_F;
  if (lists_sorted && k_in >= k_out) {
    for (int l = threadIdx.x; l < n_lists; l += kMergeThreads) {
      const int64_t o = ((int64_t)l * n_queries + qy) * k_in + (k_out - 1);
      const float sc = score[o];
      if (idx[o] >= 0 && sc == sc) tmax = fmaxf(tmax, sc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) sm_s[wid] = tmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = sm_s[0];
    for (int w = 1; w < kMergeWarps; ++w) t = fmaxf(t, sm_s[w]);
    bound_s = t;
  }
  __syncthreads();
  const float bound = bound_s;
  __syncthreads();
  const int total = n_lists * k_in;
  // Fast path (everything a single-query scan or an all-gather merge produces: total <= 8192 entries): an exact
  // radix select of the k_out-th largest score over order-preserving keys held in registers (8 passes of 4 bits,
  // 16-bin shared-memory histograms), then the few entries at or above it are ranked by counting how many beat them
  // (score desc, row id asc -- a strict order, every row sits in one list only).  No serial list insertion: the
  // per-list bound above still leaves thousands of entries when hundreds of short lists are merged.
  constexpr int kPerT = 8, kCap = 1024;
  __shared__ unsigned hist[16];
  __shared__ unsigned sel_prefix, sel_need;
  __shared__ float cand_s[kCap];
  __shared__ int cand_i[kCap];
  __shared__ int cand_n;
  if (total <= kMergeThreads * kPerT) {
    unsigned key[kPerT];
    int cid[kPerT];
    float csc[kPerT];
#pragma unroll
    for (int u = 0; u < kPerT; ++u) {
      const int e = u * kMergeThreads + threadIdx.x;
      key[u] = 0u;
      cid[u] = -1;
      csc[u] = -CUDART_INF_F;
      if (e < total) {
        const int l = e / k_in, j = e - l * k_in;
        const int64_t o = ((int64_t)l * n_queries + qy) * k_in + j;
        cid[u] = idx[o];
        csc[u] = score[o];
        if (cid[u] >= 0 && csc[u] == csc[u] && csc[u] >= bound) {
          const unsigned b = __float_as_uint(csc[u]);
          key[u] = ((b & 0x80000000u) ? ~b : (b | 0x80000000u)) | 1u;   // valid keys are never 0 (lowest bit is noise)
        } else {
          cid[u] = -1;
        }
      }
    }
    if (threadIdx.x == 0) {
      sel_prefix = 0u;
      sel_need = (unsigned)k_out;
      cand_n = 0;
    }
    __syncthreads();
    for (int shift = 28; shift >= 0; shift -= 4) {
      if (threadIdx.x < 16) hist[threadIdx.x] = 0u;
      __syncthreads();
      const unsigned prefix = sel_prefix;
#pragma unroll
      for (int u = 0; u < kPerT; ++u)
        if (key[u] && (shift == 28 || (key[u] >> (shift + 4)) == (prefix >> (shift + 4))))
          atomicAdd(&hist[(key[u] >> shift) & 15u], 1u);
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned need = sel_need, d = 0;
        for (int dd = 15; dd >= 0; --dd) {
          if (hist[dd] >= need) { d = (unsigned)dd; break; }
          need -= hist[dd];
          if (dd == 0) { d = 0; need = 0; }      // fewer valid entries than k_out: keep them all
        }
        sel_need = need;
        sel_prefix = prefix | (d << shift);
      }
      __syncthreads();
    }
    const unsigned kth = sel_need ? sel_prefix : 0u;
#pragma unroll
    for (int u = 0; u < kPerT; ++u)
      if (key[u] && key[u] >= (kth & ~1u)) {       // the key's lowest bit was forced: compare without it
        const int pos = atomicAdd(&cand_n, 1);
        if (pos < kCap) {
          cand_s[pos] = csc[u];
          cand_i[pos] = cid[u];
        }
      }
    __syncthreads();
    const int m = cand_n;
    if (m <= kCap) {
      if ((int)threadIdx.x < m) {
        const float cs = cand_s[threadIdx.x];
        const int ci = cand_i[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
          const float os = cand_s[j];
          rank += (os > cs || (os == cs && cand_i[j] < ci)) ? 1 : 0;
        }
        if (rank < k_out) {
          out_idx[qy * k_out + rank] = ci;
          out_score[qy * k_out + rank] = cs;
        }
      }
      for (int r = m + threadIdx.x; r < k_out; r += kMergeThreads) {
        out_idx[qy * k_out + r] = -1;
        out_score[qy * k_out + r] = -CUDART_INF_F;
      }
      return;
    }
    __syncthreads();
  }
  WarpList wl;
  wl.init();
  // every thread fetches its (up to 8) entries of a pass BEFORE any is ranked: one exposed load latency per
  // pass instead of one per entry (the 592 x 11 lists of a single-query scan are one pass of 1024 threads)
  constexpr int kPer = 8;
  for (int p0 = 0; p0 < total; p0 += kMergeThreads * kPer) {
    float cs[kPer];
    int ci[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int e = p0 + (u * kMergeWarps + wid) * 32 + lane;
      cs[u] = -CUDART_INF_F;
      ci[u] = -1;
      if (e < total) {
        const int l = e / k_in, j = e - l * k_in;
        const int64_t o = ((int64_t)l * n_queries + qy) * k_in + j;
        ci[u] = idx[o];
        cs[u] = score[o];
      }
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const bool valid = ci[u] >= 0 && cs[u] == cs[u] && cs[u] >= bound;
      if (__any_sync(0xffffffffu, valid)) wl.insert_lanes(cs[u], ci[u], valid, k_out, lane);
    }
  }
  cta_merge(wl, sm_s, sm_i, k_out, kMergeWarps);
  if (wid == 0 && lane < k_out) {
    out_idx[qy * k_out + lane] = (wl.i == 0x7fffffff) ? -1 : wl.i;
    out_score[qy * k_out + lane] = wl.s;
  }
}

// Exact fp32 re-rank: warp per query row.  Candidates come as n_lists lists per query (layout
// [list][query][list_cap]; cand_cnt[list][query] entries are valid, or every non-negative entry when cand_cnt
// is null), e.g. the per-chunk lists of the tensor-core pass.
// Certification (optional): list l left out only candidates whose selection (bf16-operand) score is
// <= cand_thr[l][query]; such a row has fp32 score <= that + eps.  The fp32 top-k is therefore provably the
// exact top-k over ALL rows the lists were drawn from when the k-th re-ranked score clears
// max_l cand_thr[l] + eps (+ q_eps[query]); flag = 1 then, else 0.
template <int NV>
__global__ void __launch_bounds__(kTopkThreads)
rerank_kernel(const float* __restrict__ Wq, int64_t q0, int64_t n_queries, const float* __restrict__ Wc,
              int dim, const int* __restrict__ cand, const int* __restrict__ cand_cnt,
              const float* __restrict__ cand_thr, int n_lists, int list_cap, int k, float eps,
              const float* __restrict__ q_eps, int* __restrict__ out_idx, float* __restrict__ out_score,
              unsigned char* __restrict__ certified) {
  const int lane = threadIdx.x & 31;
  const int d4 = dim >> 2;
  const int64_t qi = (int64_t)blockIdx.x * kTopkWarps + (threadIdx.x >> 5);
  if (qi >= n_queries) return;
  float4 qv[NV];
  float qs = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    qv[i] = (j < d4) ? ld4(Wq + (q0 + qi) * dim + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    qs += dot4(qv[i], qv[i]);
  }
  const float qn = sqrtf(warp_sum(qs));
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    qv[i].x /= qn; qv[i].y /= qn; qv[i].z /= qn; qv[i].w /= qn;
  }
  WarpList wl;
  wl.init();
  float bound = -CUDART_INF_F;
  for (int l = 0; l < n_lists; ++l) {
    const int64_t lq = (int64_t)l * n_queries + qi;
    const int* cl = cand + lq * list_cap;
    const int cnt = cand_cnt ? min(cand_cnt[lq], list_cap) : list_cap;
    if (cand_thr) bound = fmaxf(bound, cand_thr[lq]);
    for (int j0 = 0; j0 < cnt; j0 += 32) {
      const int mine = (j0 + lane < cnt) ? cl[j0 + lane] : -1;
      const int m = min(32, cnt - j0);
      for (int j = 0; j < m; j += 2) {  // two candidate rows in flight
        const int r0 = __shfl_sync(0xffffffffu, mine, j);
        const int r1 = (j + 1 < m) ? __shfl_sync(0xffffffffu, mine, j + 1) : -1;
        float d0 = 0.f, s0 = 0.f, d1 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int jj = lane + 32 * i;
          if (jj < d4) {
            const float4 x0 = (r0 >= 0) ? ld4_nc(Wc + (int64_t)r0 * dim + 4 * jj) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x1 = (r1 >= 0) ? ld4_nc(Wc + (int64_t)r1 * dim + 4 * jj) : make_float4(0.f, 0.f, 0.f, 0.f);
            d0 += dot4(x0, qv[i]); s0 += dot4(x0, x0);
            d1 += dot4(x1, qv[i]); s1 += dot4(x1, x1);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d0 += __shfl_xor_sync(0xffffffffu, d0, o); s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          d1 += __shfl_xor_sync(0xffffffffu, d1, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (r0 >= 0) {
          const float sc0 = d0 / sqrtf(s0);
          if (sc0 == sc0) wl.insert(sc0, r0, lane, true);
        }
        if (r1 >= 0) {
          const float sc1 = d1 / sqrtf(s1);
          if (sc1 == sc1) wl.insert(sc1, r1, lane, true);
        }
      }
    }
  }
  if (lane < k) {
    out_idx[qi * k + lane] = (wl.i == 0x7fffffff) ? -1 : wl.i;
    out_score[qi * k + lane] = wl.s;
  }
  if (certified) {
    const float kth = __shfl_sync(0xffffffffu, wl.s, k - 1);  // -inf when fewer than k candidates exist
    const float e = eps + (q_eps ? q_eps[qi] : 0.f);
    if (lane == 0) certified[qi] = (bound == -CUDART_INF_F || kth >= bound + e) ? 1 : 0;
  }
}

// watched-bit matrix for the scoring path: set bit idx[j] of row r for j in [indptr[r], indptr[r+1])
__global__ void bits_from_csr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                     int64_t n_rows, int64_t stride_words, int64_t n_bits,
                                     uint32_t* __restrict__ out) {
  const int64_t r = blockIdx.x;
  if (r >= n_rows) return;
  for (int64_t j = indptr[r] + threadIdx.x; j < indptr[r + 1]; j += blockDim.x) {
    const int c = idx[j];
    if (c >= 0 && c < n_bits) atomicOr(out + r * stride_words + (c >> 5), 1u << (c & 31));
  }
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
static bool dim_ok(int dim) { return dim > 0 && dim <= 512 && (dim % 4) == 0; }
// Grid of the single-query scan: exactly one wave (SMs x resident CTAs of the instantiation) -- with 80
// registers per thread only 3 CTAs fit an SM, and a 4-per-SM grid ran as 1.33 waves (ncu: warps active 30 %).
constexpr int kMaxQueryCtasPerSm = 8;
template <int NQ>
static int query_occupancy() {
  static int occ = 0;
  if (!occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, query_topk_kernel<NQ>, kTopkThreads, 0) != cudaSuccess || occ <= 0) occ = 2;
    occ = std::min(occ, kMaxQueryCtasPerSm);
  }
  return occ;
}
static int query_blocks(int64_t n_rows, int occ) {
  int64_t need = (n_rows + kTopkWarps * 4 - 1) / (kTopkWarps * 4);
  return (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)sm_count() * occ));
}

}  // namespace ar

using namespace ar;

extern "C" int ar_rownorm(const float* W, int64_t n_rows, int32_t dim, float* out, void* stream) {
  AR_REQUIRE(W && out, "ar_rownorm: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_rownorm: dim %d unsupported", dim);
  if (n_rows <= 0) return AR_OK;
  int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)sm_count() * 8);
  rownorm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, n_rows, dim, out, nullptr, nullptr);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_rownorm_bf16(const float* W, int64_t n_rows, int32_t dim, void* out_bf16, float* resid,
                               void* stream) {
  AR_REQUIRE(W && out_bf16, "ar_rownorm_bf16: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_rownorm_bf16: dim %d unsupported", dim);
  if (n_rows <= 0) return AR_OK;
  int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)sm_count() * 8);
  rownorm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, n_rows, dim, nullptr, (__nv_bfloat16*)out_bf16, resid);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int64_t ar_topk_query_workspace(int64_t n_rows, int32_t k) {
  if (k <= 0 || k > kMaxK || n_rows <= 0) return 0;
  return (int64_t)query_blocks(n_rows, kMaxQueryCtasPerSm) * k * 8;
}

extern "C" int ar_cosine_topk_query(const float* W, int64_t n_rows, int32_t dim, int64_t q,
                                    const uint32_t* cand_mask, int64_t exclude, int32_t k, int32_t* out_idx,
                                    float* out_score, void* workspace, void* stream) {
  AR_REQUIRE(W && out_idx && out_score && workspace, "ar_cosine_topk_query: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_cosine_topk_query: dim %d unsupported", dim);
  AR_REQUIRE(k > 0 && k <= kMaxK, "ar_cosine_topk_query: k %d outside [1,%d]", k, kMaxK);
  AR_REQUIRE(n_rows > 0 && q >= 0 && q < n_rows, "ar_cosine_topk_query: query row %lld outside [0,%lld)", (long long)q, (long long)n_rows);
  cudaStream_t st = (cudaStream_t)stream;
  const int nq = ((dim >> 2) + 7) / 8;
  const int occ = nq <= 4 ? query_occupancy<4>() : (nq <= 8 ? query_occupancy<8>() : query_occupancy<16>());
  const int blocks = query_blocks(n_rows, occ);
  int* pidx = (int*)workspace;
  float* pscore = (float*)workspace + (int64_t)blocks * k;
  if (nq <= 4) query_topk_kernel<4><<<blocks, kTopkThreads, 0, st>>>(W, n_rows, dim, q, cand_mask, exclude, k, pidx, pscore);
  else if (nq <= 8) query_topk_kernel<8><<<blocks, kTopkThreads, 0, st>>>(W, n_rows, dim, q, cand_mask, exclude, k, pidx, pscore);
  else query_topk_kernel<16><<<blocks, kTopkThreads, 0, st>>>(W, n_rows, dim, q, cand_mask, exclude, k, pidx, pscore);
  AR_LAUNCH_CHECK();
  topk_merge_kernel<<<1, kMergeThreads, 0, st>>>(pidx, pscore, blocks, 1, k, k, 1, out_idx, out_score);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

// Many exact fp32 queries in one call (embedding sizes the tensor-core pass is not built for): query i is row
// q0 + i, excluded from its own result; the launches are queued back to back on the stream from C, so a whole
// table costs no host round trips (similarity.allpairs_topk used to loop in Python here).
extern "C" int ar_cosine_topk_queries(const float* W, int64_t n_rows, int32_t dim, int64_t q0, int64_t n_queries,
                                      const uint32_t* cand_mask, int32_t k, int32_t* out_idx, float* out_score,
                                      void* workspace, void* stream) {
  AR_REQUIRE(q0 >= 0 && n_queries >= 0 && q0 + n_queries <= n_rows, "ar_cosine_topk_queries: query range outside the table");
  for (int64_t i = 0; i < n_queries; ++i) {
    const int rc = ar_cosine_topk_query(W, n_rows, dim, q0 + i, cand_mask, q0 + i, k, out_idx + i * k, out_score + i * k,
                                        workspace, stream);
    if (rc) return rc;
  }
  return AR_OK;
}

extern "C" int ar_topk_merge(const int32_t* idx, const float* score, int32_t n_lists, int64_t n_queries,
                             int32_t k_in, int32_t k_out, int32_t lists_sorted, int32_t* out_idx, float* out_score,
                             void* stream) {
  AR_REQUIRE(idx && score && out_idx && out_score, "ar_topk_merge: null pointer");
  AR_REQUIRE(k_out > 0 && k_out <= kMaxK && k_in > 0 && n_lists > 0, "ar_topk_merge: bad k/n_lists");
  if (n_queries <= 0) return AR_OK;
  topk_merge_kernel<<<(unsigned)n_queries, kMergeThreads, 0, (cudaStream_t)stream>>>(idx, score, n_lists, n_queries, k_in, k_out, lists_sorted, out_idx, out_score);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_cosine_rerank(const float* Wq, int64_t q0, int64_t n_queries, const float* Wc, int32_t dim,
                                const int32_t* cand, const int32_t* cand_cnt, const float* cand_thr,
                                int32_t n_lists, int32_t list_cap, int32_t k, float eps, const float* q_eps,
                                int32_t* out_idx, float* out_score, uint8_t* certified, void* stream) {
  AR_REQUIRE(Wq && Wc && cand && out_idx && out_score, "ar_cosine_rerank: null pointer");
  AR_REQUIRE(dim_ok(dim), "ar_cosine_rerank: dim %d unsupported", dim);
  AR_REQUIRE(k > 0 && k <= kMaxK && n_lists > 0 && list_cap > 0, "ar_cosine_rerank: bad k/n_lists/list_cap");
  AR_REQUIRE(!certified || cand_thr, "ar_cosine_rerank: certification needs the lists' admission thresholds");
  if (n_queries <= 0) return AR_OK;
  const int blocks = (int)((n_queries + kTopkWarps - 1) / kTopkWarps);
  cudaStream_t st = (cudaStream_t)stream;
#define AR_RERANK(NVV) rerank_kernel<NVV><<<blocks, kTopkThreads, 0, st>>>(Wq, q0, n_queries, Wc, dim, cand, cand_cnt, cand_thr, \
                                                                            n_lists, list_cap, k, eps, q_eps, out_idx, out_score, certified)
  switch ((dim + 127) / 128) {
    case 1: AR_RERANK(1); break;
    case 2: AR_RERANK(2); break;
    case 3: AR_RERANK(3); break;
    default: AR_RERANK(4); break;
  }
#undef AR_RERANK
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_bits_from_csr(const int64_t* indptr, const int32_t* idx, int64_t n_rows, int64_t stride_words,
                                int64_t n_bits, uint32_t* out, void* stream) {
  AR_REQUIRE(indptr && idx && out, "ar_bits_from_csr: null pointer");
  AR_REQUIRE(stride_words * 32 >= n_bits, "ar_bits_from_csr: stride_words too small for n_bits");
  if (n_rows <= 0) return AR_OK;
  bits_from_csr_kernel<<<(unsigned)n_rows, 128, 0, (cudaStream_t)stream>>>(indptr, idx, n_rows, stride_words, n_bits, out);
  AR_LAUNCH_CHECK();
  return AR_OK;
}
