// Collaborative aggregation of user_recs (user_recs.py:348-387, 708-794) as a batched GPU job:
//   favourites of a user   = the ratings at or above the user's 80th percentile (np.percentile, linear
//                            interpolation) -- user_recs.py:359-361
//   recommendations of q   = the anime most often among the favourites of q's similar users, excluding q's own
//                            favourites, ranked by that count (value_counts, user_recs.py:761-774)
// Ratings come as a CSR by user (indptr, anime index, rating).  Integer work except for the percentile, which is
// evaluated in double exactly as NumPy does.
#include <algorithm>

#include "common.cuh"

namespace ar {

// order-preserving uint32 key of a float
__device__ __forceinline__ unsigned f2key(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of the warp's strided view of x[0..n): bitwise radix select, 32 passes of n/32 steps.
__device__ float warp_select(const float* __restrict__ x, int64_t n, int64_t k, int lane) {
  unsigned prefix = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    // how many keys share the decided prefix and have a 0 in this bit?
    const unsigned mask_hi = bit == 31 ? 0u : (0xffffffffu << (bit + 1));
    int64_t zeros = 0;
    for (int64_t i = lane; i < n; i += 32) {
      const unsigned key = f2key(x[i]);
      zeros += ((key & mask_hi) == prefix && !((key >> bit) & 1u)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if (k >= zeros) {
      k -= zeros;
      prefix |= 1u << bit;
    }
  }
  return key2f(prefix);
}

// Warp per user: threshold = np.percentile(ratings, q) with NumPy's default (linear) method:
//   virtual = (n-1) * (q/100); lo = floor(virtual); g = virtual - lo;
//   thr = a + (b-a)*g  (g < 0.5)   or   b - (b-a)*(1-g)  (g >= 0.5)        [numpy _lerp]
// then flag every rating >= thr (compared in double, user_recs.py:361).
__global__ void __launch_bounds__(256)
user_fav_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ rating, int n_users, double q,
                uint8_t* __restrict__ fav, double* __restrict__ thr_out) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= n_users) return;
  const int64_t beg = indptr[u], n = indptr[u + 1] - beg;
  if (n <= 0) {
    if (lane == 0 && thr_out) thr_out[u] = 0.0;
    return;
  }
  const float* x = rating + beg;
  const double virt = __dmul_rn((double)(n - 1), q / 100.0);
  const int64_t lo = (int64_t)floor(virt);
  const double g = virt - (double)lo;
  const double a = (double)warp_select(x, n, lo, lane);
  const double b = (lo + 1 < n) ? (double)warp_select(x, n, lo + 1, lane) : a;
  const double d = b - a;
  // explicit roundings (no FMA contraction): the same two operations NumPy performs
  const double thr = g >= 0.5 ? __dsub_rn(b, __dmul_rn(d, 1.0 - g)) : __dadd_rn(a, __dmul_rn(d, g));
  for (int64_t i = lane; i < n; i += 32) fav[beg + i] = ((double)x[i] >= thr) ? 1 : 0;
  if (lane == 0 && thr_out) thr_out[u] = thr;
}

struct RecsArgs {
  const int64_t* indptr;
  const int32_t* anime;
  const uint8_t* fav;
  int n_anime;
  const int32_t* query;
  const int32_t* sim;   // [n_query][k_sim], entries < 0 are padding
  int k_sim, n_recs;
  int32_t* out_idx;     // [n_query][n_recs], -1 beyond the number of recommendations
  int32_t* out_cnt;
};

// CTA per query user.  counts[] (16 bit, two per word) in shared memory; the ranking is by (count desc, anime
// index asc): counts are at most k_sim, so one ordered compaction pass per count level emits the list.
constexpr int kRecsThreads = 256;
__global__ void __launch_bounds__(kRecsThreads) user_recs_kernel(RecsArgs a) {
  extern __shared__ unsigned cnt2[];   // ceil(n_anime/2) words
  __shared__ int warp_tot[kRecsThreads / 32];
  __shared__ int base_s;
  const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int words = (a.n_anime + 1) >> 1;
  for (int i = tid; i < words; i += kRecsThreads) cnt2[i] = 0u;
  __syncthreads();
  for (int j = 0; j < a.k_sim; ++j) {
    const int u = a.sim[(size_t)qi * a.k_sim + j];
    if (u < 0) continue;
    const int64_t beg = a.indptr[u], end = a.indptr[u + 1];
    for (int64_t i = beg + tid; i < end; i += kRecsThreads)
      if (a.fav[i]) {
        const int an = a.anime[i];
        atomicAdd(&cnt2[an >> 1], 1u << (16 * (an & 1)));
      }
  }
  __syncthreads();
  {  // the query's own favourites are never recommended (user_recs.py:757)
    const int u = a.query[qi];
    const int64_t beg = a.indptr[u], end = a.indptr[u + 1];
    for (int64_t i = beg + tid; i < end; i += kRecsThreads)
      if (a.fav[i]) {
        const int an = a.anime[i];
        atomicAnd(&cnt2[an >> 1], (an & 1) ? 0x0000ffffu : 0xffff0000u);
      }
  }
  __syncthreads();
  int32_t* oi = a.out_idx + (size_t)qi * a.n_recs;
  int32_t* oc = a.out_cnt + (size_t)qi * a.n_recs;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int level = a.k_sim; level >= 1; --level) {
    if (base_s >= a.n_recs) break;
    for (int i0 = 0; i0 < a.n_anime; i0 += kRecsThreads) {
      const int i = i0 + tid;
      const bool hit = i < a.n_anime && (int)((cnt2[i >> 1] >> (16 * (i & 1))) & 0xffffu) == level;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) warp_tot[wid] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kRecsThreads / 32; ++w) {
        if (w < wid) before += warp_tot[w];
        total += warp_tot[w];
      }
      const int base = base_s;
      if (hit) {
        const int pos = base + before + __popc(m & ((1u << lane) - 1u));
        if (pos < a.n_recs) {
          oi[pos] = i;
          oc[pos] = level;
        }
      }
      __syncthreads();
      if (tid == 0) base_s = base + total;
      __syncthreads();
      if (base_s >= a.n_recs) break;
    }
  }
  for (int p = min(base_s, a.n_recs) + tid; p < a.n_recs; p += kRecsThreads) {
    oi[p] = -1;
    oc[p] = 0;
  }
}

}  // namespace ar

extern "C" int ar_user_favourites(const int64_t* indptr, const float* rating, int32_t n_users, double percentile,
                                  uint8_t* fav_flag, double* thr_out, void* stream) {
  AR_REQUIRE(indptr && rating && fav_flag, "ar_user_favourites: null pointer");
  AR_REQUIRE(percentile >= 0.0 && percentile <= 100.0, "ar_user_favourites: percentile %f outside [0,100]", percentile);
  if (n_users <= 0) return AR_OK;
  ar::user_fav_kernel<<<ar::ceil_div(n_users, 8), 256, 0, (cudaStream_t)stream>>>(indptr, rating, n_users, percentile,
                                                                                 fav_flag, thr_out);
  AR_LAUNCH_CHECK();
  return AR_OK;
}

extern "C" int ar_user_recs(const int64_t* indptr, const int32_t* anime_idx, const uint8_t* fav_flag, int32_t n_anime,
                            const int32_t* query_users, int32_t n_query, const int32_t* sim_users, int32_t k_sim,
                            int32_t n_recs, int32_t* out_idx, int32_t* out_cnt, void* stream) {
  AR_REQUIRE(indptr && anime_idx && fav_flag && query_users && sim_users && out_idx && out_cnt, "ar_user_recs: null pointer");
  AR_REQUIRE(n_anime > 0 && k_sim > 0 && k_sim < 1024 && n_recs > 0, "ar_user_recs: bad sizes (n_anime %d, k_sim %d, n_recs %d)",
             n_anime, k_sim, n_recs);
  const size_t smem = (size_t)((n_anime + 1) / 2) * sizeof(unsigned);
  AR_REQUIRE(smem <= 200 * 1024, "ar_user_recs: %d anime do not fit the shared-memory counters", n_anime);
  if (n_query <= 0) return AR_OK;
  static bool attr_set = false;
  if (!attr_set) {
    AR_CUDA(cudaFuncSetAttribute(ar::user_recs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  ar::RecsArgs a{indptr, anime_idx, fav_flag, n_anime, query_users, sim_users, k_sim, n_recs, out_idx, out_cnt};
  ar::user_recs_kernel<<<n_query, ar::kRecsThreads, smem, (cudaStream_t)stream>>>(a);
  AR_LAUNCH_CHECK();
  return AR_OK;
}
