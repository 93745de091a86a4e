// Half B, tensor-core part (placeholder until the tcgen05 kernel lands in this file).
#include "common.cuh"

extern "C" int64_t ar_allpairs_workspace(int64_t n_q, int32_t kprime) { return 0; }

extern "C" int ar_cosine_topk_allpairs(const void* Qn_bf16, int64_t q0, int64_t n_q, const void* Cn_bf16,
                                       int64_t c0, int64_t n_c, int64_t c_total, int32_t dim, int32_t kprime,
                                       int32_t exclude_self, const uint32_t* watched, int64_t watched_stride,
                                       float sign, int32_t* out_idx, float* out_score, void* workspace,
                                       void* stream) {
  ar::set_error("ar_cosine_topk_allpairs: not built yet");
  return AR_ERR_UNSUPPORTED;
}
