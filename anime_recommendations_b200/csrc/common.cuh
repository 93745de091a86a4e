// Shared helpers for libanimerec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "animerec.h"

namespace ar {

void set_error(const char* fmt, ...);

#define AR_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ar::set_error(__VA_ARGS__);        \
      return AR_ERR_INVALID;             \
    }                                    \
  } while (0)

#define AR_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ar::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return AR_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define AR_LAUNCH_CHECK() AR_CUDA(cudaGetLastError())

constexpr int kWarp = 32;
constexpr float kBeta1 = 0.9f;
constexpr float kBeta2 = 0.999f;
// (1 - beta) evaluated in double then rounded, as Keras/NumPy do (oracle/train.py _adam_apply)
constexpr float kOneMinusBeta1 = (float)(1.0 - 0.9);
constexpr float kOneMinusBeta2 = (float)(1.0 - 0.999);
constexpr float kAdamEps = 1e-7f;
constexpr float kBnEps = 1e-3f;
constexpr float kBnOneMinusMomentum = (float)(1.0 - 0.99);
constexpr float kL2NormEps = 1e-12f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming variants: table rows are touched once per step, keep them out of L1
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld4_nc(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 acc) {
  return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One Keras-2.12 Adam update of a single element (oracle/train.py::_adam_apply, assumption A6):
//   m += (g-m)(1-b1);  v += (g^2-v)(1-b2);  w -= (m*alpha)/(sqrt(v)+eps)
// Every operation is spelled as an explicit intrinsic (no compiler contraction), so the SAME bits come
// out wherever this is inlined -- that is what makes the deferred replay bit-identical to the dense
// mode.  sqrt and the reciprocal use the MUFU approximations (<= 2 ulp each, ~3e-7 relative on a step of
// size ~alpha, i.e. ~1e-11 absolute): 2 MUFU + 9 FP32 ops per element-step instead of ~40 for the IEEE
// sequences, which is what bounds the replay (DESIGN.md "rows_catchup").
__device__ __forceinline__ void adam1(float& w, float& m, float& v, float g, float alpha) {
  m = __fmaf_rn(__fsub_rn(g, m), kOneMinusBeta1, m);
  v = __fmaf_rn(__fmaf_rn(g, g, -v), kOneMinusBeta2, v);
  const float r = rcp_approx(__fadd_rn(sqrt_approx(v), kAdamEps));
  w = __fmaf_rn(-__fmul_rn(m, alpha), r, w);
}
__device__ __forceinline__ void adam4(float4& w, float4& m, float4& v, float4 g, float alpha) {
  adam1(w.x, m.x, v.x, g.x, alpha);
  adam1(w.y, m.y, v.y, g.y, alpha);
  adam1(w.z, m.z, v.z, g.z, alpha);
  adam1(w.w, m.w, v.w, g.w, alpha);
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace ar
