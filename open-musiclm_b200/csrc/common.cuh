// Shared host/device helpers for libomlm_b200: error plumbing, TMA descriptor cache, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace omlm {

// ---- error reporting across the C ABI (no exceptions cross the boundary) -----------------
void set_last_error(const char* fmt, ...);
#define OMLM_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::omlm::set_last_error(__VA_ARGS__); \
      return 1;                            \
    }                                      \
  } while (0)
#define OMLM_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::omlm::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1000 + static_cast<int>(_e);                                                     \
    }                                                                                         \
  } while (0)
#define OMLM_LAUNCH_CHECK() OMLM_CUDA(cudaGetLastError())

// ---- TMA descriptors ----------------------------------------------------------------------
// 2-D bf16 tensor map: inner (contiguous) extent dim0, outer extent dim1, row pitch in bytes,
// box {box0, box1}, SWIZZLE_128B (box0 * 2 bytes must be 128).
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t dim0, uint64_t dim1,
                      uint64_t pitch_bytes, uint32_t box0, uint32_t box1);

// 3-D bf16 tensor map (dim0 contiguous), SWIZZLE_128B, box {box0, box1, 1}: out-of-range rows/planes are clipped.
int make_tmap_bf16_3d(CUtensorMap* out, const void* gptr, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0, uint32_t box1);

int num_sms();

// ---- device math --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// ---- GELU (exact erf) ------------------------------------------------------------------------
// Exact-erf GELU pieces from ONE exponential: e = exp(-x^2/2) gives both the normal pdf and, through the
// Abramowitz-Stegun 7.1.26 rational form (|error| <= 1.5e-7, below fp32 resolution of the products here),
// erf(x/sqrt(2)).  cdf = Phi(x), pdf = phi(x);  gelu(x) = x*cdf, gelu'(x) = cdf + x*pdf.
__device__ __forceinline__ void normal_cdf_pdf(float x, float& cdf, float& pdf) {
  const float e = __expf(-0.5f * x * x);
  const float ax = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float half_erfc = 0.5f * p * t * e;            // 0.5 * erfc(|x|/sqrt2)
  cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float c, p;
  normal_cdf_pdf(x, c, p);
  return x * c;
}

// Philox-4x32 counter RNG, 7 rounds (the shortest variant that passes BigCrush): dropout / forgetful-mask
// randomness, replayable in the backward pass from (seed, layer, row, chunk).
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

}  // namespace omlm
